"""Warm per-GEMM times of the CAMA layer shapes at b = 1 (M = 250) for the K5 tile variants.
Run on the GPU box: python scripts/cama_gemm_bench.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from motionrag_b200 import cama  # noqa: E402

M = int(os.environ.get("M", 250))
SHAPES = [("qkv", 3072, 1024, (1,)), ("out", 1024, 1024, (1, 2, 4, 8)), ("ffn1", 4096, 1024, (1,)),
          ("ffn2", 1024, 4096, (1, 4, 8))]
torch.manual_seed(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, N, K, split_opts in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda").bfloat16()
    for bn in ("64", "128"):
        for sp in split_opts:
            os.environ["MRAG_K5_BN"] = bn
            try:
                for _ in range(5):
                    cama.linear(a, w, bias if sp == 1 else None, splits=sp)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(20):
                        cama.linear(a, w, bias if sp == 1 else None, splits=sp)
                g.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 200 * 1e3
                print(f"{name:5s} M={M} N={N} K={K} BN={bn:>3s} splits={sp}: {us:6.2f} us/launch "
                      f"({2.0 * M * N * K / us / 1e6:6.1f} TF/s)")
            except Exception as ex:  # noqa: BLE001
                print(f"{name} BN={bn} splits={sp}: {type(ex).__name__}: {ex}")
