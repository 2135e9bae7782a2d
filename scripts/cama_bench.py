"""CAMA transformer forward: libmrag kernels vs torch (eager bf16, and torch under a CUDA graph)."""
import statistics
import sys
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import motionrag_b200 as m  # noqa: E402
from motionrag_b200.context import block_causal_mask  # noqa: E402

torch.manual_seed(0)
layer = nn.TransformerEncoderLayer(1024, 16, 4096, 0.0, "gelu", batch_first=True, norm_first=False, bias=True)
enc = nn.TransformerEncoder(layer, 4, enable_nested_tensor=False).eval().cuda().bfloat16()
mask = block_causal_mask(10, 25, "cuda")
cama = m.CamaTransformer(enc, groups=10, group_tokens=25, max_batch=16, device=0)


def timed(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


BATCHES = (1, 2, 16)
for a in sys.argv[1:]:
    if a.startswith("--b="):
        BATCHES = tuple(int(v) for v in a[4:].split(","))
for b in BATCHES:
    x = torch.randn(b, 250, 1024, device="cuda").bfloat16()
    cama.input_view(b).copy_(x)
    with torch.no_grad():
        t_eager = float("nan") if "--no-torch" in sys.argv else timed(lambda: enc(x, mask))
        t_graph = float("nan")
        if "--torch-graph" in sys.argv:   # torch's MHA fast path is not always capturable
            try:
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for _ in range(3):
                        enc(x, mask)
                torch.cuda.current_stream().wait_stream(s)
                with torch.cuda.graph(g):
                    y = enc(x, mask)
                t_graph = timed(g.replay)
            except Exception as e:  # noqa: BLE001
                print("torch CUDA-graph capture failed:", type(e).__name__)
    t_ours = timed(lambda: cama.forward(b=b))
    t_ours_nograph = timed(lambda: cama.forward(b=b, use_graph=False))
    flops = b * 4 * (2 * 250 * 1024 * (3072 + 1024 + 4096 + 4096))
    print(f"b={b:2d}: libmrag graph {t_ours*1e3:7.1f} us ({flops/t_ours/1e9:6.1f} TF/s) | libmrag direct launches "
          f"{t_ours_nograph*1e3:7.1f} us | torch eager bf16 {t_eager*1e3:7.1f} us | torch CUDA-graph {t_graph*1e3:7.1f} us")
