#!/usr/bin/env bash
# CAMA forward A/B: parity tests, then whole-forward times with the new forms on / off
# usage: gpurun --timeout 900 -- 'bash scripts/r2_cama_ab.sh'
set -u
mkdir -p gpurun_out; OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cama.py tests/test_gpu_context.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15
echo "== default"; timeout 200 python scripts/cama_bench.py --b=1,4,8,10,12,16 --no-torch
echo "== MRAG_K6_WIDE_MIN=0 (old attention)"; MRAG_K6_WIDE_MIN=0 timeout 200 python scripts/cama_bench.py --b=8,10,16 --no-torch
echo "== MRAG_K6_WIDE_MIN=64"; MRAG_K6_WIDE_MIN=64 timeout 200 python scripts/cama_bench.py --b=4,8 --no-torch
echo "== MRAG_K5_PAIR_MIN=148 (old pair threshold)"; MRAG_K5_PAIR_MIN=148 timeout 200 python scripts/cama_bench.py --b=12,16 --no-torch
echo "== MRAG_K5_PAIR_MIN=36"; MRAG_K5_PAIR_MIN=36 timeout 200 python scripts/cama_bench.py --b=10,12 --no-torch
