#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_rag.py -m gpu -x -q -k "not 10m" -p no:cacheprovider 2>&1 | tail -5
for w in c2 c0; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $OUT/k3s_$w.json 2> $OUT/k3s_$w.err
  echo "$w rc=$?"; python - $w <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/k3s_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, "kernel_ms", d["roofline"]["kernel_ms"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["parity"]["mismatches"], d["parity"]["near_ties"], d["gpu_launches"])
PY
  tail -2 $OUT/k3s_$w.err
done
