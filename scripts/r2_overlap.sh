#!/usr/bin/env bash
# A/B of the programmatic-dependent-launch overlap of consecutive single-query searches
set -u
mkdir -p gpurun_out; OUT=gpurun_out
for w in c1s c1; do
  for o in 1 0; do
    MRAG_K1_OVERLAP=$o timeout 300 python bench.py --workload $w --steps 300 --warmup 20 --no-extras --no-cpu-baseline > $OUT/ov_${w}_$o.json 2> $OUT/ov_${w}_$o.err
    echo "$w overlap=$o rc=$?"; python - $w $o <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/ov_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, "kernel_ms", d["roofline"]["kernel_ms"], "grid", d["roofline"]["plan"]["grid"], "e2e", d["e2e"]["ms_per_step"], d["parity"]["mismatches"], d["parity"]["near_ties"], d["gpu_launches"])
PY
    tail -2 $OUT/ov_${w}_$o.err
  done
done
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_rag.py -m gpu -x -q -k "not 10m" -p no:cacheprovider 2>&1 | tail -5
