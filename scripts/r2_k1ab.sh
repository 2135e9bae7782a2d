#!/usr/bin/env bash
# A/B of the single-query search: fused tail (one launch) vs K1 + K3, with the phase timeline
set -u
mkdir -p gpurun_out; OUT=gpurun_out
for f in 1 0; do
  MRAG_K1_FUSE=$f MRAG_K3_STAMPS=1 timeout 300 python bench.py --steps 200 --warmup 10 --no-extras --no-cpu-baseline > $OUT/ab_fuse$f.json 2> $OUT/ab_fuse$f.err
  echo "fuse=$f rc=$?"; python - $f <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/ab_fuse{sys.argv[1]}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","p50_latency_ms")}, "kernel_ms", d["roofline"]["kernel_ms"], "total", d.get("extra",{}), "e2e", d["e2e"]["ms_per_step"], d["gpu_launches"])
PY
  tail -4 $OUT/ab_fuse$f.err
done
for w in c1s; do
  for f in 1 0; do
  MRAG_K1_FUSE=$f MRAG_K3_STAMPS=1 timeout 300 python bench.py --workload $w --steps 200 --warmup 10 --no-extras --no-cpu-baseline > $OUT/ab_${w}_fuse$f.json 2> $OUT/ab_${w}_fuse$f.err
  echo "$w fuse=$f rc=$?"; python - $w $f <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/ab_{sys.argv[1]}_fuse{sys.argv[2]}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","p50_latency_ms")}, "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
PY
  tail -3 $OUT/ab_${w}_fuse$f.err
  done
done
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "not 10m" -p no:cacheprovider 2>&1 | tail -5
