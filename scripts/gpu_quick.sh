#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/pytest_all.log 2>&1
echo "pytest all rc=$?" | tee -a $OUT/summary.txt; tail -5 $OUT/pytest_all.log | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 300 --warmup 20 --no-extras --no-cpu-baseline > $OUT/bench_c1.json 2> $OUT/bench_c1.err
python - <<'PY' | tee -a gpurun_out/summary.txt
import json
try:
    d=json.loads(open("gpurun_out/bench_c1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("c1 value=%.0f ms=%.4f p50=%.4f kernel_ms=%.4f frac=%.3f e2e=%.0f (%.4f ms) launches=%d"%(d["value"],d["ms_per_step"],d["p50_latency_ms"],r["kernel_ms"],r["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["gpu_launches"]))
except Exception as e:
    print("ERR", e, open("gpurun_out/bench_c1.err").read()[-1500:])
PY
