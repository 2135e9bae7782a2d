#!/usr/bin/env bash
# CAMA (row f-1) check: parity tests, then forward / predict / torch baselines and the K4 kernel-only times
# usage: gpurun --timeout 900 -- 'bash scripts/r2_cama.sh'
set -u
mkdir -p gpurun_out; OUT=gpurun_out
timeout 500 python -m pytest tests/test_gpu_cama.py tests/test_gpu_context.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -12
timeout 300 python bench.py --workload c1s --steps 50 --warmup 5 --no-cpu-baseline --extras cama_forward,k4_gather > $OUT/bench_cama.json 2> $OUT/bench_cama.err
echo "bench rc=$?"; tail -3 $OUT/bench_cama.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cama.json").read().strip().splitlines()[-1])["extra"]
print(json.dumps(d, indent=1))
PY
