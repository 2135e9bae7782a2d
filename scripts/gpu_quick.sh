#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/pytest_all.log 2>&1
echo "pytest all rc=$?" | tee -a $OUT/summary.txt; tail -5 $OUT/pytest_all.log | tee -a $OUT/summary.txt
timeout 900 python bench.py --steps 300 --warmup 20 ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' | tee -a gpurun_out/summary.txt
import json
try:
    d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("c1 value=%.0f ms=%.4f p50=%.4f kernel_ms=%.4f frac=%.3f e2e=%.0f (%.4f ms) launches=%d cpu=%s"%(d["value"],d["ms_per_step"],d["p50_latency_ms"],r["kernel_ms"],r["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["gpu_launches"], d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],1)))
    for k,v in d.get('extra',{}).items():
        rf=v.get('roofline',{})
        print("   ", k, {kk:(round(v[kk],4) if isinstance(v.get(kk),float) else v.get(kk)) for kk in ('value','ms_per_step','p50_ms','error')}, 'kernel_ms', rf.get('kernel_ms'), 'ach', rf.get('achieved'), 'frac', rf.get('frac'), 'e2e', (v.get('e2e') or {}).get('value'))
except Exception as e:
    print("ERR", e, open("gpurun_out/bench.err").read()[-1500:])
PY
timeout 300 python scripts/e2e_breakdown.py > $OUT/e2e_breakdown.txt 2>&1; cat $OUT/e2e_breakdown.txt | tee -a $OUT/summary.txt
