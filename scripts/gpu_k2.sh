#!/usr/bin/env bash
# K2 iteration loop: tensor tests under a hard timeout, then c2 bench for both kernel forms.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "tensor or 1m or golden" -p no:cacheprovider > $OUT/pytest_tensor.log 2>&1
echo "pytest tensor rc=$?" | tee $OUT/summary.txt; tail -8 $OUT/pytest_tensor.log | tee -a $OUT/summary.txt
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $OUT/bench_c2_pair.json 2> $OUT/bench_c2_pair.err
echo "bench pair rc=$?" | tee -a $OUT/summary.txt
MRAG_K2_SINGLE=1 timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $OUT/bench_c2_single.json 2> $OUT/bench_c2_single.err
echo "bench single rc=$?" | tee -a $OUT/summary.txt
timeout 300 python bench.py --workload c1 --steps 300 --warmup 20 --no-extras --no-cpu-baseline > $OUT/bench_c1.json 2> $OUT/bench_c1.err
echo "bench c1 rc=$?" | tee -a $OUT/summary.txt
python - <<'PY' | tee -a gpurun_out/summary.txt
import json
for f in ("bench_c2_pair","bench_c2_single","bench_c1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r=d["roofline"]; print(f, "value=%.0f q/s ms=%.3f p50=%.3f kernel_ms=%.3f frac=%.3f ach=%.1f share=%.3f e2e=%.0f"%(d["value"],d["ms_per_step"],d["p50_latency_ms"],r["kernel_ms"],r["frac"],r["achieved"],r["kernel_share_of_step"],d["e2e"]["value"]), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
