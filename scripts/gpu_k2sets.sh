#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 240 python -m pytest tests -m gpu -x -q -k "tensor or golden" -p no:cacheprovider > $OUT/pytest_tensor.log 2>&1
echo "pytest tensor sets=1 rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/pytest_tensor.log | tee -a $OUT/summary.txt
MRAG_K2_SETS=2 timeout 240 python -m pytest tests -m gpu -x -q -k "tensor or golden" -p no:cacheprovider > $OUT/pytest_tensor2.log 2>&1
echo "pytest tensor sets=2 rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/pytest_tensor2.log | tee -a $OUT/summary.txt
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload ${WL:-c2} --steps ${STEPS:-20} --warmup 5 --no-extras --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  python - "$name" <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "ms=%.3f p50=%.3f kernel_ms=%.4f ach=%.1f frac=%.3f"%(d["ms_per_step"],d["p50_latency_ms"],r["kernel_ms"],r["achieved"],r["frac"]), d.get("clocks"))
except Exception as e:
    print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
}
run pair_s1 MRAG_K2_SETS=1
run pair_s2 MRAG_K2_SETS=2
run single_s1 MRAG_K2_SINGLE=1 MRAG_K2_SETS=1
run single_s2 MRAG_K2_SINGLE=1 MRAG_K2_SETS=2
run pair_s1_noins MRAG_K2_SETS=1 MRAG_K2_DEBUG=2
run pair_s2_noins MRAG_K2_SETS=2 MRAG_K2_DEBUG=2
WL=c1; STEPS=300
for v in 0 1 2 3 4 5; do run k1_f32_v$v MRAG_K1_VARIANT=$v; done
