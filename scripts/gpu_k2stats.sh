#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 400 python -m pytest tests -m gpu -x -q -k "tensor or golden or 1m or filter or width or margin" -p no:cacheprovider > $OUT/pytest_tensor.log 2>&1
echo "pytest tensor rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/pytest_tensor.log | tee -a $OUT/summary.txt
run() { local name=$1; shift
  env "$@" MRAG_K2_STATS=1 timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  echo "== $name" | tee -a $OUT/summary.txt; grep "k2 stats" $OUT/$name.err | tail -1 | tee -a $OUT/summary.txt
}
run pair_s1_kc16 MRAG_K2_SETS=1
run single_s1_kc16 MRAG_K2_SINGLE=1 MRAG_K2_SETS=1
run pair_s1_noins MRAG_K2_SETS=1 MRAG_K2_DEBUG=2
bench() { local name=$1; shift
  env "$@" timeout 300 python bench.py --workload ${WL:-c2} --steps ${STEPS:-50} --warmup 5 --no-extras --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  python - "$name" <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "ms=%.3f p50=%.3f kernel_ms=%.4f ach=%.1f frac=%.3f e2e=%.0f"%(d["ms_per_step"],d["p50_latency_ms"],r["kernel_ms"],r["achieved"],r["frac"],d["e2e"]["value"]), d.get("clocks"), r.get("plan"))
except Exception as e:
    print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
}
bench b_pair MRAG_K2_SETS=1
bench b_pair_again MRAG_K2_SETS=1
bench b_single MRAG_K2_SINGLE=1
WL=c3q4096 STEPS=6 bench b_pair_10m A=1
WL=c4 STEPS=100 bench b_c4 A=1
