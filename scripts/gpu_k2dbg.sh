#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  python - "$name" <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "ms=%.3f kernel_ms=%.3f ach=%.1f TF frac=%.3f"%(d["ms_per_step"],r["kernel_ms"],r["achieved"],r["frac"]), d.get("clocks"))
except Exception as e:
    print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
}
run pair_normal A=1
run single_normal MRAG_K2_SINGLE=1
run pair_noepi MRAG_K2_DEBUG=1
run single_noepi MRAG_K2_SINGLE=1 MRAG_K2_DEBUG=1
run pair_noinsert MRAG_K2_DEBUG=2
run single_noinsert MRAG_K2_SINGLE=1 MRAG_K2_DEBUG=2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_batch -s 2 -c 1 -f -o $OUT/prof_k2b \
   python bench.py --workload c2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k2b.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/summary.txt
