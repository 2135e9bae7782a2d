#!/usr/bin/env bash
# K1 variants on a 125k-row shard (what each GPU scans for c1 at N=8)
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
for v in 0 1 2 3 4 5; do
  MRAG_K1_VARIANT=$v timeout 200 python bench.py --workload c1s --steps 600 --warmup 20 --no-extras --no-cpu-baseline > $OUT/k1s$v.json 2> $OUT/k1s$v.err
  python - $v <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/k1s{v}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("c1s bf16 variant %s: value=%.0f ms=%.4f kernel_ms=%.4f ach=%.0f frac=%.3f grid=%s"%(v,d["value"],d["ms_per_step"],r["kernel_ms"],r["achieved"],r["frac"],r["plan"]["grid"]))
except Exception as e:
    print(v,"ERR", e, open(f"gpurun_out/k1s{v}.err").read()[-800:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 4 -c 1 -f -o $OUT/prof_k1_small \
    python bench.py --workload c1s --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k1s.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/summary.txt
