#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
for v in 0 1 2 3 4 5; do
  MRAG_K1_VARIANT=$v timeout 300 python bench.py --workload c1 --steps 400 --warmup 20 --no-extras --no-cpu-baseline > $OUT/k1v$v.json 2> $OUT/k1v$v.err
  python - $v <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/k1v{v}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("bf16 variant %s: value=%.0f ms=%.4f kernel_ms=%.4f ach=%.0f frac=%.3f grid=%s e2e=%.0f"%(v,d["value"],d["ms_per_step"],r["kernel_ms"],r["achieved"],r["frac"],r["plan"]["grid"],d["e2e"]["value"]))
except Exception as e:
    print(v,"ERR", e, open(f"gpurun_out/k1v{v}.err").read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 4 -c 2 -f -o $OUT/prof_k1_bf16 \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
echo "ncu k1 rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/summary.txt
