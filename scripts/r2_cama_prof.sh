#!/usr/bin/env bash
# CAMA forward at b=16: per-launch durations of one forward, ncu --set full of the attention kernel and the FFN1 GEMM
# usage: gpurun --timeout 600 -- 'bash scripts/r2_cama_prof.sh'
set -u
mkdir -p gpurun_out; OUT=gpurun_out
b=${1:-16}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[567]_" -s 56 -c 28 --csv --log-file $OUT/cama_launches_b$b.csv python scripts/cama_profile.py $b > $OUT/cama_prof_b$b.log 2>&1
python - $b <<'PY'
import csv,sys
b=sys.argv[1]
rows=list(csv.reader(open(f'gpurun_out/cama_launches_b{b}.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; data=rows[hdr+1:]
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
print(f"== b={b}: 28 launches")
for r in data:
    print(f"  {r[ki][:50]:52s} grid {r[gi]:>16s} {float(r[vi].replace(',',''))/1e3:8.1f} us")
print("  total of 28 launches: %.1f us"%(sum(float(r[vi].replace(',','')) for r in data)/1e3))
PY
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k6_attention" -s 8 -c 1 -f -o $OUT/prof_k6_b$b python scripts/cama_profile.py $b > $OUT/ncu_k6.log 2>&1
echo "ncu k6 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k5_linear_pair" -s 17 -c 1 -f -o $OUT/prof_k5pair_ffn1_b$b python scripts/cama_profile.py $b > $OUT/ncu_k5pair.log 2>&1
echo "ncu k5 pair rc=$?"
