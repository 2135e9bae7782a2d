#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
for b in 1 16; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[567]_" -s 56 -c 28 --csv --log-file $OUT/cama_launches_b$b.csv python scripts/cama_profile.py $b > $OUT/cama_prof_b$b.log 2>&1
python - $b <<'PY'
import csv,sys,collections
b=sys.argv[1]
rows=list(csv.reader(open(f'gpurun_out/cama_launches_b{b}.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; data=rows[hdr+1:]
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
print(f"== b={b}: one layer (7 launches)")
for r in data[:7]:
    print(f"  {r[ki][:50]:52s} grid {r[gi]:>16s} {float(r[vi].replace(',',''))/1e3:8.1f} us")
print("  total of 28 launches: %.1f us"%(sum(float(r[vi].replace(',','')) for r in data)/1e3))
PY
done
