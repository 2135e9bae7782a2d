#!/usr/bin/env bash
# Round-2 final single-GPU validation, the driver's own three commands: pytest -m gpu, smoke(), bench.py (defaults) + reference arm
# usage: gpurun --timeout 1500 -- 'bash scripts/r2_final.sh'
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/final_summary.txt
S=$(date +%s)
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=15 > $OUT/final_pytest.log 2>&1
echo "pytest rc=$? ($(( $(date +%s) - S )) s)" | tee -a $OUT/final_summary.txt; tail -30 $OUT/final_pytest.log | tee -a $OUT/final_summary.txt
S=$(date +%s)
timeout 300 python __graft_entry__.py --smoke > $OUT/final_smoke.log 2>&1
echo "smoke rc=$? ($(( $(date +%s) - S )) s)" | tee -a $OUT/final_summary.txt; tail -3 $OUT/final_smoke.log | tee -a $OUT/final_summary.txt
S=$(date +%s)
timeout 900 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err
echo "bench rc=$? ($(( $(date +%s) - S )) s)" | tee -a $OUT/final_summary.txt; tail -3 $OUT/final_bench.err | tee -a $OUT/final_summary.txt
S=$(date +%s)
timeout 600 python bench.py --impl reference > $OUT/final_bench_ref.json 2> $OUT/final_bench_ref.err
echo "bench ref rc=$? ($(( $(date +%s) - S )) s)" | tee -a $OUT/final_summary.txt; cat $OUT/final_bench_ref.json | tee -a $OUT/final_summary.txt
