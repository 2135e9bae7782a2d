#!/usr/bin/env bash
# Round-2 single-GPU check: parity tests (search first, under their own timeout), the rest, smoke, a short bench.
# usage: gpurun --timeout 1500 -- 'bash scripts/r2_check.sh [full]'
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "not 10m" -p no:cacheprovider > $OUT/pytest_search.log 2>&1
echo "pytest search rc=$?" | tee -a $OUT/summary.txt; tail -25 $OUT/pytest_search.log | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_search.py ${1:+} > $OUT/pytest_rest.log 2>&1
echo "pytest rest rc=$?" | tee -a $OUT/summary.txt; tail -25 $OUT/pytest_rest.log | tee -a $OUT/summary.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-extras --no-cpu-baseline > $OUT/bench_c1.json 2> $OUT/bench_c1.err
echo "bench c1 rc=$?" | tee -a $OUT/summary.txt; tail -c 2500 $OUT/bench_c1.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench_c1.err | tee -a $OUT/summary.txt
if [ "${1:-}" == "full" ]; then
  timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "10m" -p no:cacheprovider > $OUT/pytest_10m.log 2>&1
  echo "pytest 10m rc=$?" | tee -a $OUT/summary.txt; tail -8 $OUT/pytest_10m.log | tee -a $OUT/summary.txt
fi
