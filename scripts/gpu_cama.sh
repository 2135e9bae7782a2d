#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 300 python -m pytest tests/test_gpu_cama.py -m gpu -x -q -p no:cacheprovider > $OUT/pytest_cama.log 2>&1
echo "pytest cama rc=$?" | tee -a $OUT/summary.txt; tail -30 $OUT/pytest_cama.log | tee -a $OUT/summary.txt
timeout 300 python scripts/cama_bench.py > $OUT/cama_bench.txt 2>&1; cat $OUT/cama_bench.txt | tail -20 | tee -a $OUT/summary.txt
bash scripts/gpu_cama_prof.sh 2>&1 | tee -a $OUT/summary.txt
