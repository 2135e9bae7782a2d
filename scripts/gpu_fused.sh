#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 300 python -m pytest tests/test_gpu_cama.py -m gpu -x -q -p no:cacheprovider > $OUT/pytest_fused.log 2>&1
echo "pytest cama (fused default) rc=$?" | tee -a $OUT/summary.txt; tail -5 $OUT/pytest_fused.log | tee -a $OUT/summary.txt
MRAG_CAMA_FUSED=0 timeout 300 python -m pytest tests/test_gpu_cama.py -m gpu -x -q -p no:cacheprovider -k "forward or causality or attach" > $OUT/pytest_chain.log 2>&1
echo "pytest cama (chain) rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/pytest_chain.log | tee -a $OUT/summary.txt
for f in 1 0; do echo "MRAG_CAMA_FUSED=$f" | tee -a $OUT/summary.txt; MRAG_CAMA_FUSED=$f timeout 200 python scripts/cama_bench.py 2>&1 | tail -3 | cut -c1-120 | tee -a $OUT/summary.txt; done
MRAG_CAMA_FUSED_STAMPS=1 timeout 120 python scripts/cama_profile.py 1 2>&1 | grep "fused layer" | tail -2 | tee -a $OUT/summary.txt
