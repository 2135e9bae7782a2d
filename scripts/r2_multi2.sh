#!/usr/bin/env bash
# usage: gpurun --gpus N --timeout 1200 -- 'bash scripts/r2_multi2.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider > $OUT/pytest_multi_n$N.log 2>&1
echo "pytest multi rc=$?"; tail -12 $OUT/pytest_multi_n$N.log
bash scripts/r2_scale8.sh $N
