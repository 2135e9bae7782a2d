#!/usr/bin/env bash
# One gpurun call: smoke, GPU parity tests (tensor-core tests isolated under their own timeout so
# a hung kernel cannot take the rest down), a short bench and the ncu launch list.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke (stream only pieces first)" | tee $OUT/summary.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "stream or context or merge or normalises or tiny" \
    -p no:cacheprovider > $OUT/pytest_stream.log 2>&1
echo "pytest stream/context rc=$?" | tee -a $OUT/summary.txt; tail -5 $OUT/pytest_stream.log | tee -a $OUT/summary.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "tensor" -p no:cacheprovider > $OUT/pytest_tensor.log 2>&1
echo "pytest tensor rc=$?" | tee -a $OUT/summary.txt; tail -5 $OUT/pytest_tensor.log | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_all.log 2>&1
echo "pytest all rc=$?" | tee -a $OUT/summary.txt; tail -15 $OUT/pytest_all.log | tee -a $OUT/summary.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
if [ "${1:-}" != "quick" ]; then
  timeout 900 python bench.py --steps 200 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err
  echo "bench rc=$?" | tee -a $OUT/summary.txt; tail -c 3000 $OUT/bench.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9]_" -c 80 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
  echo "ncu launches rc=$?" | tee -a $OUT/summary.txt
fi
if [ "${1:-}" == "prof" ] || [ "${2:-}" == "prof" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 4 -c 2 -f -o $OUT/prof_k1 \
      python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 4 -c 1 -f -o $OUT/prof_k1_f32 \
      python bench.py --workload c1f --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k1f.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:k4_gather -c 2 -f -o $OUT/prof_k4 \
      python bench.py --workload c4 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k4.log 2>&1
  echo "ncu k1 rc=$?" | tee -a $OUT/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_batch -s 2 -c 1 -f -o $OUT/prof_k2 \
      python bench.py --workload c2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k2.log 2>&1
  echo "ncu k2 rc=$?" | tee -a $OUT/summary.txt
fi
