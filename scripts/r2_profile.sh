#!/usr/bin/env bash
# Round-2 evidence run on ONE B200: the full bench line, the ncu launch list of the same command, and
# `ncu --set full` captures of the dominant kernels (K1 with the fused K3 tail, K2 pair, K3 merge).
# usage: gpurun --timeout 900 -- 'bash scripts/r2_profile.sh'
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary_profile.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 420 python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
echo "bench rc=$?" | tee -a $OUT/summary_profile.txt; tail -c 1500 $OUT/bench_n1.json | tee -a $OUT/summary_profile.txt; tail -5 $OUT/bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9]_" -c 120 --csv \
    --log-file $OUT/launches_c1.csv python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/summary_profile.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 4 -c 2 -f -o $OUT/prof_k1 \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
echo "ncu k1 rc=$?" | tee -a $OUT/summary_profile.txt
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:k2_batch|k3_merge" -s 4 -c 2 -f -o $OUT/prof_k2 \
    python bench.py --workload c2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_k2.log 2>&1
echo "ncu k2 rc=$?" | tee -a $OUT/summary_profile.txt
ls -la $OUT | tee -a $OUT/summary_profile.txt
