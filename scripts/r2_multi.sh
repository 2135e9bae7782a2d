#!/usr/bin/env bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary_multi.txt
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_multi.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "pytest multi rc=$?" | tee -a $OUT/summary_multi.txt; tail -30 $OUT/pytest_multi.log | tee -a $OUT/summary_multi.txt
for x in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-extras --exchange $x > $OUT/bench_n${N}_$x.json 2> $OUT/bench_n${N}_$x.err
echo "bench n=$N $x rc=$?" | tee -a $OUT/summary_multi.txt
python - $N $x <<'PY' | tee -a gpurun_out/summary_multi.txt
import json,sys
n,x=sys.argv[1:3]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_n{n}_{x}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"n={n} {x}: c1 value={d['value']:.0f} q/s ms={d['ms_per_step']:.4f} p50={d['p50_latency_ms']:.4f} e2e={d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.4f} ms) frac={d['roofline']['frac']:.3f} kernel_ms={d['roofline']['kernel_ms']:.4f}")
except Exception as e:
    print("ERR", e, open(f"gpurun_out/bench_n{n}_{x}.err").read()[-2500:])
PY
done
