#!/usr/bin/env bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary_multi.txt
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_multi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_rag.py tests/test_gpu_search.py -m gpu -x -q -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "pytest multi rc=$?" | tee -a $OUT/summary_multi.txt; tail -5 $OUT/pytest_multi.log | tee -a $OUT/summary_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --exchange nccl --no-extras > $OUT/bench_n${N}_nccl.json 2> $OUT/bench_n${N}_nccl.err
python - $N <<'PY' | tee -a gpurun_out/summary_multi.txt
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_n{n}_nccl.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"n={n} NCCL exchange: c1 value={d['value']:.0f} q/s ms={d['ms_per_step']:.3f} p50={d['p50_latency_ms']:.3f} e2e={d['e2e']['value']:.0f}")
except Exception as e:
    print("ERR", e, open(f"gpurun_out/bench_n{n}_nccl.err").read()[-1500:])
PY
for n in 1 $N; do
  if [ "$n" == "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  fi
  echo "bench n=$n rc=$?" | tee -a $OUT/summary_multi.txt
  python - $n <<'PY' | tee -a gpurun_out/summary_multi.txt
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"n={n} c1 value={d['value']:.0f} q/s ms={d['ms_per_step']:.3f} p50={d['p50_latency_ms']:.3f} e2e={d['e2e']['value']:.0f} frac={d['roofline']['frac']:.3f}")
    for k,v in d['extra'].items():
        print("   ", k, {kk:(round(v[kk],3) if isinstance(v.get(kk),float) else v.get(kk)) for kk in ('value','ms_per_step','p50_ms','error')}, 'frac', v.get('roofline',{}).get('frac'))
except Exception as e:
    print("ERR", e, open(f"gpurun_out/bench_n{n}.err").read()[-1500:])
PY
done
