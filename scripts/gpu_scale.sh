#!/usr/bin/env bash
# usage: gpurun --gpus 8 --timeout 1800 -- 'bash scripts/gpu_scale.sh'
set -u
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary_scale.txt
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_scale.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_search.py -m gpu -x -q -p no:cacheprovider > $OUT/pytest_scale.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary_scale.txt; tail -3 $OUT/pytest_scale.log | tee -a $OUT/summary_scale.txt
for n in 1 2 4 8; do
  EXTRA="--no-extras"; [ "$n" == "8" ] && EXTRA=""; [ "$n" == "1" ] && EXTRA=""
  if [ "$n" == "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline $EXTRA > $OUT/scale_n1.json 2> $OUT/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
      bench.py --gpus $n --steps 300 --warmup 20 --no-cpu-baseline $EXTRA > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err
  fi
  echo "bench n=$n rc=$?" | tee -a $OUT/summary_scale.txt
  python - $n <<'PY' | tee -a gpurun_out/summary_scale.txt
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/scale_n{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"n={n} c1 value={d['value']:.0f} q/s ms={d['ms_per_step']:.4f} p50={d['p50_latency_ms']:.4f} e2e={d['e2e']['value']:.0f} frac={d['roofline']['frac']:.3f} kernel_ms={d['roofline']['kernel_ms']:.4f}")
    for k,v in d.get('extra',{}).items():
        print("   ", k, {kk:(round(v[kk],4) if isinstance(v.get(kk),float) else v.get(kk)) for kk in ('value','ms_per_step','p50_ms','error')}, 'frac', v.get('roofline',{}).get('frac'))
except Exception as e:
    print("ERR", e, open(f"gpurun_out/scale_n{n}.err").read()[-1500:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 \
  bench.py --gpus 8 --steps 300 --warmup 20 --no-cpu-baseline --no-extras --exchange nccl > $OUT/scale_n8_nccl.json 2> $OUT/scale_n8_nccl.err
python - <<'PY' | tee -a gpurun_out/summary_scale.txt
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_n8_nccl.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"n=8 NCCL exchange c1 value={d['value']:.0f} q/s ms={d['ms_per_step']:.4f} p50={d['p50_latency_ms']:.4f}")
except Exception as e:
    print("ERR", e)
PY
