#!/usr/bin/env bash
# One N-GPU bench line as the driver launches it (extras limited to the 10 M-row configs)
# usage: gpurun --gpus 8 --timeout 600 -- 'bash scripts/r2_scale8.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_n$N.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 200 --warmup 10 --extras c3q1,c3q4096,c2 > $OUT/scale_n$N.json 2> $OUT/scale_n$N.err
echo "bench n=$N rc=$?"
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/scale_n{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"n={n}: c1 value={d['value']:.0f} q/s ms={d['ms_per_step']:.4f} p50={d['p50_latency_ms']:.4f} e2e={d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.4f} ms) frac={d['roofline']['frac']:.3f} kernel_ms={d['roofline']['kernel_ms']:.4f} parity={d['parity']} exchange={d['exchange']}")
    for k,v in d.get('scale_10m',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac'), v.get('parity'), v.get('error'))
    print('c2', d['c2'] and (d['c2'].get('value'), d['c2'].get('ms_per_step'), d['c2'].get('roofline',{}).get('frac'), d['c2'].get('parity')))
except Exception as e:
    print("ERR", e, open(f"gpurun_out/scale_n{n}.err").read()[-3000:])
PY
tail -5 $OUT/scale_n$N.err
