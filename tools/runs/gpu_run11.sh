#!/bin/bash
# 8-GPU box: the scaling table (weak 1/2/4/8), strong and buffer modes at 8, NCCL comparison at 8
mkdir -p gpurun_out
run() { # name nproc extra...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 400 python bench.py --gpus 1 --steps 50 --warmup 5 --skip-cpu --iter-states 0 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 50 --warmup 5 --skip-cpu --iter-states 0 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "$name rc=$?"
  python - <<PY
import json
try:
    line=[l for l in open('gpurun_out/$name.json') if l.startswith('{')][-1]
    d=json.loads(line); print('$name', round(d['value']), round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5), (d.get('e2e') or {}).get('value'), d.get('multi_gpu_selfcheck'))
except Exception as e: print('$name', 'ERR', e)
PY
}
run r2s_w1 1
run r2s_w2 2
run r2s_w4 4
run r2s_w8 8
run r2s_s8 8 --mode strong --skip-e2e
run r2s_b8 8 --mode buffer --skip-e2e
run r2s_n8 8 --nccl-exchange --skip-e2e
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2
