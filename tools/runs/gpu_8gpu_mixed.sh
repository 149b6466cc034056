#!/bin/bash
# 8-GPU box, final kernel: BASELINE configs[4] (hlg_concept + dhm_concept mixed batch) and configs[3] (8192-graph buffer)
mkdir -p gpurun_out
run() { # name nproc extra...
  name=$1; n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 50 --warmup 5 --skip-cpu --skip-e2e --iter-states 0 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    line=[l for l in open('gpurun_out/$name.json') if l.startswith('{')][-1]
    d=json.loads(line); print('$name', round(d['value']), round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5), d['roofline']['frac'], d.get('multi_gpu_selfcheck'))
except Exception as e: print('$name', 'ERR', e)
PY
}
run r2f_mixed8 8 --mixed hlg_concept,dhm_concept
run r2f_buffer8 8 --mode buffer
