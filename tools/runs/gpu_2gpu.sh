#!/bin/bash
# 2-GPU box: the multi-GPU parity tests and a weak-scaling bench line with the in-bench self-check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --skip-cpu --iter-states 0 > gpurun_out/r2f_w2.json 2> gpurun_out/r2f_w2.err; echo "w2 rc=$?"
python - <<PY
import json
line=[l for l in open('gpurun_out/r2f_w2.json') if l.startswith('{')][-1]
d=json.loads(line); print('w2', round(d['value']), round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5), (d.get('e2e') or {}).get('value'), d.get('multi_gpu_selfcheck'))
PY
