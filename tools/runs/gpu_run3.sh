#!/bin/bash
# 2-GPU: parity of the multi-GPU path + bench lines (weak, strong), TMA build
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --skip-e2e > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; echo "2gpu rc=$?"; tail -2 gpurun_out/r2c_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --skip-e2e --mode strong > gpurun_out/r2c_bench_2gpu_strong.json 2> gpurun_out/r2c_bench_2gpu_strong.err; echo "2gpu strong rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --skip-e2e --nccl-exchange > gpurun_out/r2c_bench_2gpu_nccl.json 2> gpurun_out/r2c_bench_2gpu_nccl.err; echo "2gpu nccl rc=$?"
python bench.py --steps 50 --warmup 5 --skip-cpu --iter-states 0 > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/r2c_bench_1gpu.err; echo "1gpu rc=$?"
for f in r2c_bench_1gpu r2c_bench_2gpu r2c_bench_2gpu_strong r2c_bench_2gpu_nccl; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d.get('multi_gpu_selfcheck'))"; done
python tools/tail_times.py 2>&1 | tail -5
