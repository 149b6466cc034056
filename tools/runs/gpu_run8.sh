#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python tools/tail_times.py 2>&1 | tail -5
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
