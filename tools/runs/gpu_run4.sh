#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 --skip-cpu --iter-states 0 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"
python tools/phase_times.py 2>&1 | tail -26
python tools/tail_times.py 2>&1 | tail -3
