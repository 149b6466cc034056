#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2i_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_iteration'], d['cpu_baseline'])"
