#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for spl in 1 10 50; do
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu --skip-e2e --iter-states 0 --steps-per-launch $spl > gpurun_out/r2f_bench_spl$spl.json 2> gpurun_out/r2f_bench_spl$spl.err; echo "spl $spl rc=$?"; tail -2 gpurun_out/r2f_bench_spl$spl.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_spl$spl.json')); print('spl', $spl, round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches'])"
done
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "full rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print('full', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['e2e_iteration'])"
