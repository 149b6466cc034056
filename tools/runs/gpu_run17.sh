#!/bin/bash
# evidence refresh at the final commit: launch list, full ncu capture of the step kernel, headline bench line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/r2g_ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sgnn -s 12 -c 2 -o gpurun_out/prof_r2g python bench.py --steps 6 --warmup 3 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/r2g_ncu_full.log 2>&1; echo "ncu full rc=$?"
python bench.py --steps 50 --warmup 5 > gpurun_out/r2g_bench_hlg.json 2> gpurun_out/r2g_bench_hlg.err; echo "hlg rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_hlg.json')); print('hlg', round(d['value']), round(d['ms_per_step'],5), round(d['roofline']['frac'],4), (d.get('e2e') or {}).get('value'), (d.get('e2e_iteration') or {}).get('seconds'))"
