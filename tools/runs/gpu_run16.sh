#!/bin/bash
# evidence run (final kernel of round 2): launch list, full ncu capture of the step kernel, sanitizer passes, bench lines of the BASELINE configs
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/r2f_ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sgnn -s 12 -c 2 -o gpurun_out/prof_r2f python bench.py --steps 6 --warmup 3 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/r2f_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2f_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2f_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2f_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2f_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2f_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r2f_synccheck.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2f_bench_hlg.json 2> gpurun_out/r2f_bench_hlg.err; echo "hlg rc=$?"
python bench.py --steps 50 --warmup 5 --community dhm --iter-states 0 > gpurun_out/r2f_bench_dhm.json 2> gpurun_out/r2f_bench_dhm.err; echo "dhm rc=$?"
python bench.py --steps 50 --warmup 5 --community dhm --iter-states 0 --skip-cpu --skip-e2e --tiles bf16 > gpurun_out/r2f_bench_dhm_bf16.json 2> gpurun_out/r2f_bench_dhm_bf16.err; echo "dhm bf16 rc=$?"
python bench.py --steps 50 --warmup 5 --mixed hlg_concept,dhm_concept --iter-states 0 --skip-cpu --padded-gpu > gpurun_out/r2f_bench_mixed.json 2> gpurun_out/r2f_bench_mixed.err; echo "mixed rc=$?"
python bench.py --steps 50 --warmup 5 --mode buffer --iter-states 0 --skip-cpu --skip-e2e > gpurun_out/r2f_bench_buffer.json 2> gpurun_out/r2f_bench_buffer.err; echo "buffer rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "ref rc=$?"
for f in hlg dhm dhm_bf16 mixed buffer; do python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],5), round(d['roofline']['frac'],4), (d.get('e2e') or {}).get('value'))"; done
