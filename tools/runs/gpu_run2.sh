#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_update.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err
python bench.py --steps 50 --warmup 5 --community dhm --iter-states 0 --skip-cpu > gpurun_out/r2b_bench_dhm.json 2> gpurun_out/r2b_bench_dhm.err; echo "dhm rc=$?"
python bench.py --steps 50 --warmup 5 --mixed hlg_concept,dhm_concept --iter-states 0 --skip-cpu --padded-gpu > gpurun_out/r2b_bench_mixed.json 2> gpurun_out/r2b_bench_mixed.err; echo "mixed rc=$?"; tail -3 gpurun_out/r2b_bench_mixed.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
python tools/phase_times.py > gpurun_out/r2b_phase.txt 2>&1
python tools/balance_check.py > gpurun_out/r2b_balance.txt 2>&1
python tools/tail_times.py > gpurun_out/r2b_tail.txt 2>&1
cat gpurun_out/r2b_phase.txt gpurun_out/r2b_balance.txt gpurun_out/r2b_tail.txt
