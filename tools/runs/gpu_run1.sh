#!/bin/bash
# round-2 GPU check: parity tests, bench line, sanitizer passes on the small fixture path
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json | cut -c1-600
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2a_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2a_racecheck.log
