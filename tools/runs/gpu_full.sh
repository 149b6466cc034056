#!/bin/bash
# what the driver runs at round end on one GPU: the GPU test tier, smoke(), both bench arms with default flags
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
/usr/bin/time -v python bench.py --impl reference > gpurun_out/full_ref.json 2> gpurun_out/full_ref.err; echo "ref rc=$? $(grep Elapsed gpurun_out/full_ref.err)"
/usr/bin/time -v python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$? $(grep Elapsed gpurun_out/full_bench.err)"
tail -c 600 gpurun_out/full_bench.json; echo; tail -c 300 gpurun_out/full_ref.json
