#!/bin/bash
# quick kernel check: parity tests of the SGNN path + bench line + phase table (+ the same for variant libraries)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_update.py -m gpu -x -q 2>&1 | tail -2
run() {
timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print(round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
python tools/phase_times.py 2>&1 | grep -E "cycles|total|busy"
}
run
for v in variants/*.so; do [ -f "$v" ] || continue; echo "== $v"; export UPB_LIB=$PWD/$v
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
run; done
unset UPB_LIB
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gradient_and_steps and (hlg or small)" 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|passed|failed" | head -8
