#!/bin/bash
for v in NO_XEARLY NO_STATS_RED; do
echo "== variant $v"
UPB_LIB=$PWD/variants/libupb200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gradient_and_steps" 2>&1 | tail -3
done
echo "== full, sanitizer memcheck on the failing case"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gradient_and_steps and hlg" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed|at 0x|sgnn_kernel" | head -20
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gradient_and_steps and hlg" 2>&1 | grep -E "RACECHECK SUMMARY|hazard|passed|failed|sgnn_kernel" | head -20
