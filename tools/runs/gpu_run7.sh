#!/bin/bash
mkdir -p gpurun_out
for v in inline outline; do
  if [ $v = outline ]; then export UPB_LIB=$PWD/drl_urban_planning_b200/libupb200_outline.so; else unset UPB_LIB; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
  for spl in 1 10; do
    timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu --skip-e2e --iter-states 0 --steps-per-launch $spl > gpurun_out/r2g_${v}_spl$spl.json 2> gpurun_out/r2g_${v}_spl$spl.err; echo "$v spl $spl rc=$?"; tail -2 gpurun_out/r2g_${v}_spl$spl.err
    python -c "
import json; d=json.load(open('gpurun_out/r2g_${v}_spl$spl.json')); print('$v spl', $spl, round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches'])"
  done
done
unset UPB_LIB
python tools/phase_times.py 2>&1 | tail -24
