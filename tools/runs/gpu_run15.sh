#!/bin/bash
mkdir -p gpurun_out
for v in base varA varB; do
  if [ $v = base ]; then unset UPB_LIB; else export UPB_LIB=$PWD/drl_urban_planning_b200/libupb200_$v.so; fi
  for rep in 1 2; do
  timeout 300 python bench.py --steps 100 --warmup 10 --skip-cpu --skip-e2e --iter-states 0 > gpurun_out/r2v_$v.json 2> gpurun_out/r2v_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2v_$v.json')); print('$v', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done
