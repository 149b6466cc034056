#!/bin/bash
mkdir -p gpurun_out
cd tools/tc_probe && timeout 120 ./epq_tc 310 200 > ../../gpurun_out/r2e_tc_probe.txt 2>&1; echo "probe rc=$?"; cat ../../gpurun_out/r2e_tc_probe.txt
timeout 120 ./epq_tc 460 200 >> ../../gpurun_out/r2e_tc_probe.txt 2>&1; tail -3 ../../gpurun_out/r2e_tc_probe.txt
