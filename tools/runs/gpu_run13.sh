#!/bin/bash
mkdir -p gpurun_out
for t in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $t --print-limit 10 python __graft_entry__.py --smoke > gpurun_out/r2_final_$t.log 2>&1; echo "$t rc=$?"; grep -E "SUMMARY|smoke ok" gpurun_out/r2_final_$t.log | cut -c1-160
done
