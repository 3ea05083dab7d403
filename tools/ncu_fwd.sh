#!/bin/bash
# one ncu --set full capture (with source) of the fused forward kernel at C2, plus the step breakdown
mkdir -p gpurun_out
timeout 300 python tools/step_breakdown.py > gpurun_out/breakdown.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd3_kernel -s 1 -c 1 -f -o gpurun_out/prof_fwd3 python tools/prof_fwd.py > gpurun_out/prof.log 2>&1
cat gpurun_out/breakdown.log; tail -3 gpurun_out/prof.log
