#!/bin/bash
# ncu --set full of the two outer-stage kernels of the C3 path (N=32768 gated, L=16384)
mkdir -p gpurun_out
N=32768 B=8 H=1024 L=16384 GATED=1 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^(fwd_kernel|inv_kernel)$" -s 2 -c 2 -f -o gpurun_out/prof_c3_outer python tools/prof_fwd.py > gpurun_out/prof_c3.log 2>&1
tail -3 gpurun_out/prof_c3.log
