#!/bin/bash
touch flash-fft-conv_b200/libbffc.so
mkdir -p gpurun_out
BFFC_INNER=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd4_kernel -s 1 -c 1 -f -o gpurun_out/prof_fwd4 \
  python tools/prof_fwd.py > gpurun_out/prof_fwd4.log 2>&1; tail -3 gpurun_out/prof_fwd4.log
