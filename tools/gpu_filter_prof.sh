#!/bin/bash
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "kf_from_filter or dk_from_dkf or filter_fft" --maxfail=8 2>&1 | tail -5
timeout 600 python tools/filter_bench.py 2>&1 | tail -40 | tee gpurun_out/filter_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"filter_|dk_rows|dk_cols|kf_from" -c 5 -o gpurun_out/r2_filter python tools/filter_bench.py --once c4 c2 > gpurun_out/ncu_filter.log 2>&1; tail -3 gpurun_out/ncu_filter.log
