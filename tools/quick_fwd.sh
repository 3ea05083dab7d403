#!/bin/bash
# quick correctness + timing check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/step_breakdown.py 2>&1 | tail -9
