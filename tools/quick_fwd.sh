#!/bin/bash
# quick correctness + kernel-time check of the fused forward kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fwd or small or composite or long" 2>&1 | tail -4
timeout 300 python tools/step_breakdown.py 2>&1 | head -4
