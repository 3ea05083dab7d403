#!/bin/bash
# backward parity + timing, pipelined dk_f kernel (default) vs BFFC_DKF3=0
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "bwd or ragged or fp16 or dk_from" 2>&1 | tail -4
for v in 1 0; do echo "BFFC_DKF3=$v"; BFFC_DKF3=$v timeout 300 python tools/step_breakdown.py 2>&1 | grep -E "forward \+ backward"; done
