#!/bin/bash
# last GPU seconds of the round: default path sanity (mbarrier wait change), then the warp-specialised dk_f kernel
timeout 60 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fwd_8192_vs or host_matches or composite" 2>&1 | tail -2
echo "== BFFC_DKF3=1"
BFFC_DKF3=1 timeout 75 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "bwd or ragged or fp16_golden" 2>&1 | tail -3
BFFC_DKF3=1 timeout 40 python tools/step_breakdown.py 2>&1 | grep -E "backward"
