#!/bin/bash
mkdir -p gpurun_out
echo "== tests with fwd3"; BFFC_FWD3=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for v in 0 1; do
  echo "== BFFC_FWD3=$v"
  BFFC_FWD3=$v timeout 200 python tools/gpu_bringup.py 2>&1 | grep -E "C2|rel-L2 vs|B=16"
done
