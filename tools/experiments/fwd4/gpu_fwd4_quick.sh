#!/bin/bash
touch flash-fft-conv_b200/libbffc.so
mkdir -p gpurun_out
echo "== one box per plane"; BFFC_INNER=4 timeout 300 python tools/bringup_fwd4.py 2>&1 | tail -16
echo "== eight boxes per plane"; BFFC_INNER=4 BFFC_FWD4_BOXES=8 timeout 300 python tools/bringup_fwd4.py 2>&1 | tail -9
BFFC_INNER=4 timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fwd_8192_vs_oracle or composite or long or fwd_against or fp16_vs or full_size or bwd_8192 or bwd_composite or bwd_long" --deselect "tests/test_parity_gpu.py::test_fwd_against_reference_golden[n8192_bf16_gated]" 2>&1 | tail -5
