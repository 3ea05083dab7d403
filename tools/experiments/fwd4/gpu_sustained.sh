#!/bin/bash
touch flash-fft-conv_b200/libbffc.so
timeout 120 python tools/sustained_c2.py 2>&1 | tail -8
BFFC_INNER=4 timeout 120 python tools/sustained_c2.py 2>&1 | tail -8
