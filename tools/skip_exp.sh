#!/bin/bash
for s in 0 1 2 3 4 7; do BFFC_SKIP=$s timeout 120 python tools/skip_exp.py 2>&1 | tail -1; done
