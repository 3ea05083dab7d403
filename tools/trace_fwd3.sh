#!/bin/bash
BFFC_TRACE=/tmp/trace.bin ITERS=2 timeout 120 python tools/prof_fwd.py > /dev/null 2>&1
python tools/trace_fwd3.py /tmp/trace.bin
