#!/bin/bash
# A/B of the L2 read-ahead distance of the level-0 outer kernels at C3 (0 = off)
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "composite or long or gated" 2>&1 | tail -2
for la in 0 296 592 1184 2368; do
  BFFC_LOOKAHEAD=$la timeout 300 python bench.py --workload c3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lookahead $la: step_ms %.4f kern_ms %.4f fwdbwd_ms %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fwd_bwd']['ms_per_step']))"
done
