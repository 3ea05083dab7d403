#!/bin/bash
# A/B: filter-side FFT kernels (default) vs cuFFT + pack / unpack (BFFC_FILTER_FFT=0), whole bench step at C2
for rep in 1 2; do for v in 1 0; do
  BFFC_FILTER_FFT=$v timeout 300 python bench.py --workload c2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('BFFC_FILTER_FFT=$v step_ms %.4f kern_ms %.4f fwdbwd_ms %.3f sm_mhz %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fwd_bwd']['ms_per_step'], d['clocks']['sm_mhz']))"
done; done
