#!/bin/bash
# full parity suite + c3 / c5 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in c3 c5; do
  timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w step_ms %.3f kern_ms %.3f hbm_frac %.3f fwdbwd_ms %.2f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['fwd_bwd']['ms_per_step']))"
done
