#!/bin/bash
# full parity suite + step breakdown + c3 / c4 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/step_breakdown.py 2>&1 | grep -E "kernel only|module forward"
for w in c3 c4; do
  timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w step_ms %.3f kern_ms %.3f fwdbwd_ms %.2f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fwd_bwd']['ms_per_step']))"
done
