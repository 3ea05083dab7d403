#!/bin/bash
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=12 -k "bwd or backward or c2_full or c3_full or ragged or golden or error_table" 2>&1 | tail -8 > gpurun_out/tests_bwd.log; cat gpurun_out/tests_bwd.log
timeout 120 python tools/step_breakdown.py 2>&1 | tail -12
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dkf3_kernel -s 1 -c 1 -f -o gpurun_out/prof_dkf3b \
  python tools/prof_bwd.py > gpurun_out/prof_dkf3b.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('headline', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
for k,v in d['roofline']['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, 'fwd_ms %.3f kern_ms %.3f frac %.3f fwdbwd_ms %.3f (x%.2f) e2e_ms %.2f' % (v['fwd']['ms_per_step'], v['kernels']['ms'], v['kernels']['frac'], v['fwd_bwd']['ms_per_step'], v['fwd_bwd']['ratio_to_fwd'], v['e2e']['ms_per_step']))
PY
