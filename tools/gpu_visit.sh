#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_full_gpu.py -q -m gpu -k "small or 256 or 512 or 1024 or 2048 or 4096 or kf_from_filter or dk_from_dkf" --maxfail=15 2>&1 | tail -40 > gpurun_out/tests_small.log; cat gpurun_out/tests_small.log
timeout 1800 python -m pytest tests -q -m gpu --maxfail=12 2>&1 | tail -12 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('headline', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
for k,v in d['roofline']['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, 'fwd_ms %.3f eval_ms %.3f kern_ms %.3f frac %.3f fwdbwd_ms %.3f (x%.2f) e2e_ms %.2f peak_mb %s' % (v['fwd']['ms_per_step'], v['fwd_eval_cached_kf']['ms_per_step'], v['kernels']['ms'], v['kernels']['frac'], v['fwd_bwd']['ms_per_step'], v['fwd_bwd']['ratio_to_fwd'], v['e2e']['ms_per_step'], {a: round(b) for a, b in v['peak_mem_mb'].items() if a != 'note'}))
PY
