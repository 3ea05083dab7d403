#!/bin/bash
# GPU visit: filter-side FFT tests first (fast fail), then the whole suite, bench
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "kf_from_filter or dk_from_dkf or filter_fft" --maxfail=8 2>&1 | tail -25 > gpurun_out/tests_filter.log; cat gpurun_out/tests_filter.log
timeout 1800 python -m pytest tests -q -m gpu --maxfail=12 2>&1 | tail -25 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
for f in ('bench_default',):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'headline', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
        for k,v in d['roofline']['configs'].items():
            if 'error' in v: print(k, v); continue
            print(k, 'fwd_ms %.3f eval_ms %.3f kern_ms %.3f frac %.3f fwdbwd_ms %.3f (x%.2f) e2e_ms %.2f peak_mb %s' % (v['fwd']['ms_per_step'], v['fwd_eval_cached_kf']['ms_per_step'], v['kernels']['ms'], v['kernels']['frac'], v['fwd_bwd']['ms_per_step'], v['fwd_bwd']['ratio_to_fwd'], v['e2e']['ms_per_step'], {a: round(b) for a, b in v['peak_mem_mb'].items() if a != 'note'}))
    except Exception as e:
        print(f, 'parse error', e)
PY
