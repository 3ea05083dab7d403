#!/bin/bash
# GPU visit: parity tests, fwd4 bring-up (BFFC_INNER=4), bench, same-box reference kernels
touch flash-fft-conv_b200/libbffc.so
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=12 2>&1 | tail -25 > gpurun_out/tests.log; cat gpurun_out/tests.log
BFFC_INNER=4 timeout 300 python tools/bringup_fwd4.py > gpurun_out/bringup_fwd4.log 2>&1; cat gpurun_out/bringup_fwd4.log
BFFC_INNER=4 timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fwd_8192_vs_oracle or composite or long or fwd_against or fp16_vs or full_size or bwd_8192 or bwd_composite or bwd_long" 2>&1 | tail -15 > gpurun_out/tests_fwd4.log; cat gpurun_out/tests_fwd4.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
BFFC_INNER=4 timeout 600 python bench.py > gpurun_out/bench_fwd4.json 2> gpurun_out/bench_fwd4.err; tail -3 gpurun_out/bench_fwd4.err
python - <<'PY'
import json
for f in ('bench_default', 'bench_fwd4'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'headline', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
        for k,v in d['roofline']['configs'].items():
            if 'error' in v: print(k, v); continue
            print(k, 'fwd_ms %.3f kern_ms %.3f frac %.3f fwdbwd_ms %.3f (x%.2f) e2e_ms %.2f peak_mb %s' % (v['fwd']['ms_per_step'], v['kernels']['ms'], v['kernels']['frac'], v['fwd_bwd']['ms_per_step'], v['fwd_bwd']['ratio_to_fwd'], v['e2e']['ms_per_step'], {a: round(b) for a, b in v['peak_mem_mb'].items() if a != 'note'}))
    except Exception as e:
        print(f, 'parse error', e)
PY
