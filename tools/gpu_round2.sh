#!/bin/bash
# One GPU visit (round 2): parity tests (incl. full BASELINE shapes + error table), the default bench line (all configs),
# the reference arm, MMA microbenchmark for the flop-lean kernel, same-box reference CUDA kernels, ncu captures.
# the tree may be mid-edit when the snapshot is taken: always run the prebuilt library
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_ref.err
timeout 60 ./tools/mbu4 > gpurun_out/mbu4.log 2>&1; cat gpurun_out/mbu4.log
timeout 900 python baseline/run_ref.py > gpurun_out/run_ref.log 2>&1; tail -12 gpurun_out/run_ref.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv \
  python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/launches_c2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:dkf3_kernel -s 1 -c 1 -f -o gpurun_out/prof_dkf3 \
  python tools/prof_bwd.py > gpurun_out/prof_dkf3.log 2>&1
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print('headline', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e'], d['clocks'])
    for k,v in d['roofline']['configs'].items():
        if 'error' in v: print(k, v); continue
        print(k, 'fwd_ms %.3f kern_ms %.3f frac %.3f fwdbwd_ms %.3f (x%.2f) e2e_ms %.2f peak_mb %s' % (v['fwd']['ms_per_step'], v['kernels']['ms'], v['kernels']['frac'], v['fwd_bwd']['ms_per_step'], v['fwd_bwd']['ratio_to_fwd'], v['e2e']['ms_per_step'], {a: round(b) for a, b in v['peak_mem_mb'].items() if a != 'note'}))
except Exception as e:
    print('bench parse error', e)
PY
