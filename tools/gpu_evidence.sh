#!/bin/bash
# Evidence visit: full GPU test suite, the driver's bench command, its ncu launch list, ncu --set full captures of the
# dominant kernels, per-config launch lists, the reference's own kernels on the same box.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=12 2>&1 | tail -8 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('headline', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['roofline']['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, 'fwd_ms %.3f eval_ms %.3f kern_ms %.3f frac %.3f fwdbwd_ms %.3f (x%.2f) e2e_ms %.2f' % (v['fwd']['ms_per_step'], v['fwd_eval_cached_kf']['ms_per_step'], v['kernels']['ms'], v['kernels']['frac'], v['fwd_bwd']['ms_per_step'], v['fwd_bwd']['ratio_to_fwd'], v['e2e']['ms_per_step']))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
# ncu launch list of the bench command itself (workload restricted to the headline config to bound the list)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --workload c2 > gpurun_out/bench_under_ncu.log 2>&1
grep -c "fwd3_kernel" gpurun_out/r2_launches_bench.csv
# full captures: C2 forward + backward (fwd3 ungated x2, dkf3), r8k (fwd3 gated), r1k (fwd3 gated, small)
W=c2 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd3_kernel|dkf3_kernel" -c 3 -o gpurun_out/r2_c2_fwd_bwd python tools/prof_step.py > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
W=r8k ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd3_kernel" -c 3 -o gpurun_out/r2_r8k_gated python tools/prof_step.py > gpurun_out/ncu_r8k.log 2>&1; tail -2 gpurun_out/ncu_r8k.log
W=r1k ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd3_kernel" -c 1 -o gpurun_out/r2_r1k_small python tools/prof_step.py > gpurun_out/ncu_r1k.log 2>&1; tail -2 gpurun_out/ncu_r1k.log
bash tools/gpu_launchlists.sh > /dev/null 2>&1; grep "^## " gpurun_out/r2_launches.md
timeout 900 python baseline/run_ref.py > gpurun_out/run_ref.log 2>&1; tail -10 gpurun_out/run_ref.log | cut -c1-260
