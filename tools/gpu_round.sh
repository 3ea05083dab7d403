#!/bin/bash
# One GPU visit: parity tests, bench lines for every BASELINE config, launch list + full ncu capture of the top kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/tests.log
for w in c2 c3 c4 c5; do
  timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv \
  python bench.py --steps 3 --warmup 3 > gpurun_out/launches_c2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fwd3_kernel -s 1 -c 1 -f -o gpurun_out/prof_fwd8k \
  python tools/prof_fwd.py > gpurun_out/prof.log 2>&1
cat gpurun_out/tests.log; for w in c2 c3 c4 c5; do head -c 600 gpurun_out/bench_$w.json; echo; tail -2 gpurun_out/bench_$w.err; done
BFFC_TRACE=/tmp/trace.bin ITERS=2 timeout 120 python tools/prof_fwd.py > /dev/null 2>&1; python tools/trace_fwd3.py /tmp/trace.bin > gpurun_out/trace.txt 2>&1
timeout 300 python tools/step_breakdown.py > gpurun_out/breakdown.log 2>&1; cat gpurun_out/breakdown.log
python - <<'PY'
import json
for w in ['c2','c3','c4','c5']:
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json')); r=d['roofline']; t=d['roofline_tensor']
        print(w, 'step_ms', round(d['ms_per_step'],3), 'kern_ms', round(r['kernel_ms'],3), 'hbm frac', round(r['frac'],3), 'tensor frac', round(t['frac'],3),
              'fwdbwd_ms', round(d['fwd_bwd']['ms_per_step'],2), 'e2e_ms', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'])
    except Exception as e:
        print(w, 'ERR', e)
PY
