#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/tests.log
for w in c2 c3; do
  timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
cat gpurun_out/tests.log
python - <<'PY'
import json
for w in ['c2','c3']:
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json')); r=d['roofline']
        print(w, 'step_ms', round(d['ms_per_step'],3), 'kern_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'fwdbwd_ms', round(d['fwd_bwd']['ms_per_step'],2))
    except Exception as e:
        print(w, 'ERR', e)
PY
