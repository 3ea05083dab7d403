#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/tests.log
for w in c2 c3 c4 c5; do
  timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
for w in c3 c4; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$w.csv \
  python bench.py --workload $w --steps 1 --warmup 3 > gpurun_out/launches_$w.log 2>&1
done
cat gpurun_out/tests.log
python - <<'PY'
import json
for w in ['c2','c3','c4','c5']:
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json')); r=d['roofline']
        print(w, 'step_ms', round(d['ms_per_step'],3), 'kern_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'fwdbwd_ms', round(d['fwd_bwd']['ms_per_step'],2))
    except Exception as e:
        print(w, 'ERR', e)
PY
