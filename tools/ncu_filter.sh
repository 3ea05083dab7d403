#!/bin/bash
# correctness of the filter-side kernels, then their cold per-launch durations next to the cuFFT + pack pair they replace
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "kf_from or dk_from or bwd_8192 or bwd_small" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"kf_from_filter|dk_from_dkf|vector_fft|kf_pack|dkf_unpack" -c 60 --csv --log-file gpurun_out/filter_launches.csv python tools/step_breakdown.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/filter_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hdr+2:]:
    if len(r)>vi: agg.setdefault((r[ki][:48], r[mi]),[]).append(float(r[vi].replace(',','')))
for (k,m),v in agg.items(): print(f'{k:50s} {m:55s} n={len(v):3d} mean={sum(v)/len(v):12.1f}')
PY
