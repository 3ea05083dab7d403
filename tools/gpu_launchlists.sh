#!/bin/bash
touch flash-fft-conv_b200/libbffc.so
mkdir -p gpurun_out
for w in c3 c4 c5 r1k; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/launches_$w.log 2>&1
done
python - <<'PY'
import csv, collections
for w in ['c3','c4','c5','r1k']:
    try:
        rows=[r for r in csv.reader(open(f'gpurun_out/launches_{w}.csv')) if len(r)>10]
        h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
        per=collections.OrderedDict()
        for r in rows[1:]:
            per.setdefault(r[ii],{'k':r[ki]})[r[mi]]=float(r[vi].replace(',',''))
        print('==',w)
        for i,(id_,d) in enumerate(per.items()):
            if i<24: print(f"{d['k'][:58]:58s} {d.get('gpu__time_duration.sum',0)/1000:9.1f} us  rd {d.get('dram__bytes_read.sum',0)/1e6:8.1f} MB wr {d.get('dram__bytes_write.sum',0)/1e6:8.1f} MB")
    except Exception as e:
        print(w,'ERR',e)
PY
