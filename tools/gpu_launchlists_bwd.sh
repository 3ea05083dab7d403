#!/bin/bash
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
for w in c2 c3 c4; do
  W=$w ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_fb_$w.csv \
    python tools/prof_step.py > gpurun_out/launches_fb_$w.log 2>&1
done
python - <<'PY'
import csv, collections
for w in ['c2','c3','c4']:
    try:
        rows=[r for r in csv.reader(open(f'gpurun_out/launches_fb_{w}.csv')) if len(r)>10]
        h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
        per=collections.OrderedDict()
        for r in rows[1:]:
            per.setdefault(r[ii],{'k':r[ki]})[r[mi]]=float(r[vi].replace(',',''))
        items=list(per.values())
        # second iteration only: find the last occurrence of the first conv kernel name pattern
        names=[d['k'] for d in items]
        half=len(items)//2
        print('==',w,'(launches of the run:',len(items),')')
        tot=0
        for d in items[-(len(items)-12)//2:] if False else items:
            pass
        start=max(i for i,n in enumerate(names) if 'randn' in n or 'distribution' in n or 'copy' in n.lower() and i<len(names)//2) if False else 0
        for d in items[len(items)-((len(items)-10)//2):]:
            t=d.get('gpu__time_duration.sum',0)/1000; tot+=t
            print(f"{d['k'][:62]:62s} {t:9.1f} us  rd {d.get('dram__bytes_read.sum',0)/1e6:8.1f} wr {d.get('dram__bytes_write.sum',0)/1e6:8.1f} MB")
        print('sum', round(tot,1), 'us')
    except Exception as e:
        print(w,'ERR',e)
PY
