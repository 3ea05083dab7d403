"""gpurun_out/launches_fb_<w>.csv (ncu --csv launch lists of tools/prof_step.py) -> markdown tables, one iteration each."""
import collections, csv, re, sys

print('# Launch lists, forward + backward through the module (tools/gpu_launchlists.sh)\n')
print('ncu `gpu__time_duration.sum`, `dram__bytes_{read,write}.sum` per launch, `--clock-control none`; second of two '
      'identical iterations.  Times under ncu are serialised and partly cold-cache: read them as shares of the step.\n')
for w in sys.argv[1:]:
    try:
        rows = [r for r in csv.reader(open(f'gpurun_out/launches_fb_{w}.csv')) if len(r) > 10]
        h = rows[0]
        ki, mi, vi, ii = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
        per = collections.OrderedDict()
        for r in rows[1:]:
            per.setdefault(r[ii], {'k': r[ki]})[r[mi]] = float(r[vi].replace(',', ''))
        items = list(per.values())
        names = [d['k'] for d in items]
        P = next((p for p in range(1, len(names) // 2 + 1) if names[-p:] == names[-2 * p:-p] and
                  any('fwd' in n or 'dkf' in n for n in names[-p:])), len(names))
        it = items[-P:]
        tot = sum(d.get('gpu__time_duration.sum', 0) for d in it) / 1000
        print(f'## {w}: {P} launches, {tot:.1f} us\n')
        print('| kernel | us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|')
        for d in it:
            t = d.get('gpu__time_duration.sum', 0) / 1000
            name = re.sub(r'\(.*', '', d['k']).replace('void ', '').replace('bffc::', '')[:70]
            print(f"| `{name}` | {t:.1f} | {t / tot:.1%} | {d.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {d.get('dram__bytes_write.sum', 0) / 1e6:.1f} |")
        print()
    except Exception as e:
        print(f'## {w}: error {e}\n')
