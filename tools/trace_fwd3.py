"""Phase timeline of CTA 0 of fwd3_kernel (bring-up; run with BFFC_TRACE=<file> set, e.g. via tools/trace_fwd3.sh).
Stamps per unit: 0 unit start, 1 tiles landed (lead warp), 2 stage-1 issued, 3 stage-1 done, 4 pass 1 done, 5 barrier passed,
6 stage-2 issued + k_f requested, 7 stage-2 done, 8 pass 3 done, 9 stage-3 issued, 10 stage-3 done, 11 pass 5 done,
12 stage-4 issued, 13 stage-4 done, 14 pass 6 done, 15 store issued."""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(3, 2, 64, 16)
names = ['tma wait', 'S1 issue', 'S1 wait', 'pass1', 'sync1', 'S2 issue+kf', 'S2 wait', 'pass3', 'sync+S3 issue', 'S3 wait', 'pass5',
         'sync+S4 issue', 'S4 wait', 'pass6', 'sync+store']
t0 = a[a > 0].min()
for pipe in range(3):
    for w in range(2):
        t = a[pipe, w]
        n = int((t[:, 15] > 0).sum())
        if n < 3:
            continue
        d = np.diff(t[:n], axis=1)                      # phase durations
        unit_time = np.diff(t[:n, 0])
        print(f'pipe {pipe} warp {"3" if w else "0 (lead)"}: {n} units, unit period mean {unit_time[1:].mean():.0f} cycles; first unit starts at {t[0,0]-t0}')
        mid = d[1:n - 1]
        print('   ' + '  '.join(f'{nm} {m:.0f}' for nm, m in zip(names, mid.mean(axis=0))))
# tensor-pipe occupancy timeline of CTA 0: intervals [issue, done] of each stage for the 3 pipelines
ev = []
for pipe in range(3):
    t = a[pipe, 0]
    n = int((t[:, 15] > 0).sum())
    for u in range(n):
        for (i, j, nm) in ((2, 3, 'S1'), (6, 7, 'S2'), (9, 10, 'S3'), (12, 13, 'S4')):
            ev.append((t[u, i] - t0, t[u, j] - t0, pipe, u, nm))
ev.sort()
print('first 40 MMA stages of CTA 0 (issue, done, duration, pipe, unit, stage):')
for e in ev[12:52]:
    print(f'   {e[0]:8d} {e[1]:8d} {e[1]-e[0]:6d}  p{e[2]} u{e[3]} {e[4]}')
