#!/bin/bash
# per-config launch lists of one module-level forward + backward (second of two iterations), ncu per-launch durations
# (cold-ish caches, serialised: shares of the step, not absolute times) -> gpurun_out/r2_launches.md
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
for w in ${WORKLOADS:-c2 c3 c4 c5 r1k r8k}; do
  W=$w ITERS=2 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fb_$w.csv \
    python tools/prof_step.py > gpurun_out/launches_fb_$w.log 2>&1
done
python tools/launchlist_md.py ${WORKLOADS:-c2 c3 c4 c5 r1k r8k} > gpurun_out/r2_launches.md; cat gpurun_out/r2_launches.md
