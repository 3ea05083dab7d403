#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -2 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
for k,v in d['roofline']['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, 'fwd_ms %.3f eval_ms %.3f kern_ms %.3f fwdbwd_ms %.3f e2e_ms %.2f' % (v['fwd']['ms_per_step'], v['fwd_eval_cached_kf']['ms_per_step'], v['kernels']['ms'], v['fwd_bwd']['ms_per_step'], v['e2e']['ms_per_step']))
PY
