#!/bin/bash
# one 8-GPU visit: NCCL parity test of the sharded path, then the bench line at N = 8 (one rank per GPU, NUMA-bound staging)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -3 gpurun_out/bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_8gpu.json'))
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], d['clocks'])
for k,v in d['roofline']['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, 'fwd_ms %.3f eval_ms %.3f kern_ms %.3f fwdbwd_ms %.3f e2e_ms %.2f' % (v['fwd']['ms_per_step'], v['fwd_eval_cached_kf']['ms_per_step'], v['kernels']['ms'], v['fwd_bwd']['ms_per_step'], v['e2e']['ms_per_step']))
PY
