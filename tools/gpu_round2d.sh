#!/bin/bash
# 2-GPU visit: NCCL multi-rank parity test, bench at --gpus 2 (both arms), PCIe probe with both ranks copying
# (the library is rebuilt on the box only if its source hash stamp disagrees with the tree)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -12 gpurun_out/topo.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('2 GPUs headline', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], d['config'].get('host_numa_binding'))
for k,v in d['roofline']['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, v['workload'][:60], 'fwd_ms %.3f convs/s %.3e kern_ms %.3f fwdbwd_ms %.3f e2e_ms %.2f' % (v['fwd']['ms_per_step'], v['fwd']['convs_per_sec'], v['kernels']['ms'], v['fwd_bwd']['ms_per_step'], v['e2e']['ms_per_step']))
PY
