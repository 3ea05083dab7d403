#!/bin/bash
mkdir -p gpurun_out
for mode in "" "--no-bind"; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/pcie_probe_ranks.py $mode 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version"
done | tee gpurun_out/pcie_ranks.md
