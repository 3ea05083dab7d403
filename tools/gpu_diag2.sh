#!/bin/bash
echo "== 1 process"; timeout 200 python tools/diag_step.py plain 2>&1 | grep steps=
echo "== 1 process, OMP_NUM_THREADS=1"; OMP_NUM_THREADS=1 timeout 200 python tools/diag_step.py plain 2>&1 | grep steps=
echo "== torchrun 2 ranks, no binding"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/diag_step.py nobind 2>&1 | grep "steps=" | sort
echo "== torchrun 2 ranks, NUMA bind"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/diag_step.py bind 2>&1 | grep "steps=\|bound" | sort
