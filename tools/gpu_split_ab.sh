#!/bin/bash
# walker-split chain check (both transports) and the walker-split leg of the bench alone.  usage: gpu_split_ab.sh <world>
W=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 2>&1 | grep MULTIGPU_CHECK
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29521 tools/gpu_split_leg.py 40 2>&1 | grep SPLIT_LEG
