#!/bin/bash
TAG=${1:-r02z9}
# 2-GPU visit: walker-split chain check (p2p + nccl transports) and the N = 2 bench line
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -15 | tee gpurun_out/${TAG}_multigpu_check_w2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g2.json 2> gpurun_out/${TAG}_bench_g2.err; echo "bench rc=$?"; tail -12 gpurun_out/${TAG}_bench_g2.err
TAG=$TAG python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/%s_bench_g2.json' % __import__('os').environ['TAG']))
    print("N=2 batch brute %.4f ms %.3e wt/s | windowed %.4f ms %.3e | e2e %.3e | mcmc %s" % (d['ms_per_step'], d['value'], d['windowed']['ms_per_step'], d['windowed']['value'], d['e2e']['value'], json.dumps(d['mcmc'])[:400]))
    print("walker_split", json.dumps(d['walker_split']))
except Exception as e:
    print("parse failed", e)
PY
