#!/bin/bash
TAG=${1:-r02z9}
# 8-GPU visit: walker-split chain check at world 4 and 8 (both transports), bench lines at N = 8 and N = 4
mkdir -p gpurun_out
for w in ${WORLDS:-8 4}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 2951$w tests/multigpu_check.py 2>&1 | grep MULTIGPU_CHECK | tee gpurun_out/${TAG}_multigpu_check_w$w.log
done
for n in ${WORLDS:-8 4}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g$n.json 2> gpurun_out/${TAG}_bench_g$n.err; echo "bench N=$n rc=$?"; tail -3 gpurun_out/${TAG}_bench_g$n.err
done
TAG=$TAG WORLDS="${WORLDS:-8 4}" python - <<'PY'
import json
for n in [int(x) for x in __import__("os").environ.get("WORLDS", "8 4").split()]:
    try:
        d=json.load(open('gpurun_out/%s_bench_g%d.json' % (__import__('os').environ['TAG'], n)))
        print("N=%d batch brute %.4f ms %.3e wt/s | windowed %.4f ms %.3e | e2e %.3e | mcmc win %.1f steps/s (%.3e cand-steps/s)" % (n, d['ms_per_step'], d['value'], d['windowed']['ms_per_step'], d['windowed']['value'], d['e2e']['value'], d['mcmc']['windowed']['steps_per_s'], d['mcmc']['windowed']['candidate_steps_per_s']))
        print("   walker_split", json.dumps(d['walker_split']))
    except Exception as e:
        print("parse failed", n, e)
PY
