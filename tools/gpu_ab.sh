#!/bin/bash
for v in "" "NMB_BENCH_NOFLUSH=1"; do
 env $v python bench.py --steps 300 --warmup 5 --no-cpu > gpurun_out/ab.json 2>gpurun_out/ab.err; tail -2 gpurun_out/ab.err
 python - <<PY
import json
d=json.load(open('gpurun_out/ab.json'))
print("[$v] brute %.4f ms %.3e wt/s | windowed %.4f ms %.3e | e2e %.3e (win %.3e)" % (d['ms_per_step'], d['value'], d['windowed']['ms_per_step'], d['windowed']['value'], d['e2e']['value'], d['windowed']['e2e']))
PY
done
