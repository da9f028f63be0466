#!/bin/bash
# Round 2, visit A: GPU tests, the default bench line (N = 1: BASELINE configs[2]), the reference arm,
# and the ncu launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02a_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02a_bench.json'))
    r=d['roofline']
    print("kepler brute %.4f ms (median %.4f) %.3e wt/s | windowed %.4f ms %.3e | e2e %.3e | callable %.3e | frac %.3f exec %.3f" % (
        d['ms_per_step'], d['ms_per_step_median'], d['value'], d['windowed']['ms_per_step'], d['windowed']['value'], d['e2e']['value'],
        d['e2e_callable']['value'], r['frac'], r['frac_executed']))
    for k in r['kernels']: print("  ", k['kernel'], "%.4f ms frac %.3f exec %.3f" % (k['ms_per_launch'], k['frac'], k['frac_executed']))
    print("mcmc", d['mcmc'])
    for key in ('walker_split','config2','config1','candidate_batch','gp','cpu_baseline'):
        print(key, json.dumps(d.get(key))[:900])
except Exception as e:
    print("parse failed", e)
PY
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv, collections
try:
    rows=list(csv.reader(l for l in open('gpurun_out/r02a_launches.csv') if l.startswith('"')))
    hdr=rows[0]; i_k=hdr.index('Kernel Name'); i_v=hdr.index('Metric Value'); i_g=hdr.index('Grid Size')
    agg=collections.defaultdict(list)
    for r in rows[1:]:
        agg[r[i_k].split('(')[0][-50:]+' grid '+r[i_g]].append(float(r[i_v].replace(',','')))
    for k,v in agg.items(): print("%-80s n=%3d  mean %.1f us  min %.1f" % (k, len(v), sum(v)/len(v)/1e3, min(v)/1e3))
except Exception as e:
    print("launch list failed", e)
PY
ls -la gpurun_out | tail -8
