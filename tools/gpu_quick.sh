#!/bin/bash
# tests + bench + per-kernel launch durations (ncu light pass)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 300 --warmup 5 --no-cpu > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("brute %.4f ms %.3e wt/s | windowed %.4f ms %.3e | e2e %.3e (win %.3e) | roofline frac %.3f" % (d['ms_per_step'], d['value'], d['windowed']['ms_per_step'], d['windowed']['value'], d['e2e']['value'], d['windowed']['e2e'], d['roofline']['frac']))
PY
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/launches.csv') if l.startswith('"')))
hdr=rows[0]; i_k=hdr.index('Kernel Name'); i_v=hdr.index('Metric Value'); i_g=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[r[i_k].split('(')[0][-50:]+' grid '+r[i_g]].append(float(r[i_v].replace(',','')))
for k,v in agg.items(): print("%-80s n=%3d  mean %.1f us  min %.1f" % (k, len(v), sum(v)/len(v)/1e3, min(v)/1e3))
PY
