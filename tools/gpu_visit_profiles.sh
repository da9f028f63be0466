#!/bin/bash
TAG=${1:-r02z9}
# usage: tools/gpu_visit_profiles.sh <tag>
# ncu evidence of a build: launch list of the bench command, full captures of the two dominant kernels
# (the sweep inside the bench command on configs[2]; eval_slots on the candidate batch)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu0.log 2>&1
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_walkers -s 4 -c 1 -f -o gpurun_out/${TAG}_sweep_kepler python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu2.log 2>&1
python tools/gpu_run_mode.py batch windowed 4 > gpurun_out/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_slots -s 2 -c 1 -f -o gpurun_out/${TAG}_eval_slots_batch python tools/gpu_run_mode.py batch windowed 4 > gpurun_out/ncu1.log 2>&1
TAG=$TAG python - <<'PY'
import csv, collections, os
rows=list(csv.reader(l for l in open('gpurun_out/%s_launches.csv' % os.environ['TAG']) if l.startswith('"')))
hdr=rows[0]; i_k=hdr.index('Kernel Name'); i_v=hdr.index('Metric Value'); i_g=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[r[i_k].split('(')[0][-50:]+' grid '+r[i_g]].append(float(r[i_v].replace(',','')))
for k,v in agg.items(): print("%-80s n=%3d  mean %.1f us  min %.1f" % (k, len(v), sum(v)/len(v)/1e3, min(v)/1e3))
PY
ls -la gpurun_out/${TAG}*
