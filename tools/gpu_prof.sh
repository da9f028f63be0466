#!/bin/bash
# ncu launch list + full captures for the kernels given as regex in $1 (comma separated)
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu1.log 2>&1
for k in $(echo $1 | tr ',' ' '); do
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/prof_$k python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
