#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch list, ncu full capture of the top kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 300 --warmup 5 > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; tail -2 gpurun_out/bench_k2.err; cat gpurun_out/bench_k2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu1.log 2>&1
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lnprob_sweep -s 4 -c 2 -o gpurun_out/prof_sweep python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lnprob_window -s 4 -c 2 -o gpurun_out/prof_window python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
