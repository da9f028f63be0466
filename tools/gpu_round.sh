#!/bin/bash
# One GPU-box visit: tests, bench (all workloads), reference arm, ncu launch list + full captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; tail -2 gpurun_out/bench_k2.err
for w in kepler tess epic; do
  python bench.py --steps 20 --warmup 3 --no-cpu --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu1.log 2>&1
for k in sweep_walkers eval_mixed eval_slots window_ranges window_direct; do
python bench.py --steps 5 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_$k.log 2>&1
done
python tools/gpu_run_mode.py kepler brute 4 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_walkers -s 2 -c 1 -o gpurun_out/prof_sweep_walkers_kepler -f python tools/gpu_run_mode.py kepler brute 4 > gpurun_out/ncu_sweep_kepler.log 2>&1
gzip -f gpurun_out/*.ncu-rep   # gpurun copies back at most 64 MiB
du -sh gpurun_out; ls -la gpurun_out | head -40
