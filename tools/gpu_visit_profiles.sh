#!/bin/bash
# tests, default bench, ncu captures of the two dominant kernels (sweep on kepler, eval_slots on the candidate batch)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02s_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s_bench.json'))
r=d['roofline']
print("kepler brute %.4f ms (median %.4f) %.3e wt/s | windowed %.4f ms %.3e | e2e %.3e | callable %.3e | frac %.3f exec %.3f" % (
    d['ms_per_step'], d['ms_per_step_median'], d['value'], d['windowed']['ms_per_step'], d['windowed']['value'], d['e2e']['value'],
    d['e2e_callable']['value'], r['frac'], r['frac_executed']))
for k in r['kernels']: print("  ", k['kernel'], "%.4f ms frac %.3f exec %.3f" % (k['ms_per_launch'], k['frac'], k['frac_executed']))
print("mcmc", d['mcmc']['windowed'], d['mcmc']['brute'])
print("split", d['walker_split']['steps_per_s'], d['walker_split']['bitwise_equal'])
print("config2", d['config2']['brute']['ms_per_step'], d['config2']['windowed']['ms_per_step'], d['config2']['mcmc']['windowed']['steps_per_s'])
print("config1", d['config1']['brute']['ms_per_step'], d['config1']['windowed']['ms_per_step'], d['config1']['mcmc_windowed']['steps_per_s'])
print("batch", d['candidate_batch']['brute']['ms_per_step'], d['candidate_batch']['windowed']['ms_per_step'], d['candidate_batch']['mcmc_windowed'])
PY
python tools/gpu_run_mode.py batch windowed 4 > gpurun_out/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_slots -s 2 -c 1 -f -o gpurun_out/r02s_eval_slots_batch python tools/gpu_run_mode.py batch windowed 4 > gpurun_out/ncu1.log 2>&1
python tools/gpu_run_mode.py kepler brute 4 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_walkers -s 2 -c 1 -f -o gpurun_out/r02s_sweep_kepler python tools/gpu_run_mode.py kepler brute 4 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/r02j*
