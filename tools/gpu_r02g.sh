#!/bin/bash
# ncu --set full of eval_slots on the candidate batch (windowed), old build against the refactored one
mkdir -p gpurun_out
for v in old sA_roll_tlog; do
  export NMB_LIB=$PWD/namaste_b200/variants/$v.so
  python tools/gpu_run_mode.py batch windowed 4 > gpurun_out/plain_$v.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:eval_slots -s 2 -c 1 -f -o gpurun_out/r02g_eval_slots_$v python tools/gpu_run_mode.py batch windowed 4 > gpurun_out/ncu_$v.log 2>&1
  tail -2 gpurun_out/ncu_$v.log
done
ls -la gpurun_out/*.ncu-rep
