#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -30 gpurun_out/r02b_bench.err
python tools/gpu_ab_lib.py namaste_b200/variants/old.so namaste_b200/variants/sweep_t4_m8.so namaste_b200/variants/sweep_t4_m7.so namaste_b200/variants/sweep_t4_m6.so namaste_b200/variants/sweep_t2_m8.so namaste_b200/variants/sweep_t2_m7.so 2>&1 | tee gpurun_out/r02b_ab_sweep.txt
