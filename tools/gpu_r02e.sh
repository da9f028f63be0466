#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_ab_lib.py namaste_b200/variants/old.so namaste_b200/variants/single_m5.so namaste_b200/variants/pair_m5.so namaste_b200/variants/pair_m4.so namaste_b200/variants/pair_m3.so -- kepler k2 batch tess 2>&1 | tee gpurun_out/r02e_ab_stage2.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
