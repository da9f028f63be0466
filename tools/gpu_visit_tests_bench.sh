#!/bin/bash
# tests, bench, and compute-sanitizer memcheck over the small all-kernels driver
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02s_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s_bench.json'))
print("value %.3e e2e %.3e fp32 %s fp32_pure %s" % (d['value'], d['e2e']['value'], json.dumps(d['fp32'])[:200], json.dumps(d['fp32_pure'])[:200]))
print("detrend", json.dumps(d['detrend']))
PY
python tools/gpu_sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1 && echo "all-kernels driver ok"; tail -1 gpurun_out/sanitize_plain.log | cut -c1-200
# (compute-sanitizer --tool memcheck|racecheck python tools/gpu_sanitize_run.py: the tool is closed on this pool,
#  profiles/r02l_compute_sanitizer_closed.log)
python tools/gpu_stress_determinism.py 200 2>&1 | tail -3 | tee gpurun_out/r02s_determinism.txt
