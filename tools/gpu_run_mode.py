"""Run a few ensemble evaluations of one workload/mode (short driver for ncu captures).
usage: python tools/gpu_run_mode.py [workload] [mode] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import namaste_b200 as nb

name = sys.argv[1] if len(sys.argv) > 1 else "k2"
mode = sys.argv[2] if len(sys.argv) > 2 else "brute"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
wl = bench.make_batch(0) if name == "batch" else bench.make_workload(name)
staged = bench.stage(nb, wl)
d_pos = torch.from_numpy(np.ascontiguousarray(wl["pos"])).cuda()
ptab = wl["priors"] if wl["priors"].ndim == 3 else wl["priors"][None]
d_pri = torch.from_numpy(np.ascontiguousarray(ptab)).cuda()
out = torch.empty((wl["C"], wl["W"]), dtype=torch.float64, device="cuda")
for _ in range(reps):
    staged.lnprob_device(d_pos, d_pri, out=out, mode=mode)
torch.cuda.synchronize()
print("ok", float(out.sum()))
