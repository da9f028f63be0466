"""A short windowed MCMC run of one bench workload (driver for an ncu launch list: per-kernel durations of the
sampler's half-steps).  usage: python tools/gpu_mcmc_kernels.py [workload] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import namaste_b200 as nb
name = sys.argv[1] if len(sys.argv) > 1 else "kepler"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
wl = bench.make_batch(0) if name == "batch" else bench.make_workload(name)
staged = bench.stage(nb, wl)
smp = nb.EnsembleSampler(wl["W"], 6, nb.MonoOnlyLogProb, args=(staged, wl["priors"]), seed=1)
smp.run_mcmc(wl["pos"], 100, storechain=False)      # let the prior-box walkers join the ensemble
torch.cuda.synchronize()
t0 = time.perf_counter()
smp.run_mcmc(None, steps, storechain=False)
dt = time.perf_counter() - t0
print("%s: %.1f us per half-step" % (name, 1e6 * dt / steps / 2))
