"""Where a walker-split half-step spends its time, on ONE GPU: a W-walker ensemble on the 500 k-point light curve
(tight ball, like the bench leg) stepped by the plain sampler, for W/2 = 8192, 4096, 2048, 1024 walkers per
half-step -- the per-rank shares of the 16384-walker split on 1, 2, 4, 8 GPUs, without the exchange."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import namaste_b200 as nb
wl = bench.make_workload("tess")
lc = nb.LightCurve(wl["lc"], oversamp=wl["S"])
rng = np.random.default_rng(5)
res = {}
for W in (16384, 8192, 4096, 2048):
    pos = wl["theta"][None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
    smp = nb.EnsembleSampler(W, 6, nb.MonoOnlyLogProb, args=(lc, wl["priors"]), seed=77)
    smp.run_mcmc(pos, 5, storechain=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 40
    smp.run_mcmc(None, n, storechain=False)
    dt = time.perf_counter() - t0
    res[W // 2] = 1e6 * dt / n / 2
    print("half-step of %5d walkers: %.1f us  (%.2f G samples/s)" % (W // 2, res[W // 2], (W // 2) * 4320 / (dt / n / 2) / 1e9), flush=True)
    del smp
