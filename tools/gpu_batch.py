"""BASELINE config 4 on one GPU's share: a batch of independent single-transit candidates (4000 points
each, 100 walkers each) evaluated and sampled as ONE staged batch (no collective; with G GPUs the
candidates are dealt round-robin, namaste_b200.distributed.shard_candidates).
usage: python tools/gpu_batch.py [candidates=256]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import namaste_b200 as nb

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, N, S = 100, 4000, 11
curves, pos, tabs = [], [], []
for k in range(C):
    wl = bench.make_workload("k2", seed_offset=k)
    rng = np.random.default_rng(7 + k)
    th = wl["theta"].copy()
    th[1], th[2], th[3] = rng.uniform(0, 0.9), rng.uniform(1.5, 6.0), rng.uniform(0.03, 0.15)   # SURVEY 8d jitter
    t = wl["lc"][:, 0]
    z = np.sqrt(th[1] ** 2 + (th[2] * (t + 0.45 * wl["cad"] - th[0])) ** 2)
    lc = wl["lc"].copy()
    lc[:, 1] = 1.0 - th[3] ** 2 * (z < 1.0) + 5.4e-5 * rng.standard_normal(N)
    curves.append(lc)
    p = th[None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
    pos.append(p)
    tab = wl["priors"].copy()
    tab[2, 1], tab[2, 3], tab[2, 4] = th[2] * 1.6, th[2] * 1.1, th[2] * 0.12
    tabs.append(tab)
pos, tabs = np.stack(pos), np.stack(tabs)
staged = nb.LightCurve(curves, oversamp=S)
d_pos = torch.from_numpy(pos).cuda()
d_tab = torch.from_numpy(tabs).cuda()
out = torch.empty((C, W), dtype=torch.float64, device="cuda")
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = {"candidates": C, "walkers_each": W, "N": N, "S": S}
for mode in ("brute", "windowed"):
    for _ in range(3):
        staged.lnprob_device(d_pos, d_tab, out=out, mode=mode)
    ms = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); staged.lnprob_device(d_pos, d_tab, out=out, mode=mode); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / 10
    res[mode] = {"ms_per_step": ms, "wt_per_s": C * W * N / (ms * 1e-3)}
smp = nb.EnsembleSampler(W, 6, nb.MonoOnlyLogProb, args=(staged, tabs), seed=3)
smp.run_mcmc(pos, 50)
smp.reset()
torch.cuda.synchronize()
t0 = time.perf_counter()
smp.run_mcmc(None, 50)
dt = time.perf_counter() - t0
res["mcmc_windowed"] = {"steps_per_s": 50 / dt, "candidate_steps_per_s": 50 * C / dt, "acceptance": float(smp.acceptance_fraction.mean())}
print(json.dumps(res))
