"""Ad-hoc GPU diagnostics (not part of the product): parity statistics + quick timings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import namaste_b200 as nb
from oracle import namaste_oracle as o

def synth(cfg, n, cad, t0, sigma, S, W, seed=0):
    rng = np.random.default_rng(151203722 + cfg)
    t = t0 + cad * np.arange(n)
    theta = np.array([t[n // 2] + 0.37 * cad, 0.5, 3.0, 0.06, 0.34, 0.30])
    flux = o.get_value(theta, t, S) + sigma * rng.standard_normal(n)
    lc = np.column_stack((t, flux, np.full(n, sigma)))
    r2 = np.random.default_rng(cfg)
    pos = theta[None, :] * (1 + 1e-3 * r2.standard_normal((W, 6)))
    return lc, theta, pos

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print("fp64 peak (TFLOP/s, ms):", nb.measure_fp64_peak())
g = np.load(os.path.join(ROOT, "tests/golden/kat_epic203311200.npz"))
tab = torch.from_numpy(g["priors"]).cuda()
for name, cfg, n, cad, t0, sig, S, W in [("cfg2 K2", 2, 4000, 0.0204318, 2061.0, 5.4e-5, 11, 1024),
                                          ("cfg1-like", 1, 2828, 0.0204318, 2061.0, 5.4e-5, 10, 100),
                                          ("cfg3 Kepler", 3, 65000, 0.02043361, 131.5, 1e-4, 15, 4096)]:
    lc, theta, pos = synth(cfg, n, cad, t0, sig, S, W)
    ptab = g["priors"].copy(); ptab[0, :2] = [theta[0] - 0.2, theta[0] + 0.2]
    tab = torch.from_numpy(ptab).cuda()
    st = nb.LightCurve(lc, oversamp=S)
    dpos = torch.from_numpy(pos).cuda()
    out = torch.empty((1, W), dtype=torch.float64, device="cuda")
    res = {}
    for mode in ("brute", "windowed"):
        ms = timeit(lambda: st.lnprob_device(dpos, tab[None], out=out, mode=mode))
        res[mode] = out.cpu().numpy().copy()
        print("%-12s %-9s N=%d S=%d W=%d  %.3f ms  %.3e wt/s" % (name, mode, n, S, W, ms, W * n / (ms * 1e-3)))
    print("   brute vs windowed max rel:", np.max(np.abs(res["brute"] - res["windowed"]) / np.abs(res["brute"])))
    sub = slice(0, 8)
    ref = o.lnprob_ensemble(pos[sub], lc, {k: v for k, v in zip(
        ["a", "b", "c", "d", "e", "f"],
        [[r[0], r[1]] if int(r[2]) == 0 else [r[0], r[1], {1: "gaussian", 2: "normlim"}[int(r[2])], r[3], r[4]] for r in ptab])}, S)
    print("   vs oracle (8 walkers) max rel:", np.max(np.abs(res["brute"][0][sub] - ref) / np.abs(ref)))
    t0_ = time.perf_counter(); h = st.lnprob(pos, ptab, mode="brute"); t1_ = time.perf_counter()
    msh = min(timeit(lambda: st.lnprob(pos, ptab, mode="brute"), reps=10), 1e9)
    print("   host-buffer path: %.3f ms  %.3e wt/s" % (msh, W * n / (msh * 1e-3)))
