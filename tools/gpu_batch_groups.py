"""Independent candidates need no lock step: the 256-candidate batch stepped as G groups (each its own staged batch,
sampler and CUDA stream, driven from G host threads) against one group.  usage: python tools/gpu_batch_groups.py [steps]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import namaste_b200 as nb
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
wl = bench.make_batch(0)
C = wl["C"]
for G in (1, 2, 4, 8):
    per = C // G
    groups = []
    for g in range(G):
        sl = slice(g * per, (g + 1) * per)
        lc = nb.LightCurve(wl["curves"][sl], oversamp=wl["S"])
        smp = nb.EnsembleSampler(wl["W"], 6, nb.MonoOnlyLogProb, args=(lc, wl["priors"][sl]), seed=1 + g)
        smp.run_mcmc(wl["pos"][sl], 20, storechain=False)
        groups.append((lc, smp))
    torch.cuda.synchronize()
    def run(smp):
        smp.run_mcmc(None, steps, storechain=False)
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(smp,)) for _, smp in groups]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("G=%d groups of %3d candidates: %.1f us per step of all %d candidates, %.3e candidate-steps/s" % (
        G, per, 1e6 * dt / steps, C, steps * C / dt), flush=True)
    del groups
