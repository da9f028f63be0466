"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library once,
on sizes that finish in seconds under the tool.  Ensemble sizes are chosen so that each stage-2 path runs
(direct blocks, mixed, queue), the sampler takes a few steps in both modes, and the batch pre-fit kernels, the
GP solve and the posterior summary run once.
usage: compute-sanitizer --tool racecheck python tools/gpu_sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import namaste_b200 as nb
from namaste_b200 import k2flatten
from conftest import load_golden

g = load_golden("synth_k2_s11")
rng = np.random.default_rng(3)
lc = nb.LightCurve(g["lc"][1500:2500], oversamp=int(g["oversamp"]))
base = g["walkers"][0]
done = []
for W in (24, 320, 1300):
    pos = base[None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
    pos[: W // 8, 2] *= rng.uniform(0.05, 0.5, W // 8)          # slow walkers: queued ranges
    for mode in ("brute", "windowed"):
        for dt in ("f64", "strict", "f32"):
            r = lc.lnprob(pos, g["priors"], mode=mode, dtype=dt)
            assert np.isfinite(r).all()
        done.append((W, mode) + lc.last_kernels())
for mode in ("windowed", "brute"):
    smp = nb.EnsembleSampler(64, 6, nb.MonoOnlyLogProb, args=(lc, g["priors"]), seed=5, mode=mode)
    smp.run_mcmc(base[None, :] * (1 + 1e-4 * rng.standard_normal((64, 6))), 6)
    assert np.isfinite(smp.lnprobability).all()
    s = smp.summary(6)
    del smp
batch = nb.LightCurve([g["lc"][1600:2400], g["lc"][1700:2300]], oversamp=int(g["oversamp"]))
r = batch.lnprob(np.stack([pos[:40], pos[40:80]]), np.stack([g["priors"]] * 2), mode="brute")
fl = load_golden("flatten_epic203311200")
out = k2flatten.ReduceNoiseBatch([fl["lc"][:600], fl["gap_lc"][:500]])
keep = k2flatten.AnomCutDiffBatch([fl["lc"][:600, 1]], 3.75)
mask = k2flatten.window_mask_batch([fl["lc"][:600, 0]], 2070.0, 2.0)
kern = nb.RealTerm(log_a=-14.0, log_c=-1.0) + nb.JitterTerm(-9.8)
kern.freeze_parameter("terms[1]:log_sigma")
gp = nb.GP(kernel=kern, mean=nb.MonotransitModel(*base), fit_mean=True)
v = nb.MonoLogProb(np.concatenate([[-14.0, -1.0], base]), g["lc"][1500:2500], None, gp)
print("SANITIZE_RUN_OK", done, float(v))
