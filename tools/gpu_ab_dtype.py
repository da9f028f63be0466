"""A/B of the arithmetic variants of the transit model on the GPU: accuracy against the golden
fixtures (reference output) and step time on the bench workloads, per dtype and mode.
usage: python tools/gpu_ab_dtype.py [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import bench
import namaste_b200 as nb
from conftest import load_golden

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
DT = ("f64", "strict")
res = {"accuracy": {}, "timing": {}}
for name in ["config1_epic203311200", "synth_k2_s11", "synth_kepler_s15", "synth_tess_s10", "synth_tess_s1"]:
    g = load_golden(name)
    S = int(g["oversamp"]) if "oversamp" in g.files else 10
    lc = nb.LightCurve(g["lc"], oversamp=S)
    ref = g["ll"] + g["lp"]
    fin = np.isfinite(ref)
    for dt in DT:
        for mode in ("brute", "windowed"):
            lnp = lc.lnprob(g["walkers"], g["priors"], mode=mode, dtype=dt)
            r = np.abs(lnp[fin] - ref[fin]) / np.abs(ref[fin])
            res["accuracy"][f"{name}/{dt}/{mode}"] = float(r.max())
            print(name, dt, mode, "max rel", r.max(), flush=True)
    lc.close()

flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for wname in ("k2", "kepler", "tess", "epic"):
    wl = bench.make_workload(wname)
    staged = nb.LightCurve(wl["lc"], oversamp=wl["S"])
    d_pos = torch.from_numpy(np.ascontiguousarray(wl["pos"])).cuda()
    d_pri = torch.from_numpy(wl["priors"]).cuda()[None].contiguous()
    out = {dt: torch.empty((1, wl["W"]), dtype=torch.float64, device="cuda") for dt in DT}
    for mode in ("brute", "windowed"):
        for dt in DT:
            for _ in range(5):
                staged.lnprob_device(d_pos, d_pri, out=out[dt], mode=mode, dtype=dt)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in ev:
                flush.zero_()
                a.record()
                staged.lnprob_device(d_pos, d_pri, out=out[dt], mode=mode, dtype=dt)
                b.record()
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            res["timing"][f"{wname}/{mode}/{dt}"] = ms
            print(wname, mode, dt, f"{ms*1e3:.1f} us", flush=True)
        a_, b_ = out["f64"].cpu().numpy(), out["strict"].cpu().numpy()
        print(wname, mode, "relaxed vs strict max rel", float(np.max(np.abs(a_ - b_) / np.abs(b_))), flush=True)
    staged.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_dtype.json", "w"), indent=1)
