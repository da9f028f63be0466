"""Step time vs ensemble size on the K2-like light curve, back-to-back (warm L2) and with L2 flushed (diagnostics;
SIZES=256,512,... selects the sizes, NMB_MIXED_MAXCW / NMB_DIRECT select the stage-2 path)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import namaste_b200 as nb
import bench
SIZES = [int(x) for x in os.environ["SIZES"].split(",")] if os.environ.get("SIZES") else (256, 1024, 4096, 16384)
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for W in SIZES:
    bench.WORKLOADS["k2"] = (2, 4000, 0.0204318, 2061.0, 5.4e-5, 11, W)
    wl = bench.make_workload("k2")
    if os.environ.get("TIGHT"):
        wl["pos"] = wl["theta"][None, :] * (1 + 1e-3 * np.random.default_rng(1).standard_normal((W, 6)))
    st = nb.LightCurve(wl["lc"], oversamp=11)
    dpos = torch.from_numpy(wl["pos"]).cuda(); dpri = torch.from_numpy(wl["priors"]).cuda()[None].contiguous()
    out = torch.empty((1, W), dtype=torch.float64, device="cuda")
    for mode in ("brute", "windowed"):
        for _ in range(5): st.lnprob_device(dpos, dpri, out=out, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): st.lnprob_device(dpos, dpri, out=out, mode=mode)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        cold = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); st.lnprob_device(dpos, dpri, out=out, mode=mode); e1.record()
            torch.cuda.synchronize(); cold.append(e0.elapsed_time(e1))
        print("W=%5d %-8s warm %.1f us  L2-flushed %.1f us  %s" % (W, mode, ms * 1e3, 1e3 * float(np.median(cold)), "+".join(st.last_kernels())), flush=True)
