"""Phase timeline (NMB_TRACE=1) of a plain 1024-walker windowed evaluation on the 500 k-point light curve: the
per-rank share of a walker-split half-step on 8 GPUs.  usage: NMB_TRACE=1 python tools/gpu_trace_split.py [W]"""
import os, sys
os.environ.setdefault("NMB_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import namaste_b200 as nb
from namaste_b200._lib import load, check, ptr
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
wl = bench.make_workload("tess")
staged = nb.LightCurve(wl["lc"], oversamp=wl["S"])
rng = np.random.default_rng(5)
pos = wl["theta"][None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
d_pos = torch.from_numpy(np.ascontiguousarray(pos)).cuda()
d_pri = torch.from_numpy(wl["priors"]).cuda()[None].contiguous()
out = torch.empty((1, W), dtype=torch.float64, device="cuda")
lib = load()
buf = np.zeros(16, dtype=np.uint64)
names = {1: "s1 consts", 2: "s1 window done", 3: "s1 push done", 8: "s2 first block", 9: "s2 items known", 10: "s2 last slot evaluated", 11: "s2 last warp done"}
for rep in range(5):
    staged.lnprob_device(d_pos, d_pri, out=out, mode="windowed")
    check(lib.nmb_debug_trace(staged._h, ptr(buf)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); staged.lnprob_device(d_pos, d_pri, out=out, mode="windowed"); e1.record()
    torch.cuda.synchronize()
    check(lib.nmb_debug_trace(staged._h, ptr(buf)))
    t0 = int(buf[0])
    print("W=%d event %.1f us (%s) |" % (W, 1e3 * e0.elapsed_time(e1), "+".join(staged.last_kernels())),
          " | ".join("%s +%.1f" % (names[k], (int(buf[k]) - t0) / 1e3) for k in sorted(names) if buf[k]), flush=True)
