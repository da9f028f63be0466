"""Phase timeline of one ensemble evaluation (NMB_TRACE=1): where the microseconds of a step go.
usage: NMB_TRACE=1 python tools/gpu_trace.py [workload] [mode] [half]"""
import os, sys
os.environ.setdefault("NMB_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
import bench
import namaste_b200 as nb
from namaste_b200._lib import load, check, ptr

name = sys.argv[1] if len(sys.argv) > 1 else "k2"
mode = sys.argv[2] if len(sys.argv) > 2 else "brute"
half = len(sys.argv) > 3
wl = bench.make_workload(name)
staged = nb.LightCurve(wl["lc"], oversamp=wl["S"])
pos = wl["pos"][: wl["W"] // 2] if half else wl["pos"]
d_pos = torch.from_numpy(np.ascontiguousarray(pos)).cuda()
d_pri = torch.from_numpy(wl["priors"]).cuda()[None].contiguous()
out = torch.empty((1, len(pos)), dtype=torch.float64, device="cuda")
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
lib = load()
buf = np.zeros(16, dtype=np.uint64)
names = {7: "direct walker finished", 5: "direct consts in smem", 6: "direct first round done", 1: "s1 consts ready", 2: "s1 sweep/window done", 3: "s1 fold done", 4: "s1 scan done", 8: "s2 first block start",
         9: "s2 first items known", 10: "s2 last item evaluated", 11: "s2 last warp done",
         12: "direct first start", 13: "direct last start", 14: "direct rounds done", 15: "direct folded"}
for rep in range(6):
    staged.lnprob_device(d_pos, d_pri, out=out, mode=mode)
    check(lib.nmb_debug_trace(staged._h, ptr(buf)))
    if rep >= 4 or True:
        flush.zero_() if rep % 2 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); staged.lnprob_device(d_pos, d_pri, out=out, mode=mode); e1.record()
        torch.cuda.synchronize()
        check(lib.nmb_debug_trace(staged._h, ptr(buf)))
        t0 = int(buf[0])
        print("rep %d (%s) event time %.1f us" % (rep, "L2 flushed" if rep % 2 else "warm", 1e3 * e0.elapsed_time(e1)),
              " | ".join("%s +%.1f" % (names[k], (int(buf[k]) - t0) / 1e3) for k in sorted(names) if buf[k]))
