"""A/B of library builds on the GPU: brute-force step time (per-kernel split) of the bench workloads and a
bitwise comparison of lnprob against the first build.  Each build runs in its own process (NMB_LIB).
usage: python tools/gpu_ab_lib.py lib1.so lib2.so ... [-- workload ...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, bench
import namaste_b200 as nb
res = {}
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for name in sys.argv[1:]:
    wl = bench.make_batch(0) if name == "batch" else bench.make_workload(name)
    staged = bench.stage(nb, wl)
    d_pos = torch.from_numpy(np.ascontiguousarray(wl["pos"])).cuda()
    ptab = wl["priors"] if wl["priors"].ndim == 3 else wl["priors"][None]
    d_pri = torch.from_numpy(np.ascontiguousarray(ptab)).cuda()
    out = torch.empty((wl["C"], wl["W"]), dtype=torch.float64, device="cuda")
    r = {}
    for mode in ("brute", "windowed"):
        for _ in range(5): staged.lnprob_device(d_pos, d_pri, out=out, mode=mode)
        ms = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); staged.lnprob_device(d_pos, d_pri, out=out, mode=mode); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        staged.stage_timing(True)
        for _ in range(10):
            flush.zero_(); staged.lnprob_device(d_pos, d_pri, out=out, mode=mode)
        torch.cuda.synchronize()
        s1, s2, _ = staged.stage_times(); staged.stage_timing(False)
        r[mode] = {"ms": float(np.mean(ms)), "median": float(np.median(ms)), "s1": s1, "s2": s2,
                   "digest": int(np.bitwise_xor.reduce(out.cpu().numpy().view(np.uint64).ravel()))}
    res[name] = r
print("RESULT " + json.dumps(res))
''' % ROOT
args = sys.argv[1:]
libs = args[:args.index("--")] if "--" in args else args
wls = args[args.index("--") + 1:] if "--" in args else ["kepler", "k2", "batch"]
base = None
for lib in libs:
    env = dict(os.environ, NMB_LIB=os.path.abspath(lib))
    p = subprocess.run([sys.executable, "-c", CHILD] + wls, env=env, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        print(lib, "FAILED", p.stderr[-600:]); continue
    r = json.loads(line[0][7:])
    base = base or r
    for w in wls:
        for mode in ("brute", "windowed"):
            x = r[w][mode]
            print("%-34s %-7s %-8s %.4f ms (median %.4f)  s1 %.4f s2 %.4f  same_bits=%s" % (
                os.path.basename(lib), w, mode, x["ms"], x["median"], x["s1"], x["s2"], x["digest"] == base[w][mode]["digest"]))
