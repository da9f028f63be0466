#!/usr/bin/env python
"""Benchmark of the Namaste likelihood hot path on B200 (driver contract: one JSON line).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload k2|kepler|epic]

Metric (BASELINE.json): walker x timestamp log-probability evaluations per second.
One *step* = one full-ensemble lnprob evaluation (W walkers x N timestamps x S supersamples,
likelihood + priors) on synthetic input of BASELINE config 2 ("synthetic K2 long-cadence
light curve, 4k points, supersampling x11, 1024 walkers").  With N > 1 GPUs every rank runs
an independent candidate of that shape (the reference's batch-of-candidates mode: no data-path
collective), so per-GPU work is fixed ("weak" scaling) and `value` is the sum over ranks.

`value`   : brute-force kernel (every supersample of every timestamp is visited and classified),
            inputs resident in HBM, CUDA-event timed per step with an L2 flush between steps.
`e2e`     : the same evaluation through the public host API (numpy params in, numpy lnprob out):
            pinned H2D of the walker positions and D2H of the log-probabilities every step.
`windowed`: the exact prefix-sum/window kernel (product default), reported separately and never
            as a roofline fraction.
`roofline`: FP64-pipe bound; algorithmic FLOPs per SURVEY.md 8(d) convention / event time,
            against a DFMA microbenchmark measured in the same process.
`cpu_baseline`: the NumPy oracle (port of the reference's CPU path) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cfg id, N, cadence, t0, sigma, S, W)
    "k2": (2, 4000, 0.0204318, 2061.0, 5.4e-5, 11, 1024),
    "kepler": (3, 65000, 0.02043361, 131.5, 1e-4, 15, 4096),
    "epic": (1, 2828, 0.0204318, 2061.0, 5.4e-5, 10, 100),
    "tess": (5, 500000, 2.0 / 1440, 1325.0, 6e-4, 10, 2048),
}
WORKLOAD_DESC = {
    "k2": "BASELINE configs[1]: synthetic K2 long-cadence, N=4000, S=11, W=1024",
    "kepler": "BASELINE configs[2]: synthetic Kepler long-cadence, N=65000, S=15, W=4096",
    "epic": "BASELINE configs[0] shape: N=2828, S=10, W=100 (synthetic noise)",
    "tess": "BASELINE configs[4] per-GPU share: synthetic TESS 2-min, N=500000, S=10, W=2048",
}


# --------------------------------------------------------------------------------------------
# synthetic workload (SURVEY.md 8d recipe).  The noiseless model is drawn with plain NumPy
# geometry + the GPU model itself is NOT used here; the oracle is only needed for cpu_baseline.
def make_workload(name, seed_offset=0):
    cfg, n, cad, t0, sigma, S, W = WORKLOADS[name]
    rng = np.random.default_rng(151203722 + cfg + 1000 * seed_offset)
    t = t0 + cad * np.arange(n)
    theta = np.array([t[n // 2] + 0.37 * cad, 0.5, 3.0, 0.06, 0.34, 0.30])
    # box-shaped stand-in for the transit (depth p^2) + white noise: timing does not depend on the
    # exact injected shape, and this keeps bench.py independent of oracle/ for the GPU arm
    z = np.sqrt(theta[1] ** 2 + (theta[2] * (t + 0.45 * cad - theta[0])) ** 2)
    flux = 1.0 - theta[3] ** 2 * (z < 1.0) + sigma * rng.standard_normal(n)
    lc = np.column_stack((t, flux, np.full(n, sigma)))
    tdur = 2 * (1 + theta[3]) * np.sqrt(1 - (theta[1] / (1 + theta[3])) ** 2) / theta[2]
    ptab = np.array([
        [theta[0] - 0.3 * tdur, theta[0] + 0.3 * tdur, 0, 0, 0, 0],
        [0.0, 1.25, 0, 0, 0, 0],
        [0, theta[2] * 1.6, 2, theta[2] * 1.1, theta[2] * 0.12, 0],
        [0.02, 0.25, 0, 0, 0, 0],
        [0, 1.0, 1, theta[4], 0.038, 0],
        [0, 1.0, 1, theta[5], 0.022, 0],
    ])
    # initial ensemble: 95% tight ball around the truth, 5% uniform in the prior box
    pos = theta[None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
    nbox = max(1, W // 20)
    pos[:nbox] = ptab[:, 0] + (ptab[:, 1] - ptab[:, 0]) * rng.uniform(0, 1, (nbox, 6))
    return dict(name=name, lc=lc, theta=theta, priors=ptab, pos=pos, S=S, W=W, N=n, cad=cad)


def count_samples(wl):
    """Per-ensemble counts for the FLOP formula of SURVEY.md 8(d): fine samples in total, interior
    (planet inside the disc) and ingress/egress samples, and timestamps with at least one
    in-transit sample (those are evaluated by stage 2, the others by stage 1 alone)."""
    t, S, cad = wl["lc"][:, 0], wl["S"], wl["cad"]
    off = cad * np.arange(0, 1.0, 1 / S)
    m_int = m_edge = t_in = 0
    for par in wl["pos"]:
        tcen, b, vel, rp = par[0], par[1], par[2], abs(par[3])
        h2 = (1 + rp) ** 2 - b * b
        if h2 <= 0:
            continue
        if vel == 0:
            sel = np.arange(len(t))
        else:
            half = np.sqrt(h2) / abs(vel) + 2 * cad
            sel = np.where(np.abs(t - tcen) < half)[0]
        if len(sel) == 0:
            continue
        z = np.sqrt(b * b + (vel * ((t[sel, None] + off[None, :]) - tcen)) ** 2)
        inner = z < abs(1 - rp)
        limb = (z < 1 + rp) & ~inner
        m_int += int(np.sum(inner))
        m_edge += int(np.sum(limb))
        t_in += int(np.sum(np.any(inner | limb, axis=1)))
    return wl["W"] * wl["N"] * S, m_int, m_edge, t_in


def algorithmic_flops(wl):
    """7 per fine sample + 143 per interior + 180 per ingress/egress sample + 5 per
    timestamp + 40 per walker, split between the two kernels of the brute-force pipeline: the sweep
    visits every fine sample and closes the out-of-transit timestamps, stage 2 adds the in-transit
    terms, the residuals of the in-transit timestamps and the priors (its recomputation of z for
    those samples is implementation work and is not counted)."""
    m, m_int, m_edge, t_in = count_samples(wl)
    wn = wl["W"] * wl["N"]
    sweep = m * 7 + (wn - t_in) * 5
    evalk = m_int * 143 + m_edge * 180 + t_in * 5 + wl["W"] * 40
    return sweep + evalk, sweep, evalk, (m, m_int, m_edge, t_in)


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t_begin or ts > t_end + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def _oracle_walker(args):
    from oracle import namaste_oracle as o
    par, lc, pri, S = args
    return o.mono_only_logprob(par, lc, pri, S)


def _priors_dict(tab):
    names = ["mean:tcen", "mean:b", "mean:vel", "mean:RpRs", "mean:LD1", "mean:LD2"]
    kinds = {1: "gaussian", 2: "normlim", 3: "evans"}
    return {nm: ([r[0], r[1]] if int(r[2]) == 0 else [r[0], r[1], kinds[int(r[2])], r[3], r[4]])
            for nm, r in zip(names, tab)}


def cpu_reference_rate(wl, budget_s, pool, cores, nwalk=None):
    """wt/s of the NumPy oracle (the reference's CPU path restated) on `cores` processes."""
    pri = _priors_dict(wl["priors"])
    W = wl["W"]
    pool.map(_oracle_walker, [(wl["pos"][i % W], wl["lc"], pri, wl["S"]) for i in range(cores)])  # imports
    t0 = time.perf_counter()
    pool.map(_oracle_walker, [(wl["pos"][i % W], wl["lc"], pri, wl["S"]) for i in range(cores)])
    per_walker = (time.perf_counter() - t0)  # one walker per core, in parallel
    if nwalk is None:
        nwalk = int(max(cores, min(64 * W, budget_s / max(per_walker, 1e-6) * cores)))
    jobs = [(wl["pos"][i % W], wl["lc"], pri, wl["S"]) for i in range(nwalk)]
    t0 = time.perf_counter()
    pool.map(_oracle_walker, jobs, chunksize=max(1, nwalk // (cores * 4)))
    dt = time.perf_counter() - t0
    return nwalk * wl["N"] / dt, nwalk, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (NumPy oracle port,
    multiprocessing pool like emcee's threads=) on this box's host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    wl = make_workload(args.workload)
    cores = len(os.sched_getaffinity(0))
    pool = mp.Pool(cores)
    # size each step so that warmup+steps finish within ~2 minutes
    pri = _priors_dict(wl["priors"])
    pool.map(_oracle_walker, [(wl["pos"][i % wl["W"]], wl["lc"], pri, wl["S"]) for i in range(cores)])  # imports
    t0 = time.perf_counter()
    pool.map(_oracle_walker, [(wl["pos"][i % wl["W"]], wl["lc"], pri, wl["S"]) for i in range(cores)])
    per = time.perf_counter() - t0
    total_steps = args.steps + args.warmup
    nwalk = int(max(cores, min(wl["W"], (100.0 / total_steps) / max(per, 1e-6) * cores)))
    for _ in range(args.warmup):
        cpu_reference_rate(wl, 0, pool, cores, nwalk=nwalk)
    t_begin = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, _, _ = cpu_reference_rate(wl, 0, pool, cores, nwalk=nwalk)
        rates.append(r)
    elapsed = time.perf_counter() - t_begin
    pool.close()
    value = args.steps * nwalk * wl["N"] / elapsed
    sample = "%d of %d walkers per step (full N=%d, S=%d), NumPy oracle in a %d-process pool" % (
        nwalk, wl["W"], wl["N"], wl["S"], cores)
    line = {
        "impl": "reference", "metric": "lnprob_walker_timestamp_evals_per_sec", "value": value, "unit": "wt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "wt/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "wt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import namaste_b200 as nb

    wl = make_workload(args.workload, seed_offset=rank)
    W, N, S = wl["W"], wl["N"], wl["S"]
    staged = nb.LightCurve(wl["lc"], oversamp=S)
    d_pos = torch.from_numpy(wl["pos"]).cuda()
    d_pri = torch.from_numpy(wl["priors"]).cuda()[None].contiguous()
    out = torch.empty((1, W), dtype=torch.float64, device="cuda")
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def device_loop(mode, steps, sampler=None):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_begin = time.perf_counter()
        for e0, e1 in evs:
            if not os.environ.get('NMB_BENCH_NOFLUSH'):
                flush.zero_()                  # L2 flush between timed iterations (not timed)
            e0.record(stream)
            staged.lnprob_device(d_pos, d_pri, out=out, mode=mode)
            e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_end = time.perf_counter()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        return ms, t_begin, t_end

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident measurement (brute force = `value`) ---------------------------------
    device_loop("brute", max(args.warmup, 3))
    clocks = ClockSampler(local)
    clocks.start()
    l0 = nb.launch_count()
    ms_brute, t_b, t_e = device_loop("brute", args.steps)
    launches = nb.launch_count() - l0
    # keep the same loop running ~1 s longer so nvidia-smi (100 ms period) sees it under load
    t_stop = time.perf_counter() + (0.0 if args.profile else 1.0)
    while time.perf_counter() < t_stop:
        for _ in range(200):
            staged.lnprob_device(d_pos, d_pri, out=out, mode="brute")
        torch.cuda.synchronize()
    clk = clocks.stop(t_b, time.perf_counter())
    ms_brute = max_over_ranks(ms_brute)
    # per-kernel durations of the same loop (CUDA events between the two kernels of every step)
    staged.stage_timing(True)
    device_loop("brute", args.steps)
    ms_s1, ms_s2, _ = staged.stage_times()
    k_s1, k_s2 = staged.last_kernels()   # stage 2 is eval_mixed_kernel or eval_slots_kernel, by ensemble size
    staged.stage_timing(False)
    device_loop("windowed", max(args.warmup, 3))
    ms_win, _, _ = device_loop("windowed", args.steps)
    ms_win = max_over_ranks(ms_win)
    ref_f64 = out.cpu().numpy().copy()
    # FP32 variant (reported separately): windowed stage 1 + FP32 transit model in stage 2
    def device_loop_dtype(steps, dtype, mode="windowed"):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for e0, e1 in evs:
            flush.zero_()
            e0.record(stream)
            staged.lnprob_device(d_pos, d_pri, out=out, mode=mode, dtype=dtype)
            e1.record(stream)
        torch.cuda.synchronize()
        return sum(e0.elapsed_time(e1) for e0, e1 in evs)
    fin = np.isfinite(ref_f64) & (np.abs(ref_f64) < 1e15)
    device_loop_dtype(max(args.warmup, 3), "f32")
    ms_f32 = max_over_ranks(device_loop_dtype(args.steps, "f32"))
    f32_out = out.cpu().numpy().copy()
    f32_err = float(np.max(np.abs(f32_out[fin] - ref_f64[fin]) / np.abs(ref_f64[fin])))
    # strict variant (the reference's own rounding sequence in the hot occultation cases), brute force
    device_loop_dtype(max(args.warmup, 3), "strict", "brute")
    ms_strict = max_over_ranks(device_loop_dtype(args.steps, "strict", "brute"))
    strict_out = out.cpu().numpy().copy()
    strict_diff = float(np.max(np.abs(strict_out[fin] - ref_f64[fin]) / np.abs(ref_f64[fin])))
    staged.lnprob_device(d_pos, d_pri, out=out, mode="windowed")  # leave the FP64 result in `out`
    ref_out = out.cpu().numpy().copy()

    # ---- end to end through the public host API (numpy in / numpy out) -----------------------
    def e2e_loop(mode, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = staged.lnprob(wl["pos"], wl["priors"], mode=mode)
        dt = time.perf_counter() - t0
        return max_over_ranks(dt), res

    e2e_loop("brute", max(args.warmup, 3))
    dt_e2e, res_host = e2e_loop("brute", args.steps)
    e2e_loop("windowed", max(args.warmup, 3))
    dt_e2e_win, res_host_w = e2e_loop("windowed", args.steps)
    assert np.allclose(res_host_w, ref_out[0], rtol=1e-12, atol=0, equal_nan=True)
    assert np.allclose(res_host, ref_out[0], rtol=1e-12, atol=0, equal_nan=True)

    # ---- MCMC steps/s with the device-resident sampler (proposal + lnprob + accept on the GPU) ---
    mcmc = {}
    n_mcmc = 20 if args.profile else 200
    for mode in ("windowed", "brute"):
        smp = nb.EnsembleSampler(W, 6, nb.MonoOnlyLogProb, args=(staged, wl["priors"]), mode=mode, seed=1 + rank)
        smp.run_mcmc(wl["pos"], n_mcmc)     # warm-up: also grows the device-resident chain to its final capacity
        smp.reset()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        smp.run_mcmc(None, n_mcmc)          # returns after the device finished and pos/lnprob were read back
        dt = max_over_ranks(time.perf_counter() - t0)
        mcmc[mode] = {"steps_per_s": n_mcmc / dt, "wt_per_s": n_mcmc * W * N * world / dt,
                      "acceptance": float(smp.acceptance_fraction.mean())}
        del smp

    # ---- GP noise-model objective (the reference's default, MonoLogProb): reported separately ----------
    import ctypes
    from namaste_b200._lib import load as _load, check as _check, ptr as _ptr
    kfix = np.array([np.log(np.var(wl["lc"][:, 1])), -np.log(3.0), np.log(np.median(wl["lc"][:, 2]))])
    gp_pos = np.hstack([kfix[None, :2] + 0.1 * np.random.default_rng(3 + rank).standard_normal((W, 2)), wl["pos"]])
    d_gp = torch.from_numpy(np.ascontiguousarray(gp_pos)).cuda()
    d_kfix = kfix.copy()
    gp_out = torch.empty((1, W), dtype=torch.float64, device="cuda")

    def gp_loop(steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for e0, e1 in evs:
            flush.zero_()
            e0.record(stream)
            _check(_load().nmb_gp_lnprob(staged._h, _ptr(d_gp), W, 0, _ptr(d_kfix), 3, None, _ptr(gp_out), None,
                                         ctypes.c_void_p(stream.cuda_stream)), "gp")
            e1.record(stream)
        torch.cuda.synchronize()
        return sum(e0.elapsed_time(e1) for e0, e1 in evs)
    gp_steps = max(3, min(args.steps, 20))
    gp_loop(3)
    ms_gp = max_over_ranks(gp_loop(gp_steps)) / gp_steps

    # ---- one ensemble split by walker over the ranks (BASELINE config 5), NCCL all-gather per half-step ----
    split = None
    if world > 1 and not args.profile:
        from namaste_b200.distributed import SplitEnsembleSampler
        wl5 = make_workload("tess")
        W5 = 16384
        rng5 = np.random.default_rng(5)      # same ensemble on every rank (replicated, then sliced)
        pos5 = wl5["theta"][None, :] * (1 + 1e-3 * rng5.standard_normal((W5, 6)))
        lc5 = nb.LightCurve(wl5["lc"], oversamp=wl5["S"])
        sp = SplitEnsembleSampler(lc5, wl5["priors"], W5, seed=77)
        sp.run_mcmc(pos5, 3, storechain=False)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n5 = 20
        sp.run_mcmc(None, n5, storechain=False)
        dt5 = max_over_ranks(time.perf_counter() - t0)
        split = {"steps_per_s": n5 / dt5, "wt_per_s": n5 * W5 * wl5["N"] / dt5, "walkers": W5, "N": wl5["N"], "S": wl5["S"],
                 "scaling": "strong", "collective": "NCCL all-gather of the updated half-ensemble slices, twice per step",
                 "note": "BASELINE configs[4]: synthetic TESS 2-min light curve, one 16384-walker ensemble split by walker"}
        del sp, lc5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    units = W * N * world
    value = units / (ms_brute / args.steps * 1e-3)
    flops, fl_sweep, fl_eval, (m, m_int, m_edge, t_in) = algorithmic_flops(wl)
    peak_tf, _ = nb.measure_fp64_peak()
    step_s = ms_brute / args.steps * 1e-3
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload) or {}
        except Exception:
            traffic = {}
    kernels = []
    for name, ms, fl in ((k_s1, ms_s1, fl_sweep), (k_s2, ms_s2, fl_eval)):
        tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        kernels.append({"kernel": name, "ms_per_launch": ms,
                        "share_of_step": ms / (ms_s1 + ms_s2) if ms_s1 + ms_s2 > 0 else None,
                        "flops_per_launch": fl, "achieved": tf, "frac": tf / peak_tf, "traffic": traffic.get(name)})
    dom = max(kernels, key=lambda k: k["ms_per_launch"])
    line = {
        "metric": "lnprob_walker_timestamp_evals_per_sec", "value": value, "unit": "wt/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_brute / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "N": N, "S": S, "W": W,
                   "step": "one full-ensemble lnprob evaluation (likelihood + priors), brute-force kernel",
                   "arithmetic": "NMB_F64: FP64 with contracted multiply-adds and <= 1.5-ulp quotients / roots in the two "
                                 "hot occultation cases (lnprob within 1e-13 of the reference, tolerance 1e-10); the "
                                 "reference's own rounding sequence is timed under 'strict'",
                   "parallelism": "independent candidate per GPU, no collective" if world > 1 else "single GPU",
                   "l2": "192 MiB buffer zeroed between timed iterations (inputs are 0.15 MB, L2/SMEM resident by design)"},
        "clocks": clk,
        "e2e": {"value": units / (dt_e2e / args.steps), "unit": "wt/s",
                "h2d_bytes_per_step": int(W * 6 * 8 + 36 * 8), "d2h_bytes_per_step": int(W * 8),
                "ms_per_step": 1e3 * dt_e2e / args.steps,
                "note": "LightCurve.lnprob(numpy params) -> numpy lnprob; light curve staged once like emcee args="},
        "gpu_launches": int(launches),
        "mcmc": {"steps": n_mcmc, "walkers_per_gpu": W, "windowed": mcmc["windowed"], "brute": mcmc["brute"],
                 "note": "one step = every walker proposed once (two half-ensemble phases); chain kept in HBM"},
        "windowed": {"value": units / (ms_win / args.steps * 1e-3), "unit": "wt/s",
                     "ms_per_step": ms_win / args.steps, "speedup_vs_brute": ms_brute / ms_win,
                     "e2e": units / (dt_e2e_win / args.steps),
                     "note": "exact prefix-sum + in-transit-window kernel; effective wt/s, not a roofline figure"},
        "fp32": {"value": units / (ms_f32 / args.steps * 1e-3), "unit": "wt/s", "ms_per_step": ms_f32 / args.steps,
                 "max_rel_err_vs_fp64": f32_err,
                 "note": "NMB_F32: FP32 transit model (depth space) in the windowed pipeline; reported separately"},
        "strict": {"value": W * N * world * args.steps / (ms_strict * 1e-3), "unit": "wt/s", "ms_per_step": ms_strict / args.steps,
                   "max_rel_diff_default_vs_strict": strict_diff,
                   "note": "NMB_F64_STRICT, brute force: the hot occultation cases in the reference's own rounding sequence "
                           "(bit-identical to NumPy in most elements); the default NMB_F64 contracts multiply-adds and uses "
                           "<= 1.5-ulp quotients / roots there (lnprob within 1e-13 of the reference, tolerance 1e-10)"},
        "walker_split": split,
        "gp": {"value": units / (ms_gp * 1e-3), "unit": "wt/s", "ms_per_step": ms_gp, "ndim": 8,
               "note": "MonoLogProb: celerite RealTerm + jitter with the transit model as mean (2 free kernel "
                       "parameters + 6 transit parameters); sequential O(N) solve per walker, reported separately"},
        "roofline": {"bound": "fp64", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": dom["frac"], "traffic": dom["traffic"],
                     "peak_source": "DFMA microbenchmark in this process (MEASURED_PEAKS.json has no FP64 entry)",
                     "flops_convention": "SURVEY.md 8(d): 7/fine sample + 143 interior + 180 ingress + 5/timestamp + 40/walker",
                     "fine_samples": m, "interior": m_int, "edge": m_edge, "in_transit_timestamps": t_in,
                     "kernels": kernels,
                     "whole_step": {"flops": flops, "ms": ms_brute / args.steps, "achieved": flops / step_s / 1e12,
                                    "frac": flops / step_s / 1e12 / peak_tf},
                     "note": "dominant kernel = the longer of the two launches of a brute-force step; per-kernel "
                             "durations from CUDA events between the kernels (they do not overlap while measured)"},
    }
    if world == 1 and not args.no_cpu and not args.profile:
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        pool = mp.Pool(cores)
        rate, nwalk, dt = cpu_reference_rate(wl, 24.0, pool, cores)  # lands at 10-15 s of CPU work
        pool.close()
        line["cpu_baseline"] = {"value": rate, "unit": "wt/s", "cores": cores, "kind": "port",
                                "sample": "%d walker evaluations of the same workload (N=%d, S=%d) in %.1f s, "
                                          "NumPy oracle in a %d-process pool" % (nwalk, N, S, dt, cores)}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _emit(line):
    """Print the one JSON line on the real stdout (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries (NCCL's version banner, torchrun notices) write to fd 1; the driver expects exactly
    # one JSON line there, so everything else is diverted to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="k2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: no clock-sampling continuation, no cpu_baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
