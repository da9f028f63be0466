"""Edge cases and full-size properties of the likelihood path on the GPU: ragged candidate batches
(BASELINE config 4), generic supersampling factors, tiny inputs, and -- at the full sizes of BASELINE
configs 2 and 3 -- properties that do not need the oracle on every walker (brute force vs prefix-sum
mode, invariance under permutation of the ensemble) plus oracle spot checks."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, priors_from_table

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@pytest.fixture(scope="module")
def nb():
    import namaste_b200
    return namaste_b200


def test_ragged_candidate_batch(nb):
    from oracle import namaste_oracle as o
    g = load_golden("synth_tess_s10")
    base, pos, tab = g["lc"], g["walkers"][:12], g["priors"]
    mid = len(base) // 2
    cuts = [(mid - 700, mid + 650), (mid - 40, mid + 23), (0, len(base)), (mid - 3000, mid + 100), (mid - 1, mid + 1)]
    curves = [np.ascontiguousarray(base[a:b]) for a, b in cuts]
    batch = nb.LightCurve(curves, oversamp=10)
    assert batch.ncand == 5 and list(batch.npoints) == [b - a for a, b in cuts]
    params = np.stack([pos + 1e-6 * k for k in range(5)])            # [5, 12, 6]
    tabs = np.stack([tab] * 5)
    for mode in ("brute", "windowed"):
        got = batch.lnprob(params, tabs, mode=mode)
        assert got.shape == (5, 12)
        for k, lc in enumerate(curves):
            single = nb.LightCurve(lc, oversamp=10)
            one = single.lnprob(params[k], tab, mode=mode)
            assert np.array_equal(got[k], one, equal_nan=True), (mode, k)   # same bits as on its own
            single.close()
        with np.errstate(all="ignore"):
            ref = o.lnprob_ensemble(params[1], curves[1], priors_from_table(tab))
        ok = np.isfinite(ref)
        assert rel(got[1][ok], ref[ok]).max() < 1e-11, mode
    batch.close()


@pytest.mark.parametrize("S", [1, 3, 7, 16])
def test_generic_supersampling_factors(nb, S):
    """Only S = 1, 10, 11, 15 have compiled-in loops; everything else takes the run-time path."""
    from oracle import namaste_oracle as o
    g = load_golden("synth_k2_s11")
    lc, pos, pri = g["lc"][1700:2300], g["walkers"][:10], priors_from_table(g["priors"])
    with np.errstate(all="ignore"):
        ref = o.lnprob_ensemble(pos, lc, pri, S)
    staged = nb.LightCurve(lc, oversamp=S)
    for mode in ("brute", "windowed", "fused"):
        got = staged.lnprob(pos, pri, mode=mode)
        ok = np.isfinite(ref)
        assert rel(got[ok], ref[ok]).max() < 1e-11, (S, mode)
    staged.close()


@pytest.mark.parametrize("n", [2, 3, 63, 64, 65, 129])
def test_tiny_light_curves_and_single_walker(nb, n):
    from oracle import namaste_oracle as o
    g = load_golden("kat_epic203311200")
    lc, pri = g["lc"][100:100 + n], priors_from_table(g["priors"])
    th = g["params"][0].copy()
    th[0] = lc[n // 2, 0]
    staged = nb.LightCurve(lc)
    for mode in ("brute", "windowed", "fused"):
        got = staged.lnprob(th[None, :], pri, mode=mode)
        ref = o.mono_only_logprob(th, lc, pri)
        assert got.shape == (1,) and abs(got[0] - ref) <= 1e-11 * abs(ref), (n, mode)
    staged.close()


def _workload(name):
    sys.path.insert(0, ROOT)
    import bench
    return bench.make_workload(name)


def _special_walkers(wl, rng, n_random):
    """Walkers that exercise every branch of the path at full size: uniform in the prior box, from the tight
    ball, slow (vel ~ 0: nearly every timestamp in transit), grazing (b ~ 1 + p), a large planet (p > 1/2,
    outside the hot cases), a planet over the disc centre for a long time (b ~ 0), and outside the soft walls."""
    th = wl["theta"] if "theta" in wl else wl["pos"][-1]
    tab = wl["priors"] if wl["priors"].ndim == 2 else wl["priors"][0]
    box = tab[:, 0] + (tab[:, 1] - tab[:, 0]) * rng.uniform(0, 1, (n_random, 6))
    ball = th[None, :] * (1 + 1e-3 * rng.standard_normal((4, 6)))
    special = np.repeat(th[None, :], 6, axis=0)
    special[0, 2] = 1e-4                       # vel ~ 0: z ~ b for every sample of the light curve
    special[1, 1] = 1.0 + 0.999 * th[3]        # grazing: only the limb case, shallow
    special[2, 3] = 0.62                       # p > 1/2: the general path (cases i02 / i08 / i09 with a large planet)
    special[3, 1] = 1e-3                       # central transit: case i09 around mid-transit
    special[4, 0] = th[0] + 40.0; special[4, 3] = 0.3   # beyond the tcen and RpRs walls: soft-wall terms of 1e20
    special[5, 2] = 0.05 * th[2]; special[5, 1] = 0.9   # slow and nearly grazing: a long limb-dominated range
    return np.vstack([box, ball, special])


@pytest.mark.parametrize("name,nbox", [("k2", 6), ("kepler", 6), ("tess", 6)])
def test_full_size_properties(nb, name, nbox):
    """BASELINE configs 2 (N=4000, S=11, W=1024), 3 (N=65000, S=15, W=4096) and one GPU's share of 5 (N=500000,
    S=10, W=2048) at full size: properties that need no oracle on every walker, plus 16 oracle spot checks that
    cover the prior box, the tight ball and the special walkers of `_special_walkers`."""
    from oracle import namaste_oracle as o
    wl = _workload(name)
    staged = nb.LightCurve(wl["lc"], oversamp=wl["S"])
    pos, tab = wl["pos"].copy(), wl["priors"]
    spot = _special_walkers(wl, np.random.default_rng(77), nbox)
    pos[-len(spot):] = spot                         # the last 16 walkers of the ensemble are the spot checks
    brute = staged.lnprob(pos, tab, mode="brute")
    win = staged.lnprob(pos, tab, mode="windowed")
    assert brute.shape == (wl["W"],) and np.isfinite(brute).all()
    # every fine sample classified vs prefix sums + window: same likelihood
    assert rel(brute, win).max() < 1e-12
    # the ensemble is a set: permuting the walkers permutes the results, bit for bit
    perm = np.random.default_rng(1).permutation(wl["W"])
    assert np.array_equal(staged.lnprob(pos[perm], tab, mode="brute"), brute[perm])
    assert np.array_equal(staged.lnprob(pos[perm], tab, mode="windowed"), win[perm])
    # idempotence: evaluating again gives the same bits (workspace left clean)
    assert np.array_equal(staged.lnprob(pos, tab, mode="brute"), brute)
    pri = priors_from_table(tab)
    idx = list(range(wl["W"] - len(spot), wl["W"]))
    assert len(idx) >= 16
    with np.errstate(all="ignore"):
        for i in idx:
            ref = o.mono_only_logprob(pos[i], wl["lc"], pri, wl["S"])
            assert abs(brute[i] - ref) <= 1e-10 * abs(ref) and abs(win[i] - ref) <= 1e-10 * abs(ref), (name, i, brute[i], win[i], ref)
    staged.close()


def test_candidate_batch_256_distinct_light_curves(nb):
    """One GPU's share of BASELINE config 4: 256 candidates with their own light curves, transits, priors and
    100-walker ensembles, staged and evaluated as one batch.  Oracle spot checks on 48 (candidate, walker)
    pairs, and every candidate of a sample evaluated on its own gives the bits it has inside the batch."""
    from oracle import namaste_oracle as o
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.make_batch(0)
    rng = np.random.default_rng(5)
    pos = wl["pos"].copy()
    # some prior-box walkers in every candidate (the bench ensemble is a tight ball only)
    for c in range(wl["C"]):
        tab = wl["priors"][c]
        pos[c, :4] = tab[:, 0] + (tab[:, 1] - tab[:, 0]) * rng.uniform(0, 1, (4, 6))
    staged = nb.LightCurve(wl["curves"], oversamp=wl["S"])
    res = {mode: staged.lnprob(pos, wl["priors"], mode=mode) for mode in ("brute", "windowed")}
    assert res["brute"].shape == (256, 100) and np.isfinite(res["brute"]).all()
    assert rel(res["brute"], res["windowed"]).max() < 1e-12
    pairs = [(int(c), int(w)) for c, w in zip(rng.integers(0, 256, 48), rng.integers(0, 100, 48))]
    pairs[:8] = [(c, w % 4) for c, w in pairs[:8]]          # prior-box walkers among them
    with np.errstate(all="ignore"):
        for c, w in pairs:
            ref = o.mono_only_logprob(pos[c, w], wl["curves"][c], priors_from_table(wl["priors"][c]), wl["S"])
            for mode in ("brute", "windowed"):
                assert abs(res[mode][c, w] - ref) <= 1e-10 * abs(ref), (mode, c, w)
    for c in (0, 37, 255):
        single = nb.LightCurve(wl["curves"][c], oversamp=wl["S"])
        for mode in ("brute", "windowed"):
            assert np.array_equal(single.lnprob(pos[c], wl["priors"][c], mode=mode), res[mode][c]), (mode, c)
        single.close()
    staged.close()


_PATH_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import namaste_b200 as nb
from conftest import load_golden
g = load_golden("synth_k2_s11")
rng = np.random.default_rng(12)
base = g["walkers"][0]
out = {}
for W in (40, 320, 1000):
    pos = base[None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
    pos[: W // 16, 2] *= rng.uniform(0.03, 0.6, W // 16)      # slow walkers: long ranges, always queued
    pos[W // 16: W // 8, 1] = rng.uniform(1.0, 1.3, W // 8 - W // 16)   # grazing or never in transit
    lc = nb.LightCurve(g["lc"], oversamp=int(g["oversamp"]))
    for mode in ("brute", "windowed"):
        for dt in ("f64", "strict"):
            out["%d_%s_%s" % (W, mode, dt)] = lc.lnprob(pos, g["priors"], mode=mode, dtype=dt)
            s1, s2 = lc.last_kernels()
            out["k_%d_%s" % (W, mode)] = np.array([s1 + "+" + s2])
    lc.close()
np.savez(sys.argv[2], **out)
"""


def test_direct_and_mixed_stage2_give_the_queue_paths_bits(tmp_path):
    """window_direct_kernel (block per walker, the default for small ensembles) and eval_mixed_kernel (switched on
    with NMB_MIXED_* in a third process) against the pure queue pipeline (NMB_DIRECT=0, NMB_MIXED_MAXCW=0 in a second
    process): identical bits for every walker, mode and arithmetic, on ensembles that take each of the paths."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "paths.py"
    script.write_text(_PATH_SCRIPT)
    res = {}
    for tag, env in (("default", {}), ("queue", {"NMB_DIRECT": "0", "NMB_MIXED_MAXCW": "0"}),
                     ("mixed", {"NMB_MIXED_MINCW": "256", "NMB_MIXED_MAXCW": "1184"})):
        out = tmp_path / (tag + ".npz")
        subprocess.run([sys.executable, str(script), root, str(out)], check=True, env={**os.environ, **env})
        res[tag] = np.load(out)
    d, q, m = res["default"], res["queue"], res["mixed"]
    assert str(d["k_40_windowed"][0]) == "window_direct_kernel+window_direct_kernel(queue workers)"
    assert str(d["k_320_windowed"][0]) == "window_direct_kernel+window_direct_kernel(queue workers)"
    assert str(d["k_1000_windowed"][0]) == "window_ranges_kernel+eval_slots_kernel"
    assert str(d["k_320_brute"][0]) == "sweep_walkers_kernel+eval_slots_kernel"
    assert str(d["k_40_brute"][0]) == "sweep_walkers_kernel+eval_slots_kernel"
    assert str(m["k_1000_windowed"][0]) == "window_ranges_kernel+eval_mixed_kernel"
    assert str(m["k_320_brute"][0]) == "sweep_walkers_kernel+eval_mixed_kernel"
    for k in ("k_40_windowed", "k_320_windowed", "k_1000_windowed", "k_320_brute"):
        assert str(q[k][0]).endswith("eval_slots_kernel") and "direct" not in str(q[k][0])
    for key in m.files:
        if not key.startswith("k_"):
            assert np.array_equal(m[key], q[key], equal_nan=True), ("mixed", key)
    for key in d.files:
        if key.startswith("k_"):
            continue
        assert np.array_equal(d[key], q[key], equal_nan=True), key
        assert np.isfinite(d[key]).all()


@pytest.mark.parametrize("W", [3, 37, 41, 100, 132])
def test_sweep_tail_warp_layout_has_the_one_lane_bits(nb, W):
    """The brute-force sweep gives the walkers of a short tail warp (W mod 32 <= 16) several lanes each
    (csrc/eval_pipeline.cuh, `SUB`).  The same walkers inside an ensemble without such a tail (padded to a multiple
    of 32 with copies) take the one-lane-per-walker layout and must get the same bits; both agree with the oracle."""
    from oracle import namaste_oracle as o
    g = load_golden("synth_k2_s11")
    rng = np.random.default_rng(W)
    pos = g["walkers"][0][None, :] * (1 + 2e-3 * rng.standard_normal((W, 6)))
    pos[: max(1, W // 10), 2] *= 0.2                      # slow walkers: long in-transit ranges
    pos[W // 2, 1] = 1.2                                   # never in transit
    padded = np.vstack([pos, np.repeat(pos[:1], (-W) % 32 or 32, axis=0)])
    assert len(padded) % 32 == 0
    lc = nb.LightCurve(g["lc"], oversamp=int(g["oversamp"]))
    tail = lc.lnprob(pos, g["priors"], mode="brute")
    full = lc.lnprob(padded, g["priors"], mode="brute")[:W]
    assert np.array_equal(tail, full)
    pri = priors_from_table(g["priors"])
    with np.errstate(all="ignore"):
        ref = np.array([o.mono_only_logprob(p, g["lc"], pri, int(g["oversamp"])) for p in pos[:6]])
    assert rel(tail[:6], ref).max() < 1e-11
    lc.close()


def test_lnprob_host_with_pinned_buffers_and_out(nb):
    """nmb_lnprob_host copies page-locked caller buffers to and from the device directly and ordinary ones through
    its staging buffer; `out=` writes into the caller's array.  Same bits every way."""
    import torch
    g = load_golden("synth_k2_s11")
    lc = nb.LightCurve(g["lc"], oversamp=int(g["oversamp"]))
    pos = np.ascontiguousarray(g["walkers"][:48])
    ref, ref_ll = lc.lnprob(pos, g["priors"], return_loglike=True)
    pin_in = torch.from_numpy(pos.copy()).pin_memory()
    pin_out = torch.empty(48, dtype=torch.float64).pin_memory()
    got = lc.lnprob(pin_in.numpy(), g["priors"], out=pin_out.numpy())
    assert np.array_equal(got, ref) and np.array_equal(pin_out.numpy(), ref)
    got2, ll2 = lc.lnprob(pin_in.numpy(), g["priors"], return_loglike=True, out=pin_out.numpy())   # pinned lnprob, pageable loglike
    assert np.array_equal(got2, ref) and np.array_equal(ll2, ref_ll)
    plain = np.empty(48)
    assert np.array_equal(lc.lnprob(pos, g["priors"], out=plain), ref) and np.array_equal(plain, ref)
    with pytest.raises(ValueError):
        lc.lnprob(pos, g["priors"], out=np.empty(3))
    with pytest.raises(ValueError):
        lc.lnprob(pos, g["priors"], out=np.empty(48, dtype=np.float32))
    lc.close()
