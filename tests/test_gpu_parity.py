"""GPU parity tests: the CUDA path (through the C ABI) against fixtures produced by the
reference's own code (tests/golden, see oracle/gen_golden.py) and against the oracle.

Tolerances: BASELINE.json north_star asks for 1e-10 relative on lnprob in FP64.  Two FP64
arithmetic variants are tested: dtype "strict" keeps the reference's operation order
(differences come only from libm and from per-element Bulirsch convergence) and is held to
2e-13; the default dtype "f64" contracts multiply-adds and uses <= 1.5-ulp quotients / roots
in the two hot occultation cases (measured <= 1e-13 on every fixture) and is held to 2e-12."""
import numpy as np
import pytest

from conftest import load_golden, priors_from_table

pytestmark = pytest.mark.gpu

LNPROB_RTOL = 1e-10   # north-star tolerance
LNPROB_RTOL_TIGHT = 2e-13  # what dtype="strict" actually achieves (regression guard)
LNPROB_RTOL_DEFAULT = 2e-12  # regression guard for the default dtype="f64" (relaxed arithmetic)
GUARD = {"f64": LNPROB_RTOL_DEFAULT, "strict": LNPROB_RTOL_TIGHT}
FP32_RTOL = 1e-5       # north-star tolerance of the FP32 variant (NMB_F32: FP32 geometry, FP64 elliptic integrals)
FP32_PURE_RTOL = 2e-4  # NMB_F32_PURE, measured 0.8e-5 .. 9.3e-5 (the lambda_d combination cancels ~60x)


@pytest.fixture(scope="module")
def nb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import namaste_b200
    return namaste_b200


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_occultquad_all_cases(nb):
    g = load_golden("occultquad_cases")
    z, p, g1, g2 = g["z"], g["p"], g["g1"], g["g2"]
    key = np.column_stack((p, g1, g2))
    worst = 0.0
    start = 0
    while start < len(z):
        stop = start
        while stop < len(z) and np.array_equal(key[stop], key[start]):
            stop += 1
        out = np.column_stack(nb.occultquad(z[start:stop], p[start], (g1[start], g2[start]), retall=True))
        ref_each = g["each"][start:stop]
        ref_batch = g["batch"][start:stop]
        nan_ref = np.isnan(ref_each)
        assert np.array_equal(np.isnan(out), nan_ref)
        ok = ~nan_ref
        d = np.abs(out[ok] - ref_each[ok])
        worst = max(worst, d.max())
        # per-element reference values: libm-level agreement
        assert d.max() < 5e-14, (p[start], g1[start], g2[start], d.max())
        # batch-coupled reference values (what the reference returns for the whole array)
        okb = ~np.isnan(ref_batch)
        assert np.abs(out[okb] - ref_batch[okb]).max() < 5e-14
        start = stop
    print("occultquad worst abs deviation:", worst)


@pytest.mark.parametrize("dtype", ["f64", "strict"])
@pytest.mark.parametrize("mode", ["brute", "windowed", "fused"])
def test_logged_known_answers(nb, mode, dtype):
    if mode == "fused" and dtype != "strict":
        pytest.skip("the single-launch diagnostic kernel has one arithmetic (strict)")
    g = load_golden("kat_epic203311200")
    pri = priors_from_table(g["priors"])
    model = nb.MonotransitModel(*g["params"][0])
    for th, nll, ll, lp in zip(g["params"], g["nll_logged"], g["ll"], g["lp"]):
        got = nb.MonoOnlyNegLogProb(th, g["lc"], pri, model, mode=mode, dtype=dtype)
        assert abs(got - nll) / nll < 5e-11          # vs the value the reference logged in 2018
        assert abs(got + (ll + lp)) / nll < GUARD[dtype]  # vs the reference run here on the same lc
    assert np.array_equal(model.get_parameter_vector(), g["params"][-1])


@pytest.mark.parametrize("name", ["config1_epic203311200", "synth_k2_s11", "synth_kepler_s15",
                                  "synth_tess_s10", "synth_tess_s1"])
@pytest.mark.parametrize("dtype", ["f64", "strict"])
@pytest.mark.parametrize("mode", ["brute", "windowed", "fused"])
def test_lnprob_fixtures(nb, name, mode, dtype):
    if mode == "fused" and dtype != "strict":
        pytest.skip("the single-launch diagnostic kernel has one arithmetic (strict)")
    g = load_golden(name)
    S = int(g["oversamp"]) if "oversamp" in g.files else 10
    lc = nb.LightCurve(g["lc"], oversamp=S)
    lnp, ll = lc.lnprob(g["walkers"], g["priors"], mode=mode, return_loglike=True, dtype=dtype)
    ref_ll, ref_lp = g["ll"], g["lp"]
    finite = np.isfinite(ref_ll + ref_lp)
    assert np.array_equal(np.isfinite(lnp), finite)
    r_ll = rel(ll[finite], ref_ll[finite]).max()
    r_lnp = rel(lnp[finite], (ref_ll + ref_lp)[finite]).max()
    print(name, mode, dtype, "max rel ll", r_ll, "lnprob", r_lnp)
    assert r_ll < LNPROB_RTOL and r_lnp < LNPROB_RTOL
    assert r_ll < GUARD[dtype] and r_lnp < GUARD[dtype]
    lc.close()


def test_model_values(nb):
    g = load_golden("kat_epic203311200")
    lc = nb.LightCurve(g["lc"], oversamp=10)
    m = lc.model(g["params"])
    assert np.abs(m - g["model"]).max() < 4e-16
    for name in ["synth_k2_s11", "synth_kepler_s15", "synth_tess_s1"]:
        s = load_golden(name)
        lcs = nb.LightCurve(s["lc"], oversamp=int(s["oversamp"]))
        m = lcs.model(s["walkers"][:4])
        assert np.abs(m - s["model4"]).max() < 4e-16, name
    mm = nb.MonotransitModel(*g["params"][0])
    assert np.abs(mm.get_value(g["lc"][:, 0]) - g["model"][0]).max() < 4e-16


def test_lnprior(nb):
    g = load_golden("synth_k2_s11")
    pri = priors_from_table(g["priors"])
    got = nb.MonoLnPrior(g["walkers"], pri)
    assert rel(got, g["lp"]).max() < 1e-13
    assert abs(nb.MonoLnPrior(g["walkers"][0], pri) - g["lp"][0]) <= 1e-13 * abs(g["lp"][0])
    assert nb.MonoLnPrior(g["walkers"][0], {}) == 0.0


def test_brute_vs_windowed_vs_oracle_random_box(nb):
    """Seeded random parameter boxes on a small synthetic light curve, checked against the
    oracle evaluated on the same inputs (sizes the oracle finishes in seconds)."""
    from oracle import namaste_oracle as o
    rng = np.random.default_rng(5)
    n, cad = 700, 0.0204318
    t = 2000.0 + cad * np.arange(n)
    t = t[(np.arange(n) % 97) != 13]            # a few single-cadence gaps
    theta = np.array([t[len(t) // 2] + 0.3 * cad, 0.3, 4.0, 0.08, 0.4, 0.25])
    flux = o.get_value(theta, t) + 1e-4 * rng.standard_normal(len(t))
    lc = np.column_stack((t, flux, np.full(len(t), 1e-4)))
    W = 64
    pos = np.column_stack([rng.uniform(t[0] - 0.5, t[-1] + 0.5, W), rng.uniform(-0.2, 1.4, W),
                           rng.uniform(0.0, 8.0, W), rng.uniform(-0.1, 0.7, W),
                           rng.uniform(-0.2, 1.2, W), rng.uniform(-0.3, 1.0, W)])
    g = load_golden("kat_epic203311200")
    tab = g["priors"].copy()
    tab[0, :2] = [theta[0] - 0.2, theta[0] + 0.2]
    pri = priors_from_table(tab)
    ref = o.lnprob_ensemble(pos, lc, pri)
    staged = nb.LightCurve(lc, oversamp=10)
    for mode in ("brute", "windowed", "fused"):
        for dtype in ("f64", "strict"):
            got = staged.lnprob(pos, pri, mode=mode, dtype=dtype)
            assert rel(got, ref).max() < GUARD[dtype], (mode, dtype)
    assert np.allclose(nb.MonoOnlyLogProb(pos, lc, pri), ref, rtol=LNPROB_RTOL_DEFAULT, atol=0)


def test_errors_are_loud(nb):
    g = load_golden("kat_epic203311200")
    lc = g["lc"].copy()
    lc[5, 0] = lc[4, 0] + 1e-6  # violates the supersampling precondition
    with pytest.raises(ValueError):
        nb.LightCurve(lc)
    with pytest.raises(ValueError):
        nb.LightCurve(g["lc"][:1])
    with pytest.raises(ValueError):
        nb.LightCurve(g["lc"], oversamp=0)
    nan_par = g["params"][0].copy()
    nan_par[4] = np.nan
    assert np.isnan(nb.MonoOnlyLogProb(nan_par, g["lc"], priors_from_table(g["priors"])))


@pytest.mark.parametrize("name", ["config1_epic203311200", "synth_k2_s11", "synth_kepler_s15", "synth_tess_s10"])
@pytest.mark.parametrize("mode", ["brute", "windowed"])
def test_fp32_variant(nb, name, mode):
    """NMB_F32 holds the north star's 1e-5 relative on lnprob on every fixture: FP32 where that tolerance allows it
    (the geometry's root and the case selection), FP64 for K, E, Pi and lambda_d.  The all-FP32 occultation
    (NMB_F32_PURE) is kept as a reported variant: 1e-5 .. 1e-4 on high-S/N light curves, guard 2e-4."""
    g = load_golden(name)
    S = int(g["oversamp"]) if "oversamp" in g.files else 10
    lc = nb.LightCurve(g["lc"], oversamp=S)
    ref = g["ll"] + g["lp"]
    finite = np.isfinite(ref)
    for dtype, tol in (("f32", FP32_RTOL), ("f32_pure", FP32_PURE_RTOL)):
        lnp, ll = lc.lnprob(g["walkers"], g["priors"], mode=mode, return_loglike=True, dtype=dtype)
        r = rel(lnp[finite], ref[finite])
        rl = rel(ll[finite], g["ll"][finite])
        print(name, mode, dtype, "max rel lnprob", r.max(), "ll", rl.max())
        assert r.max() < tol and rl.max() < tol, dtype
    lc.close()


def test_degenerate_limb_darkening_matches_oracle(nb):
    """1 - LD1/3 - LD2/6 == 0 (or non-finite LD) makes the reference's flux NaN / inf everywhere, also
    out of transit; the sweep must not take its 'model == 1' shortcut for such walkers."""
    from oracle import namaste_oracle as o
    g = load_golden("synth_k2_s11")
    lc, S = g["lc"][:600], int(g["oversamp"])
    pos = np.repeat(g["walkers"][:1], 6, axis=0)
    pos[:, 0] = lc[300, 0] + 0.003
    pos[1, 4:] = [3.0, 0.0]            # four_omega == 0
    pos[2, 4:] = [0.0, 6.0]            # four_omega == 0 through LD2
    pos[3, 4] = np.inf
    pos[4, 5] = -np.inf
    pos[5, 4:] = [1.5, 3.0]            # four_omega == 0, both non-zero
    pri = priors_from_table(g["priors"])
    with np.errstate(all="ignore"):
        ref = o.lnprob_ensemble(pos, lc, pri, S)
    staged = nb.LightCurve(lc, oversamp=S)
    for mode in ("brute", "windowed", "fused"):
        got = staged.lnprob(pos, pri, mode=mode)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (mode, got, ref)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isinf(got), np.isinf(ref)), (mode, got, ref)
        assert rel(got[ok], ref[ok]).max() < LNPROB_RTOL_DEFAULT, mode
    staged.close()


def test_results_do_not_depend_on_the_launch_shape(nb):
    """Partial sums are tied to per-light-curve granules and per-walker slots, so a walker's lnprob has
    the same bits whatever ensemble it is evaluated in (what makes walker-split chains identical)."""
    g = load_golden("synth_kepler_s15")
    lc = nb.LightCurve(g["lc"], oversamp=int(g["oversamp"]))
    pos = np.tile(g["walkers"], (12, 1))[:280]
    for mode in ("brute", "windowed"):
        full = lc.lnprob(pos, g["priors"], mode=mode)
        for lo, hi in ((0, 1), (3, 40), (100, 280), (17, 147)):
            part = lc.lnprob(pos[lo:hi], g["priors"], mode=mode)
            assert np.array_equal(part, full[lo:hi], equal_nan=True), (mode, lo, hi)
    lc.close()


def test_literal_callable_finds_its_staged_light_curve(nb):
    """MonoOnlyLogProb(params, lc ndarray, priors, model) -- the reference's call (namaste.py:590, emcee args=) --
    stages `lc` once and finds it again by address + content; an in-place edit is seen, and a read-only array is
    keyed by identity alone."""
    from namaste_b200 import namaste as host
    g = load_golden("kat_epic203311200")
    pri = priors_from_table(g["priors"])
    lc = g["lc"].copy()
    host.clear_cache()
    a = nb.MonoOnlyLogProb(g["params"], lc, pri)
    assert len(host._LC_CACHE) == 1
    b = nb.MonoOnlyLogProb(g["params"], lc, pri)
    assert len(host._LC_CACHE) == 1 and np.array_equal(a, b)
    lc[100, 1] *= 1.0 + 1e-6                                   # same address, one flux value changed
    c = nb.MonoOnlyLogProb(g["params"], lc, pri)
    assert len(host._LC_CACHE) == 2 and not np.array_equal(a, c)
    frozen = g["lc"].copy()
    frozen.setflags(write=False)
    d = nb.MonoOnlyLogProb(g["params"], frozen, pri)
    assert np.array_equal(d, a) and len(host._fingerprint(frozen)) == 4
    host.clear_cache()
