"""Host-side rows of the path (a11 prior construction, a13 post-run trimming).

a11 is pinned against the reference's own output: the priors dict it logged for EPIC 203311200
(Example.ipynb cell 47 / namaste_runlog.log), reproduced here from the PDFs the reference saved and
its example light curve (fixture made by oracle/gen_golden_priors.py)."""
import numpy as np
import pytest

from conftest import load_golden


@pytest.fixture(scope="module")
def nm():
    from namaste_b200 import namaste
    return namaste


def test_prior_rows_match_the_reference_log_exactly(nm):
    g = load_golden("priors_epic203311200")
    lc = np.column_stack((g["t"], np.ones_like(g["t"]), np.ones_like(g["t"])))
    tcen, tdur = float(g["tcen"]), float(g["tdur"])
    maxvel, uerr, derr, minp = nm.calcmaxvel(lc, g["sdenss"], tcen, tdur)
    assert maxvel == 3.5736491649686775 and uerr == 0.3833318468962794   # logged 'normlim' mu / sigma
    pri = nm.BuildMeanPriors(lc, tcen, tdur, g["sdenss"], g["LD1s"], g["LD2s"])
    assert list(pri.keys()) == ["mean:tcen", "mean:b", "mean:vel", "mean:RpRs", "mean:LD1", "mean:LD2"]
    from namaste_b200.core import prior_table
    tab = prior_table(pri)
    exp = g["expected"]
    # tcen bounds were logged to two decimals (2120.87, 2121.17); everything else is full precision
    assert np.allclose(tab[0, :2], exp[0, :2], rtol=0, atol=1e-9)
    assert np.array_equal(tab[1:], exp[1:])
    assert pri["mean:vel"][2] == "normlim" and pri["mean:LD1"][2] == "gaussian"


def test_calcminp_branches(nm):
    t = np.concatenate([np.arange(0, 10, 0.02), np.arange(30, 40, 0.02)])
    lc = np.column_stack((t, t * 0 + 1, t * 0 + 1))
    # transit in the first chunk: folded distance |t - tcen| first jumps by > tdur at the gap
    d = abs(t - 5.0)
    j = np.where(np.diff(d) > 0.5)[0][0]
    assert nm.calcminp(lc, 5.0, 0.5) == d[j] + 0.5 * 0.33
    # no gap wider than tdur: farthest point
    t2 = np.arange(0, 10, 0.02)
    lc2 = np.column_stack((t2, t2 * 0 + 1, t2 * 0 + 1))
    assert nm.calcminp(lc2, 2.0, 0.5) == np.max(abs(t2 - 2.0)) + 0.5 * 0.33


def test_window_mask_reproduces_the_reference_quirk(nm):
    g = load_golden("config1_epic203311200")
    kat = load_golden("kat_epic203311200")
    th = kat["params"][0]
    # config 1's light curve IS the reference's one-sided window applied to the flattened curve
    t = g["lc"][:, 0]
    m = nm.transit_window_mask(t, th[0], th[2], th[1], th[3])
    assert m.all() and len(t) == 2828
    width = 6 * nm.CalcTdur(th[2], th[1], th[3])
    assert t.max() - th[0] < width and (th[0] - t.min()) > width        # nothing cut before the transit
    two = nm.transit_window_mask(t, th[0], th[2], th[1], th[3], two_sided=True)
    assert two.sum() == 298


def test_initial_walkers_follow_namaste_nan_rule(nm):
    g = load_golden("priors_epic203311200")
    pdfs = g["pdfs"].copy()
    pdfs[::7, 2] = np.nan
    rng = np.random.RandomState(5)
    pos = nm.initial_walkers(pdfs, 100, rng)
    assert pos.shape == (100, 6) and np.isfinite(pos).all()
    assert (pos[:, 2] == 0.0).sum() > 0           # NaN -> nanmedian(isnan(column)) = 0.0 (quirk Q5)
    rows = np.random.RandomState(5).choice(len(pdfs), 100, replace=False)
    ok = ~np.isnan(pdfs[rows])
    assert np.array_equal(pos[ok], pdfs[rows][ok])


def test_trim_samples_matches_walkerwise_restatement(nm):
    from oracle import namaste_oracle as o
    rng = np.random.default_rng(11)
    W, T, nsteps = 16, 41, 40
    chain = rng.normal(size=(W, T, 6))
    lnp = rng.normal(loc=-100, scale=2, size=(W, T))
    lnp[3] -= 50.0      # a stuck walker
    lnp[9, :15] -= 1e6  # bad only during burn-in: kept
    samples, good, ncut = nm.trim_samples(chain, lnp, nsteps)
    ref, good_ref, ncut_ref = o.trim_samples(chain, lnp, nsteps)
    assert ncut == ncut_ref == 10
    assert np.array_equal(good, good_ref) and not good[3] and good[9]
    assert samples.shape == (good.sum() * (T - ncut), 7)
    assert np.array_equal(samples, ref)
    # long runs: the cut saturates at 3000 steps
    big = rng.normal(size=(4, 3005))
    s2, g2, n2 = nm.trim_samples(np.zeros((4, 3005, 6)), big, 20000)
    assert n2 == 3000 and s2.shape == (g2.sum() * 5, 7)
