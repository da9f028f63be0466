"""Single-point outlier cut before detrending (namaste/namaste.py:1456-1467):
the host layer's ``AnomCutDiff`` / ``nonan`` against golden masks produced by the reference's own
function (oracle/gen_golden_anomcut.py), against the oracle's restatement, and - in the build container -
against the reference's source run on random series."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import _refimport, namaste_oracle as o
from namaste_b200.namaste import AnomCutDiff, nonan


@pytest.mark.parametrize("series", ["epic", "syn"])
def test_golden_masks(series):
    g = load_golden("anomcut")
    flux = g[series]
    for key, thresh in (("375", 3.75), ("420", 4.2)):
        want = g["%s_%s" % (series, key)]
        assert np.array_equal(AnomCutDiff(flux, thresh), want)
        assert np.array_equal(o.anom_cut_diff(flux.copy(), thresh), want)
    assert np.array_equal(AnomCutDiff(flux), g[series + "_420"])          # default threshold
    assert (~g["syn_420"]).sum() > 40 and (~g["epic_375"]).sum() == 2     # the fixtures do cut something
    assert flux.tolist() == g[series].tolist()                            # inputs are not modified


def test_what_is_cut():
    f = np.ones(50)
    f[10] += 1.0                      # spike: cut
    f[20:22] += 1.0                   # two in a row: kept
    f[30:] += np.linspace(0, 1, 20)   # ramp: kept
    f += 1e-3 * np.sin(np.arange(50))
    m = AnomCutDiff(f)
    assert not m[10] and m[[20, 21]].all() and m[30:].all() and m[0] and m[-1] and (~m).sum() == 1


def test_nonan():
    lc = np.ones((6, 3))
    lc[1, 1] = np.nan
    lc[4, 2] = np.nan
    assert nonan(lc).tolist() == [True, False, True, True, False, True]


@pytest.mark.skipif(not _refimport.available(), reason="reference checkout not mounted (GPU box)")
def test_against_reference_source():
    ref = _refimport.namaste_function("AnomCutDiff")
    rng = np.random.default_rng(9)
    for n in (3, 4, 17, 1000):
        for _ in range(20):
            f = 1 + 1e-3 * rng.standard_normal(n)
            f[rng.integers(0, n, max(1, n // 10))] += rng.standard_normal(max(1, n // 10)) * 1e-2
            t = float(rng.uniform(2, 6))
            assert np.array_equal(AnomCutDiff(f, t), ref(f.copy(), t)), (n, t)
