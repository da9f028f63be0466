"""Posterior reduction on the device-resident chain against NumPy on the same
chain copied to the host (namaste.py:597-608 trimming + np.percentile summaries)."""
import numpy as np
import pytest

from conftest import load_golden, priors_from_table

pytestmark = pytest.mark.gpu


def test_device_summary_equals_numpy_on_the_host_chain():
    import namaste_b200 as nb
    g = load_golden("kat_epic203311200")
    lc, pri = g["lc"], priors_from_table(g["priors"])
    rng = np.random.default_rng(5)
    W, nsteps = 24, 240
    pos0 = g["params"][0][None, :] * (1 + 1e-4 * rng.standard_normal((W, 6)))
    pos0[5, 2] *= 0.2            # a walker started far away: likely to fail the filter
    s = nb.EnsembleSampler(W, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=99)
    s.run_mcmc(pos0, 1)
    s.run_mcmc(pos0, nsteps)     # namaste.py:592-594: T = 1 + nsteps
    got = s.summary(nsteps)
    samples, good, ncut = nb.trim_samples(s.chain, s.lnprobability, nsteps)
    assert got["ncut"] == ncut == 60
    assert np.array_equal(got["good_wlkrs"], good)
    assert got["nsamples"] == len(samples)
    want = np.percentile(samples, [15.865525393145707, 50.0, 84.13447460685429], axis=0)
    assert np.array_equal(got["percentiles"], want)
    # arbitrary percentiles, including the extremes
    q = [0.0, 2.5, 97.5, 100.0]
    assert np.array_equal(s.summary(nsteps, q)["percentiles"], np.percentile(samples, q, axis=0))


def test_derived_planet_parameters():
    import namaste_b200 as nb
    a, p = nb.VelToOrbit(np.array([3.0, 5.0]), 1.0, 1.0)
    assert np.allclose(p, 18226 / np.array([27.0, 125.0])) and np.allclose(a, 13.7375 / np.array([9.0, 25.0]))
    m = nb.PlanetRtoM(np.array([1.0, 2.0, 10.0, 20.0]))
    assert np.isclose(m[1], 2.69 * 2.0 ** 0.93) and 100.0 < m[2] < 300.0 and np.isclose(m[3], 333000 * (20 / 109.1) ** 1.5)
    rng = np.random.default_rng(1)
    samples = np.column_stack([np.zeros(50), np.full(50, 0.3), rng.normal(3.5, 0.1, 50), rng.normal(0.058, 0.002, 50)])
    out = nb.MonoFinalPars(samples, np.full(50, 2.02), np.full(50, 1.33), np.full(50, 0.15))
    assert out["Rp"] > 0 and out["P"] > 0 and out["Krv"] > 0 and 0 <= out["Prob_pl"] <= 1
