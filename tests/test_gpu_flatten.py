"""Detrending row: k2flatten.ReduceNoise on the GPU against golden vectors
produced by the reference's own k2flatten.py (oracle/gen_golden_flatten.py)."""
import numpy as np
import pytest

from conftest import load_golden, priors_from_table

pytestmark = pytest.mark.gpu


def test_reduce_noise_matches_the_reference_output():
    from namaste_b200 import k2flatten
    g = load_golden("flatten_epic203311200")
    flat1 = k2flatten.ReduceNoise(g["lc"], winsize=4.5, stepsize=0.125)
    assert flat1.shape == g["lc"].shape
    assert np.array_equal(flat1[:, 0], g["lc"][:, 0]) and np.array_equal(flat1[:, 2], g["err1"])
    # the reference solves the same least-squares problems through an SVD of an ill-conditioned
    # Vandermonde matrix; the agreement is at the level of its own LAPACK noise
    assert np.abs(flat1[:, 1] - g["flat1"]).max() < 2e-9
    flat2 = k2flatten.ReduceNoise(flat1, winsize=4.5, stepsize=0.125)
    assert np.abs(flat2[:, 1] - g["flat2"]).max() < 4e-9
    gap = k2flatten.ReduceNoise(g["gap_lc"])          # defaults + a 2.5-day gap: shifted windows
    assert np.abs(gap[:, 1] - g["gap_flat"]).max() < 2e-9
    assert np.array_equal(gap[:, 1] == 0, g["gap_flat"] == 0)


def test_flatten_then_fit_reproduces_the_logged_known_answer():
    """Example.ipynb cell 45 -> 47 end to end on the GPU: flatten twice, cut 5 durations around the
    transit, evaluate the logged best-fit vector: nll 1170.0458538629673."""
    import namaste_b200 as nb
    from namaste_b200 import k2flatten
    g, kat = load_golden("flatten_epic203311200"), load_golden("kat_epic203311200")
    flat2 = k2flatten.ReduceNoise(k2flatten.ReduceNoise(g["lc"], winsize=4.5, stepsize=0.125), winsize=4.5, stepsize=0.125)
    lc = flat2[abs(flat2[:, 0] - 2121.02) < 5 * 0.5]
    assert len(lc) == 226
    nll = nb.MonoOnlyNegLogProb(kat["params"][0], lc, priors_from_table(kat["priors"]))
    assert abs(nll - kat["nll_logged"][0]) < 2e-7 * kat["nll_logged"][0]


def test_too_small_window_is_refused():
    from namaste_b200 import k2flatten
    g = load_golden("flatten_epic203311200")
    with pytest.raises(ValueError):
        k2flatten.ReduceNoise(g["lc"][:400], winsize=0.05, stepsize=0.04)
