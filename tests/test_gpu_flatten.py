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


def _candidates(g, k):
    """k light curves cut from the example (different spans and lengths, one with a gap)."""
    lc = g["lc"]
    out = [lc, g["gap_lc"]]
    rng = np.random.default_rng(4)
    while len(out) < k:
        a = int(rng.integers(0, len(lc) - 1200))
        b = a + int(rng.integers(900, 1200))
        piece = lc[a:b].copy()
        piece[:, 1] *= 1.0 + 1e-4 * rng.standard_normal(len(piece))
        out.append(piece)
    return out


def test_batch_detrending_equals_one_by_one_and_the_reference():
    """nmb_flatten_batch: ragged candidates in one call give, candidate by candidate, the bits of the single call,
    and the example light curve inside the batch still matches the reference's own output."""
    from namaste_b200 import k2flatten
    g = load_golden("flatten_epic203311200")
    cands = _candidates(g, 12)
    batch = k2flatten.ReduceNoiseBatch(cands, winsize=4.5, stepsize=0.125)
    assert len(batch) == len(cands)
    for lc, flat in zip(cands, batch):
        one = k2flatten.ReduceNoise(lc, winsize=4.5, stepsize=0.125)
        assert flat.shape == lc.shape and np.array_equal(flat, one)
    assert np.abs(batch[0][:, 1] - g["flat1"]).max() < 2e-9
    gap = k2flatten.ReduceNoiseBatch([g["gap_lc"], g["lc"][:1500]])[0]      # function defaults, gap branch
    assert np.abs(gap[:, 1] - g["gap_flat"]).max() < 2e-9 and np.array_equal(gap[:, 1] == 0, g["gap_flat"] == 0)


def test_batch_detrending_against_the_cpu_oracle():
    from namaste_b200 import k2flatten
    from oracle import flatten_oracle as fo
    g = load_golden("flatten_epic203311200")
    cands = _candidates(g, 5)[2:]
    batch = k2flatten.ReduceNoiseBatch(cands, winsize=3.0, stepsize=0.25, polydegree=2, niter=10, sigmaclip=2.5)
    for lc, flat in zip(cands, batch):
        ref = fo.reduce_noise(lc, winsize=3.0, stepsize=0.25, polydegree=2, niter=10, sigmaclip=2.5)
        assert np.abs(flat[:, 1] - ref[:, 1]).max() < 2e-9
        assert np.array_equal(flat[:, 1] == 0, ref[:, 1] == 0)


def test_batch_outlier_cut_and_window_mask():
    """AnomCutDiff on the device against masks produced by the reference's own function (tests/golden/anomcut.npz),
    and the RunMCMC window mask (both the reference's one-sided quirk and the two-sided form)."""
    import namaste_b200 as nb
    from namaste_b200 import k2flatten
    a = load_golden("anomcut")
    masks = k2flatten.AnomCutDiffBatch([a["epic"], a["syn"], a["epic"][:777]], 3.75)
    assert np.array_equal(masks[0], a["epic_375"]) and np.array_equal(masks[1], a["syn_375"])
    assert np.array_equal(masks[2], nb.AnomCutDiff(a["epic"][:777], 3.75))
    masks = k2flatten.AnomCutDiffBatch([a["epic"], a["syn"]])
    assert np.array_equal(masks[0], a["epic_420"]) and np.array_equal(masks[1], a["syn_420"])
    g = load_golden("flatten_epic203311200")
    t1, t2 = g["lc"][:, 0], g["gap_lc"][:, 0]
    tc, wd = np.array([2121.02, t2[700]]), np.array([3.2, 1.1])
    for two in (False, True):
        got = k2flatten.window_mask_batch([t1, t2], tc, wd, two_sided=two)
        for t, c, w, m in zip((t1, t2), tc, wd, got):
            want = (np.abs(t - c) < w) if two else abs(t - c < w).astype(bool)
            assert np.array_equal(m, want)
