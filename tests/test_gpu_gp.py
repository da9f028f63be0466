"""GP objective MonoLogProb on the GPU against the CPU oracle of the celerite
recurrence (oracle/gp_oracle.py; parity with celerite itself is unpinned, the package is absent)."""
import numpy as np
import pytest

from conftest import load_golden, priors_from_table

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _gp_priors(kernel_rows, table):
    from collections import OrderedDict
    pri = OrderedDict(kernel_rows)
    for k, v in priors_from_table(table).items():
        pri[k] = v
    return pri


def test_real_kernel_ensemble_matches_oracle():
    import namaste_b200 as nb
    from oracle import gp_oracle as go
    g = load_golden("kat_epic203311200")
    lc = g["lc"]
    rng = np.random.default_rng(21)
    W = 24
    mean = g["params"][0][None, :] * (1 + 1e-3 * rng.standard_normal((W, 6)))
    mean[-1, 2] = 0.0    # vel = 0: every sample in transit
    mean[-2, 1] = 1.3    # never transits
    kpar = np.column_stack([rng.uniform(-18, -10, W), rng.uniform(-2.5, 3.5, W)])
    pos = np.hstack([kpar, mean])
    kern = nb.RealTerm(log_a=-14.0, log_c=-1.0986122886681098, bounds={"log_a": (-20, 0.0), "log_c": (-3, 4.5)}) + \
        nb.JitterTerm(-9.8, bounds=dict(log_sigma=(-10.05, -9.55)))
    kern.freeze_parameter("terms[1]:log_sigma")
    gp = nb.GP(kernel=kern, mean=nb.MonotransitModel(*mean[0]), fit_mean=True)
    assert gp.get_parameter_names() == ("kernel:terms[0]:log_a", "kernel:terms[0]:log_c", "mean:tcen", "mean:b",
                                        "mean:vel", "mean:RpRs", "mean:LD1", "mean:LD2")
    pri = _gp_priors([("kernel:terms[0]:log_a", [-20, 0.0]), ("kernel:terms[0]:log_c", [-3, 4.5])], g["priors"])
    got = nb.MonoLogProb(pos, lc, pri, gp)
    ref = np.array([go.mono_logprob(p, lc, pri, "Real", [0, 0, -9.8], [True, True, False]) for p in pos])
    assert rel(got, ref).max() < RTOL
    # dense linear algebra on a few walkers (independent of the recurrence)
    dense = np.array([go.mono_logprob(p, lc, pri, "Real", [0, 0, -9.8], [True, True, False], dense=True) for p in pos[:4]])
    assert rel(got[:4], dense).max() < 1e-8
    # scalar contract: float out, vector left on the gp object (namaste.py:1335)
    one = nb.MonoLogProb(pos[3], lc, pri, gp)
    assert isinstance(one, float) and one == got[3]
    assert np.array_equal(gp.get_parameter_vector(), pos[3])
    assert nb.MonoNegLogProb(pos[3], lc, pri, gp) == -one
    # celerite's compute / log_likelihood pair
    gp.compute(lc[:, 0], lc[:, 2])
    ll = gp.log_likelihood(lc[:, 1])
    ll_ref = go.mono_logprob(pos[3], lc, {}, "Real", [0, 0, -9.8], [True, True, False])
    assert abs(ll - ll_ref) < RTOL * abs(ll_ref)


def test_rotation_kernel_with_free_jitter_matches_oracle():
    import namaste_b200 as nb
    from oracle import gp_oracle as go
    g = load_golden("synth_k2_s11")
    S = int(g["oversamp"])
    lc = g["lc"][1400:2600]
    rng = np.random.default_rng(22)
    W = 16
    mean = g["walkers"][:W].copy()
    kfull = np.array([-12.842600381780318, 0.488322328246102, 0.9162907318741551, 0.0, -8.996145323242303])
    free = [True, True, True, False, True]      # log_factor frozen (namaste.py:306), jitter left free here
    kpar = kfull[free][None, :] + 0.3 * rng.standard_normal((W, 4))
    pos = np.hstack([kpar, mean])
    kern = nb.RotationTerm(*kfull[:4]) + nb.JitterTerm(kfull[4])
    kern.freeze_parameter("terms[0]:log_factor")
    model = nb.MonotransitModel(*mean[0])
    model.oversamp = S
    gp = nb.GP(kernel=kern, mean=model, fit_mean=True)
    assert len(gp.get_parameter_vector()) == 10
    got = nb.MonoLogProb(pos, lc, None, gp)
    with np.errstate(all="ignore"):
        ref = np.array([go.mono_logprob(p, lc, {}, "quasi", kfull, free, oversamp=S) for p in pos])
    ok = np.isfinite(ref)
    assert ok.sum() >= W - 2 and np.array_equal(np.isnan(got), np.isnan(ref))
    assert rel(got[ok], ref[ok]).max() < RTOL


def test_other_kernels_are_refused():
    import namaste_b200 as nb
    with pytest.raises(TypeError):
        nb.GP(kernel=nb.RealTerm(-14.0, 0.0) + nb.RealTerm(-12.0, 1.0), mean=nb.MonotransitModel(0, 0, 1, 0.1, 0.3, 0.3))


def test_gp_sampler_chain_matches_cpu_oracle():
    """RunMCMC's default objective (namaste.py:583): EnsembleSampler over MonoLogProb, on the device,
    against the CPU stretch move driven by the GP oracle with the same Philox stream."""
    import namaste_b200 as nb
    from oracle import gp_oracle as go, stretch_oracle as so
    g = load_golden("kat_epic203311200")
    lc = g["lc"]
    rng = np.random.default_rng(33)
    W, nsteps, seed = 20, 12, 424242
    vec0 = np.concatenate([[-15.0, -1.0], g["params"][0]])
    pos0 = vec0[None, :] * (1 + 1e-4 * rng.standard_normal((W, 8)))
    kern = nb.RealTerm(log_a=-15.0, log_c=-1.0) + nb.JitterTerm(-9.8)
    kern.freeze_parameter("terms[1]:log_sigma")
    gp = nb.GP(kernel=kern, mean=nb.MonotransitModel(*g["params"][0]), fit_mean=True)
    pri = _gp_priors([("kernel:terms[0]:log_a", [-20, 0.0]), ("kernel:terms[0]:log_c", [-3, 4.5])], g["priors"])
    s = nb.EnsembleSampler(W, 8, nb.MonoLogProb, args=(lc, pri, gp), seed=seed)
    pos, lnp, _ = s.run_mcmc(pos0, nsteps)
    ref_chain, ref_ln, ref_acc = so.run(
        lambda p: np.array([go.mono_logprob(x, lc, pri, "Real", [0, 0, -9.8], [True, True, False]) for x in p]),
        pos0, nsteps, seed)
    assert s.chain.shape == (W, nsteps, 8)
    assert np.array_equal(s.chain, ref_chain)
    assert np.allclose(s.lnprobability, ref_ln, rtol=1e-10, atol=0)
    assert np.array_equal(s.naccepted, ref_acc)
    assert 0.05 < s.acceptance_fraction.mean() < 0.95
    with pytest.raises(ValueError):
        nb.EnsembleSampler(W, 6, nb.MonoLogProb, args=(lc, pri, gp))
