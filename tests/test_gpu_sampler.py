"""Device ensemble sampler vs the CPU stretch-move oracle (same Philox stream, NumPy likelihood)."""
import numpy as np
import pytest

from conftest import load_golden, priors_from_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import torch
    assert torch.cuda.is_available()
    import namaste_b200
    return namaste_b200


def _problem():
    g = load_golden("kat_epic203311200")
    lc, pri = g["lc"], priors_from_table(g["priors"])
    rng = np.random.default_rng(42)
    W = 24
    pos0 = g["params"][0][None, :] * (1 + 1e-4 * rng.standard_normal((W, 6)))
    return lc, pri, pos0


@pytest.mark.parametrize("mode", ["windowed", "brute"])
def test_chain_matches_cpu_oracle_draw_by_draw(nb, mode):
    from oracle import namaste_oracle as o, stretch_oracle as so
    lc, pri, pos0 = _problem()
    nsteps, seed = 25, 20261019
    ref_chain, ref_ln, ref_acc = so.run(lambda p: o.lnprob_ensemble(p, lc, pri), pos0, nsteps, seed)
    s = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri, nb.MonotransitModel(*pos0[0])), mode=mode, seed=seed)
    pos, lnp, _ = s.run_mcmc(pos0, nsteps)
    assert s.chain.shape == (24, nsteps, 6) and s.lnprobability.shape == (24, nsteps)
    # proposals are bit-identical as long as the accept decisions agree; they can only differ when
    # |lnpdiff - ln u| ~ 1e-12, which a 25-step chain does not hit
    assert np.array_equal(s.chain, ref_chain)
    assert np.allclose(s.lnprobability, ref_ln, rtol=1e-11, atol=0)
    assert np.array_equal(s.naccepted, ref_acc)
    assert np.array_equal(pos, ref_chain[:, -1, :]) and np.allclose(lnp, ref_ln[:, -1], rtol=1e-11)
    assert 0.1 < s.acceptance_fraction.mean() < 0.9


def test_chain_with_queued_walkers_matches_cpu_oracle(nb):
    """Slow walkers in the initial ensemble (long in-transit ranges) leave window_direct_kernel's direct blocks for
    the slot queue, which the queue workers at the end of the same launch drain: accept / reject of those walkers
    runs in the workers.  The chain must still be the CPU oracle's, draw by draw."""
    from oracle import namaste_oracle as o, stretch_oracle as so
    lc, pri, pos0 = _problem()
    pos0 = pos0.copy()
    pos0[::5, 2] *= 0.04            # five walkers ~25x slower: a few hundred timestamps in transit
    nsteps, seed = 12, 77
    with np.errstate(all="ignore"):
        ref_chain, ref_ln, ref_acc = so.run(lambda p: o.lnprob_ensemble(p, lc, pri), pos0, nsteps, seed)
    s = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri, nb.MonotransitModel(*pos0[0])), seed=seed)
    s.run_mcmc(pos0, nsteps)
    assert np.array_equal(s.chain, ref_chain)
    assert np.allclose(s.lnprobability, ref_ln, rtol=1e-11, atol=0)
    assert np.array_equal(s.naccepted, ref_acc)


def test_run_mcmc_appends_and_reset(nb):
    lc, pri, pos0 = _problem()
    s = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=7)
    s.run_mcmc(pos0, 1)
    s.run_mcmc(pos0, 9)          # namaste.py:592-594 calls run_mcmc twice from the same start: T = 1 + nsteps
    assert s.chain.shape == (24, 10, 6) and s.iterations == 10
    one = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=7)
    one.run_mcmc(pos0, 4)
    one.run_mcmc(None, 6)        # continue from the last state
    assert one.chain.shape == (24, 10, 6)
    straight = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=7)
    straight.run_mcmc(pos0, 10)
    assert np.array_equal(one.chain, straight.chain)     # counter-based RNG: splitting a run changes nothing
    assert one.flatchain.shape == (240, 6) and one.flatlnprobability.shape == (240,)
    one.reset()
    assert one.chain.shape == (24, 0, 6) and one.iterations == 0
    thin = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=7)
    thin.run_mcmc(pos0, 10, thin=5)
    assert np.array_equal(thin.chain, straight.chain[:, ::5, :])


def test_emcee_error_contract(nb):
    lc, pri, pos0 = _problem()
    with pytest.raises(TypeError):
        nb.EnsembleSampler(24, 6, lambda p: 0.0, args=(lc, pri))
    with pytest.raises(AssertionError):
        nb.EnsembleSampler(23, 6, nb.MonoOnlyLogProb, args=(lc, pri))
    with pytest.raises(AssertionError):
        nb.EnsembleSampler(10, 6, nb.MonoOnlyLogProb, args=(lc, pri))
    s = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=1)
    with pytest.raises(ValueError):
        s.run_mcmc(None, 1)
    bad = pos0.copy()
    bad[3, 2] = np.nan
    with pytest.raises(ValueError):
        s.run_mcmc(bad, 1)
    bad = pos0.copy()
    bad[3, 4] = np.inf
    with pytest.raises(ValueError):
        s.run_mcmc(bad, 1)


def test_posterior_agrees_with_cpu_chain_statistically(nb):
    """Different seeds on CPU and GPU: posterior medians must agree within Monte-Carlo error."""
    from oracle import namaste_oracle as o, stretch_oracle as so
    lc, pri, pos0 = _problem()
    rng = np.random.default_rng(3)
    pos0 = np.repeat(pos0[:1], 32, axis=0) * (1 + 1e-4 * rng.standard_normal((32, 6)))
    nsteps = 600
    s = nb.EnsembleSampler(32, 6, nb.MonoOnlyLogProb, args=(lc, pri), seed=111)
    s.run_mcmc(pos0, nsteps)
    gpu = s.chain[:, nsteps // 2:, :].reshape(-1, 6)
    cpu_chain, _, _ = so.run(lambda p: o.lnprob_ensemble(p, lc, pri), pos0, nsteps, 222)
    cpu = cpu_chain[:, nsteps // 2:, :].reshape(-1, 6)
    for d in range(6):
        sd = 0.5 * (gpu[:, d].std() + cpu[:, d].std())
        # ~ (32 walkers * 300 steps) / tau(~60) independent samples per chain -> s.e. ~ sd / 12
        assert abs(np.median(gpu[:, d]) - np.median(cpu[:, d])) < 0.5 * sd, d


def test_runmcmc_driver(nb):
    """RunMCMC (Star.RunMCMC, namaste.py:537-608, transit-only branch) end to end on the example star:
    initial ensemble from the reference's own PDF table, one-sided window quirk (N = 2828), 1 + nsteps
    appended chain, trimming; the posterior sits on the reference's logged best fit."""
    g = load_golden("priors_epic203311200")
    c1 = load_golden("config1_epic203311200")
    kat = load_golden("kat_epic203311200")
    pri = priors_from_table(c1["priors"])
    model = nb.MonotransitModel(*kat["params"][0])
    # the flattened light curve RunMCMC sees = config 1's before the window cut; here its windowed form
    # is used as the input, which the window mask keeps entirely (every t - tcen < 6 Tdur)
    nsteps = 400
    out = nb.RunMCMC(c1["lc"], pri, g["pdfs"], model, nwalkers=24, nsteps=nsteps,
                     rng=np.random.default_rng(203311200), seed=7)
    assert out["mask"].all() and out["mask"].shape == (2828,)
    s = out["sampler"]
    assert s.chain.shape == (24, 1 + nsteps, 6) and s.lnprobability.shape == (24, 1 + nsteps)
    assert out["ncut"] == 100 and out["sampleheaders"][-1] == "logprob" and len(out["sampleheaders"]) == 7
    smp = out["samples"]
    assert smp.shape == (int(out["good_wlkrs"].sum()) * (1 + nsteps - 100), 7) and np.isfinite(smp).all()
    assert out["good_wlkrs"].sum() >= 12
    # rows of the PDF table, NaN -> 0.0 as the reference does
    assert out["init_mcmc_params"].shape == (24, 6) and np.isfinite(out["init_mcmc_params"]).all()
    best = kat["params"][0]
    med = np.median(smp[:, :6], axis=0)
    assert abs(med[0] - best[0]) < 0.01            # tcen [d]
    assert abs(med[3] - best[3]) < 0.01            # RpRs
    assert np.median(smp[:, 6]) >= np.median(s.lnprobability[:, 0])   # the ensemble moved uphill from the PDF draws


def test_split_sampler_on_one_rank_is_the_plain_sampler(nb):
    """The walker-split driver (nmb_sampler_run_split) with a world of one -- what a single-GPU box can exercise --
    gives the chain of EnsembleSampler bit for bit, stored or not; the multi-rank form is checked by
    tests/multigpu_check.py on 2, 4 and 8 GPUs."""
    from namaste_b200.distributed import SplitEnsembleSampler
    lc, pri, pos0 = _problem()
    staged = nb.LightCurve(lc)
    plain = nb.EnsembleSampler(24, 6, nb.MonoOnlyLogProb, args=(staged, pri), seed=99)
    plain.run_mcmc(pos0, 12)
    split = SplitEnsembleSampler(staged, pri, 24, seed=99)
    assert split.transport == "local"
    split.run_mcmc(pos0, 5)
    pos, lnp = split.run_mcmc(None, 7)
    assert np.array_equal(split.chain, plain.chain) and np.array_equal(split.lnprobability, plain.lnprobability)
    assert np.array_equal(split.naccepted, plain.naccepted)
    assert np.array_equal(pos, plain.chain[:, -1, :]) and np.array_equal(lnp, plain.lnprobability[:, -1])
    quiet = SplitEnsembleSampler(staged, pri, 24, seed=99)
    p2, l2 = quiet.run_mcmc(pos0, 12, storechain=False)
    assert np.array_equal(p2, pos) and quiet.chain.shape == (24, 0, 6)


def test_split_api_refuses_bad_arguments(nb):
    import ctypes
    from namaste_b200._lib import load, ptr
    from namaste_b200.distributed import SplitEnsembleSampler
    lc, pri, pos0 = _problem()
    staged = nb.LightCurve(lc)
    s = SplitEnsembleSampler(staged, pri, 24, seed=1)
    lib = load()
    handles = np.zeros(64 * 4, dtype=np.uint8)
    assert lib.nmb_sampler_split_export(s._h, ptr(handles)) == 0 and handles[:64].any()
    assert lib.nmb_sampler_split_attach(s._h, 0, 17, ptr(handles)) < 0          # more ranks than the library maps
    assert lib.nmb_sampler_split_attach(s._h, 3, 2, ptr(handles)) < 0           # rank outside the world
    assert lib.nmb_sampler_split_attach(s._h, 0, 5, ptr(handles)) < 0           # 12 walkers per half do not divide by 5
    assert lib.nmb_sampler_run_split(s._h, 1, 0) < 0                            # no state yet
    assert b"never been called" in lib.nmb_last_error()
    with pytest.raises(ValueError):
        SplitEnsembleSampler(staged, pri, 24, seed=1, transport="carrier pigeon")
    with pytest.raises(ValueError):
        SplitEnsembleSampler(nb.LightCurve([lc, lc]), pri, 24, seed=1)              # one light curve only


@pytest.mark.parametrize("groups", [1, 2, 3])
def test_candidate_groups_give_the_one_batch_chains(nb, groups):
    """CandidateBatchSampler steps a catalogue of independent candidates as several groups (own staged batch,
    sampler, stream and host thread each); with every group told the index of its first candidate the chains are
    those of ONE batch under one sampler with the same seed, bit for bit."""
    g = load_golden("synth_k2_s11")
    base = g["lc"]
    mid = len(base) // 2
    cuts = [(mid - 600, mid + 500), (mid - 300, mid + 300), (0, len(base)), (mid - 900, mid + 200), (mid - 250, mid + 700)]
    curves = [np.ascontiguousarray(base[a:b]) for a, b in cuts]
    C, W, nsteps, seed = len(curves), 16, 6, 4242
    rng = np.random.default_rng(8)
    pos0 = g["walkers"][0][None, None, :] * (1 + 1e-4 * rng.standard_normal((C, W, 6)))
    tabs = np.stack([g["priors"]] * C)
    one_lc = nb.LightCurve(curves, oversamp=int(g["oversamp"]))
    one = nb.EnsembleSampler(W, 6, nb.MonoOnlyLogProb, args=(one_lc, tabs), seed=seed)
    one.run_mcmc(pos0, nsteps)
    grp = nb.CandidateBatchSampler(curves, tabs, W, oversamp=int(g["oversamp"]), seed=seed, groups=groups)
    pos, lnp = grp.run_mcmc(pos0, nsteps)
    assert grp.chain.shape == (C, W, nsteps, 6)
    assert np.array_equal(grp.chain, one.chain) and np.array_equal(grp.lnprobability, one.lnprobability)
    assert np.array_equal(grp.naccepted, one.naccepted)
    assert np.array_equal(pos, one.chain[:, :, -1, :]) and np.array_equal(lnp, one.lnprobability[:, :, -1])
    pos2, _ = grp.run_mcmc(None, 2)                       # continues
    one.run_mcmc(None, 2)
    assert np.array_equal(pos2, one.chain[:, :, -1, :])
    with pytest.raises(ValueError):
        grp.run_mcmc(pos0[:2], 1)
    grp.close()


def test_candidate_groups_on_a_second_device(nb):
    """The group threads select the device the sampler was created on (the current CUDA device is per host thread)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden("synth_k2_s11")
    curves = [np.ascontiguousarray(g["lc"][a:a + 900]) for a in (1500, 1600, 1700, 1400)]
    W, seed = 16, 99
    pos0 = g["walkers"][0][None, None, :] * (1 + 1e-4 * np.random.default_rng(2).standard_normal((4, W, 6)))
    tabs = np.stack([g["priors"]] * 4)
    chains = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            grp = nb.CandidateBatchSampler(curves, tabs, W, oversamp=int(g["oversamp"]), seed=seed, groups=2)
            grp.run_mcmc(pos0, 4)
            chains.append(grp.chain)
            grp.close()
    assert np.array_equal(chains[0], chains[1])
