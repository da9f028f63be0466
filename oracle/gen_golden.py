#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (this container only).

TEST INFRASTRUCTURE ONLY.  Run as ``python oracle/gen_golden.py`` where
/root/reference is mounted; the resulting small .npz files are committed so that
the GPU box (which has no /root/reference) can check both the oracle and the
CUDA path against numbers produced by the reference's own code:

  * occultquad / occultuniform / ellke / ellpic_bulirsch values come from the
    unmodified ``namaste/Crossfield_transit.py`` imported by path;
  * light curves are prepared with the reference's ``k2flatten.ReduceNoise`` and
    the recipe of SURVEY.md section 8(c) (Example.ipynb cells 45-47);
  * lnprob values = reference occultquad + the ~10 line wrapper of
    namaste.py:1155-1164 / :1345-1349 / :1314-1332, which cannot be imported
    (namaste.py needs autograd/celerite/emcee/astropy) and is written out here
    once, independently of oracle/namaste_oracle.py, so the fixtures also pin
    the oracle's wrapper.

Only numbers are written; no reference source is copied.
"""
import os
import sys

import numpy as np
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import _refimport  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = _refimport.REFERENCE_ROOT

# Known answers logged by the reference (Example.ipynb cell 47; namaste_runlog.log tag 1048)
KAT_PARAMS = np.array([
    [2121.023968717218, 0.7868651518286128, 2.6169560790026534, 0.058475207050375115, 0.1358330260599801, 0.29597158586826544],
    [2121.0239686579657, 0.7868626919714633, 2.6169673801508675, 0.05847517954647115, 0.13583822141944843, 0.2959726275491803],
    [2121.023968745533, 0.7868616818202295, 2.6169731129933482, 0.058475160201592394, 0.13583837575643895, 0.2959721839623829],
])
KAT_NLL = np.array([1170.0458538629673, 1170.0458540024072, 1170.0458540367601])
# prior dict printed by the reference for that run (tag 1023), as a [6,6] table:
# lo, hi, kind (0 none, 1 gaussian, 2 normlim), mu, sigma, pad
KAT_PRIORS = np.array([
    [2120.87, 2121.17, 0, 0, 0, 0],
    [0.0, 1.25, 0, 0, 0, 0],
    [0, 5.4903083994500745, 2, 3.5736491649686775, 0.3833318468962794, 0],
    [0.02, 0.25, 0, 0, 0, 0],
    [0, 1.0, 1, 0.339679200660016, 0.038476153186633194, 0],
    [0, 1.0, 1, 0.295936861708641, 0.022151103527701013, 0],
])


def table_to_dict(tab):
    names = ["mean:tcen", "mean:b", "mean:vel", "mean:RpRs", "mean:LD1", "mean:LD2"]
    kinds = {1: "gaussian", 2: "normlim", 3: "evans"}
    d = {}
    for nm, row in zip(names, tab):
        if int(row[2]) == 0:
            d[nm] = [row[0], row[1]]
        else:
            d[nm] = [row[0], row[1], kinds[int(row[2])], row[3], row[4]]
    return d


# --- the reference wrapper, written out against the reference's occultquad -------------
def ref_get_value(cf, params, t, oversamp=10):
    tcen, b, vel, rprs, ld1, ld2 = [np.float64(v) for v in params]
    cad = np.median(np.diff(t))
    finetime = np.sort(np.hstack(([t + o for o in cad * np.arange(0, 1.0, 1 / oversamp)])))
    z = np.sqrt(b ** 2 + (vel * (finetime - tcen)) ** 2)
    model = cf.occultquad(z, rprs, np.array((ld1, ld2)))
    return np.average(np.resize(model, (len(t), oversamp)), axis=1)


def ref_ln_prior(params, priors):
    lp = 0
    for n, key in enumerate(priors.keys()):
        if params[n] < priors[key][0] or params[n] > priors[key][1]:
            lp -= 1e20 * (((params[n] - 0.5 * (priors[key][0] + priors[key][1]))) / (priors[key][1] - priors[key][0])) ** 2
        if len(priors[key]) > 2:
            if priors[key][2] == 'gaussian':
                lp += stats.norm(priors[key][3], priors[key][4]).pdf(params[n])
            elif priors[key][2] == 'normlim':
                lp += ((1.0 - stats.norm.cdf(params[n], priors[key][3], priors[key][4])) * params[n]) ** 2
    return lp


def ref_lnprob(cf, params, lc, priors, oversamp=10):
    ll = np.sum(-2 * lc[:, 2] ** -2 * (lc[:, 1] - ref_get_value(cf, params, lc[:, 0], oversamp)) ** 2)
    return ll, ref_ln_prior(params, priors)


# --- light-curve recipe (namaste.py:1098-1113, planetlib.py:710-725) --------
def load_example_lc():
    lc = np.load(os.path.join(REF, "ExampleData", "EPIC203311200_bestlc_ev.npy"))
    if len(lc) == 2:
        lc = lc[0]
    lc = lc[np.logical_not(np.isnan(np.sum(lc, axis=1)))]
    lc[:, 1:] /= np.median(lc[:, 1])
    # Lightcurve.__init__
    lc = lc[~np.isnan(np.sum(lc, axis=1))]
    lc[:, 1:] /= np.nanmedian(lc[:, 1])
    from oracle.namaste_oracle import anom_cut_diff
    lc = lc[anom_cut_diff(lc[:, 1], 3.75)]
    return lc


def synthetic_lc(cf, cfg, n, cad, t0, sigma, oversamp, theta=None):
    """SURVEY.md 8(d) synthetic recipe; returns lc[n,3] and truth."""
    rng = np.random.default_rng(151203722 + cfg)
    t = t0 + cad * np.arange(n)
    if theta is None:
        theta = np.array([t[n // 2] + 0.37 * cad, 0.5, 3.0, 0.06, 0.34, 0.30])
    flux = ref_get_value(cf, theta, t, oversamp) + sigma * rng.standard_normal(n)
    return np.column_stack((t, flux, np.full(n, sigma))), theta


def synthetic_priors(theta):
    tdur = 2 * (1 + theta[3]) * np.sqrt(1 - (theta[1] / (1 + theta[3])) ** 2) / theta[2]
    tab = KAT_PRIORS.copy()
    tab[0, :2] = [theta[0] - 0.3 * tdur, theta[0] + 0.3 * tdur]
    tab[2, 1] = theta[2] * 1.6
    tab[2, 3] = theta[2] * 1.1
    tab[2, 4] = theta[2] * 0.12
    tab[4, 3], tab[5, 3] = theta[4], theta[5]
    return tab


def synthetic_walkers(rng, theta, ptab, w):
    pos = theta[None, :] * (1 + 1e-3 * rng.standard_normal((w, 6)))
    nbox = max(1, w // 20)
    lo, hi = ptab[:, 0], ptab[:, 1]
    pos[:nbox] = lo + (hi - lo) * rng.uniform(-0.05, 1.05, (nbox, 6))  # a few beyond the soft walls
    return pos


def main():
    os.makedirs(OUT, exist_ok=True)
    cf = _refimport.crossfield()
    kf = _refimport.k2flatten()
    rng = np.random.default_rng(20261019)

    # 1. special functions -------------------------------------------------------------
    k = np.concatenate([rng.uniform(0, 1, 400), [0.0, 0.3, 0.9, 0.999999, 1 - 1e-12]])
    n = np.concatenate([rng.uniform(0, 300, 400), [2.0, 120.0, 0.0, 1e-9, 1e6]])
    ee, kk = cf.ellke(k)
    # Bulirsch is batch-coupled in the reference (loops until ALL converge) -> also store
    # element-by-element values, which is what a per-thread device loop reproduces.
    pi_batch = cf.ellpic_bulirsch(n, k)
    pi_each = np.array([cf.ellpic_bulirsch(np.array([a]), np.array([b]))[0] for a, b in zip(n, k)])
    np.savez_compressed(os.path.join(OUT, "special.npz"), k=k, n=n, E=ee, K=kk, Pi_batch=pi_batch, Pi_each=pi_each)

    # 2. occultquad case table ------------------------------------------------------------
    zs, ps, g1s, g2s, out = [], [], [], [], []
    gammas = [(0.3, 0.2), (0.0, 0.0), (0.6, -0.1), (0.1358, 0.2960), (1.2, 0.4)]
    plist = [0.01, 0.0585, 0.1, 0.25, 0.4, 0.5, 0.6, 0.9, 1.0, 1.5, -0.1]
    for p in plist:
        ap = abs(p)
        special = [0.0, ap, 1 - ap, 1 + ap, abs(1 - ap), 0.5, np.nextafter(1 + ap, 0), np.nextafter(1 + ap, 9),
                   np.nextafter(ap, 0), np.nextafter(ap, 9), 1e-9, 1e-300]
        z = np.abs(np.concatenate([rng.uniform(0, 1.3 + ap, 300), special,
                                   1 + ap - 10.0 ** -rng.uniform(3, 15, 20),
                                   np.abs(1 - ap + 10.0 ** -rng.uniform(3, 15, 20) * rng.choice([-1, 1], 20)),
                                   np.abs(ap + 10.0 ** -rng.uniform(3, 15, 20) * rng.choice([-1, 1], 20))]))
        for g in gammas:
            F, le, ld, ed = cf.occultquad(z, p, np.array(g), retall=True)
            zs.append(z); ps.append(np.full(z.size, p)); g1s.append(np.full(z.size, g[0])); g2s.append(np.full(z.size, g[1]))
            out.append(np.column_stack((F, le, ld, ed)))
            # per-element evaluation (decoupled Bulirsch batches)
    z = np.concatenate(zs); p = np.concatenate(ps); g1 = np.concatenate(g1s); g2 = np.concatenate(g2s)
    batch = np.concatenate(out)
    each = np.array([np.array(cf.occultquad(np.array([zz]), pp, np.array((a, b)), retall=True)).ravel()
                     for zz, pp, a, b in zip(z, p, g1, g2)])
    uni = np.array([cf.occultuniform(np.array([zz]), pp)[0] for zz, pp in zip(z, p)])
    np.savez_compressed(os.path.join(OUT, "occultquad_cases.npz"), z=z, p=p, g1=g1, g2=g2, batch=batch, each=each, uniform=uni)

    # 3. EPIC203311200 known answers (226-point window, flattened twice) -----------------
    lc0 = load_example_lc()
    flat1 = kf.ReduceNoise(lc0, winsize=4.5, stepsize=0.125)
    flat2 = kf.ReduceNoise(flat1, winsize=4.5, stepsize=0.125)
    kat_lc = flat2[abs(flat2[:, 0] - 2121.02) < 5 * 0.5]
    pri = table_to_dict(KAT_PRIORS)
    got = np.array([ref_lnprob(cf, th, kat_lc, pri) for th in KAT_PARAMS])
    print("KAT lc points:", len(kat_lc), " nll:", -(got.sum(axis=1)), " rel:", (-(got.sum(axis=1)) - KAT_NLL) / KAT_NLL)
    models = np.array([ref_get_value(cf, th, kat_lc[:, 0]) for th in KAT_PARAMS])
    np.savez_compressed(os.path.join(OUT, "kat_epic203311200.npz"), lc=kat_lc, params=KAT_PARAMS, nll_logged=KAT_NLL,
             priors=KAT_PRIORS, ll=got[:, 0], lp=got[:, 1], model=models)

    # 4. config 1: RunMCMC input (flattened once, one-sided mask quirk Q6), W=100 ---------
    th = KAT_PARAMS[0]
    tdur = (2 * (1 + th[3]) * np.sqrt(1 - (th[1] / (1 + th[3])) ** 2)) / th[2]
    mask = abs(flat1[:, 0] - th[0] < 6 * tdur)
    lc1 = flat1[mask.astype(bool)]
    pdfs = np.load(os.path.join(REF, "InputFiles", "203311200.1_inputsamples.npy"))
    rows = np.random.default_rng(203311200).choice(6000, 100, replace=False)
    walkers = pdfs[rows][:, [0, 1, 3, 2, 9, 10]].copy()
    walkers[np.isnan(walkers)] = 0.0  # namaste.py:541-542 (nanmedian of a boolean mask == 0.0)
    res = np.array([ref_lnprob(cf, w, lc1, pri) for w in walkers])
    print("config1 N =", len(lc1), "walkers", walkers.shape, "finite:", np.isfinite(res.sum(axis=1)).sum())
    np.savez_compressed(os.path.join(OUT, "config1_epic203311200.npz"), lc=lc1, walkers=walkers, priors=KAT_PRIORS,
             ll=res[:, 0], lp=res[:, 1])

    # 5. reduced synthetic configs (same recipe as BASELINE configs 2/3/5, small W) --------
    for name, cfg, nn, cad, t0, sig, S, W in [
        ("synth_k2_s11", 2, 4000, 0.0204318, 2061.0, 5.4e-5, 11, 48),
        ("synth_kepler_s15", 3, 6500, 0.02043361, 131.5, 1e-4, 15, 24),
        ("synth_tess_s10", 5, 12000, 2.0 / 1440, 1325.0, 6e-4, 10, 24),
        ("synth_tess_s1", 6, 5000, 2.0 / 1440, 1325.0, 6e-4, 1, 24),
    ]:
        lc, theta = synthetic_lc(cf, cfg, nn, cad, t0, sig, S)
        ptab = synthetic_priors(theta)
        pos = synthetic_walkers(np.random.default_rng(cfg), theta, ptab, W)
        pos[-1, 2] = 0.0          # a vel = 0 walker (quirk Q5): every fine sample in transit
        pos[-2, 3] = 0.55         # p > 1/2 branch (i08/i09 large planet)
        pos[-3, 1] = 1.2          # grazing / non-transiting
        pos[-4, 3] = -0.06        # negative radius ratio -> |p|
        pd = table_to_dict(ptab)
        res = np.array([ref_lnprob(cf, w, lc, pd, S) for w in pos])
        mdl = np.array([ref_get_value(cf, w, lc[:, 0], S) for w in pos[:4]])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), lc=lc, walkers=pos, priors=ptab, ll=res[:, 0], lp=res[:, 1],
                 oversamp=S, theta=theta, model4=mdl)
        print(name, "N", nn, "S", S, "W", W, "ll range", res[:, 0].min(), res[:, 0].max())


if __name__ == "__main__":
    main()
