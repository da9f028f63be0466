"""CPU oracle for the Namaste single-transit likelihood hot path.

TEST INFRASTRUCTURE ONLY.  This module is the checker, never the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it.  Nothing under ``namaste_b200/`` imports it.

It is a NumPy restatement (written from the algorithm, not copied) of the
reference's arithmetic, operation order preserved so that on the same NumPy it
is bit-identical to the reference functions it follows:

  ellke            <- namaste/Crossfield_transit.py:224-270  (Hastings E,K)
  ellpic_bulirsch  <- namaste/Crossfield_transit.py:272-342  (Bulirsch Pi)
  occultuniform    <- namaste/Crossfield_transit.py:433-523  (array branch)
  occultquad       <- namaste/Crossfield_transit.py:690-970  (11-case M&A)
  get_value        <- namaste/namaste.py:1155-1164           (supersample+bin)
  mono_ln_prior    <- namaste/namaste.py:1314-1332           (MonoLnPrior)
  mono_only_logprob<- namaste/namaste.py:1345-1349           (MonoOnlyLogProb)
  calc_tdur        <- namaste/namaste.py:1469-1472
  anom_cut_diff    <- namaste/namaste.py:1459-1467

Parity pin: ``tests/test_oracle_vs_reference.py`` compares every function here
bitwise with the reference modules imported from /root/reference (when that
mount exists), and ``tests/test_oracle_kat.py`` reproduces the three
known-answer negative log-probabilities logged in Example.ipynb cell 47 /
namaste/namaste_runlog.log (tag ``1048:namaste.py``) from committed fixtures.
"""
from __future__ import annotations

import numpy as np
from scipy import stats

EPS = np.finfo(float).eps
PARAM_NAMES = ("tcen", "b", "vel", "RpRs", "LD1", "LD2")  # namaste.py:1153

# Hastings coefficients (Crossfield_transit.py:243-265)
_EA = (0.44325141463, 0.06260601220, 0.04757383546, 0.01736506451)
_EB = (0.24998368310, 0.09200180037, 0.04069697526, 0.00526449639)
_KA = (1.38629436112, 0.09666344259, 0.03590092383, 0.03742563713, 0.01451196212)
_KB = (0.5, 0.12498593597, 0.06880248576, 0.03328355346, 0.00441787012)


def ellke(k):
    """Hastings polynomial complete elliptic integrals; returns (E, K)."""
    m1 = 1. - k ** 2
    lg = np.log(m1)
    e_poly = 1. + m1 * (_EA[0] + m1 * (_EA[1] + m1 * (_EA[2] + m1 * _EA[3])))
    e_log = m1 * (_EB[0] + m1 * (_EB[1] + m1 * (_EB[2] + m1 * _EB[3]))) * (-lg)
    k_poly = _KA[0] + m1 * (_KA[1] + m1 * (_KA[2] + m1 * (_KA[3] + m1 * _KA[4])))
    k_log = (_KB[0] + m1 * (_KB[1] + m1 * (_KB[2] + m1 * (_KB[3] + m1 * _KB[4])))) * lg
    return e_poly + e_log, k_poly - k_log


def ellpic_bulirsch(n, k, tol=1000 * EPS, maxiter=1e4):
    """Complete elliptic integral of the third kind, Bulirsch (1965) AGM scheme.

    The batch iterates until *every* element is inside tolerance
    (Crossfield_transit.py:319-332), so already-converged elements keep
    being refined; that is reproduced here.
    """
    n = np.atleast_1d(np.asarray(n, dtype=float))
    k = np.atleast_1d(np.asarray(k, dtype=float))
    if n.size == 0 or k.size == 0:
        return np.array([])
    kc = np.sqrt(1. - k ** 2)
    p = np.sqrt(n + 1.)
    m0 = np.array(1.)
    c = np.array(1.)
    d = 1. / p
    e = kc.copy()
    it = 0
    while it < maxiter:
        f = c
        c = d / p + c
        g = e / p
        d = 2. * (f * g + d)
        p = g + p
        g = m0
        m0 = kc + m0
        if (np.abs(1. - kc / g) > tol).any():
            kc = 2. * np.sqrt(e)
            e = kc * m0
            it += 1
        else:
            break
    return .5 * np.pi * (c * m0 + d) / (m0 * (m0 + p))


def occultuniform(z, p):
    """Uniform-source flux 1 - (occulted fraction); array branch only."""
    z = np.abs(np.array(z, dtype=float, copy=True))
    neg = p < 0
    p = np.abs(p)
    frac = np.zeros(z.shape, float)
    p2 = p * p
    outside = (1 + p) < z
    partial = (np.abs(1 - p) < z) * (z <= (1 + p))
    inside = z <= (1 - p)
    covered = z <= (p - 1)
    if partial.any():
        zp = z[partial]
        zp2 = zp * zp
        arg1 = 1 - p2 + zp2
        ac1 = (p2 + zp2 - 1) / (2. * p * zp)
        ac2 = arg1 / (2 * zp)
        ac1[ac1 > 1] = 1.
        ac2[ac2 > 1] = 1.
        k0 = np.arccos(ac1)
        k1 = np.arccos(ac2)
        k2 = 0.5 * np.sqrt(4 * zp2 - arg1 * arg1)
        frac[partial] = (1. / np.pi) * (p2 * k0 + k1 - k2)
    frac[outside] = 0.
    if inside.any():
        frac[inside] = p2
    if covered.any():
        frac[covered] = 1.
    if neg:
        frac *= -1
    return 1. - frac


def occultquad(z, p0, gamma, retall=False):
    """Mandel & Agol (2002) quadratic limb-darkened flux, Eastman-corrected.

    Case numbering i01..i11 follows Crossfield_transit.py:809-819.
    """
    z = np.asarray(z, dtype=float)
    lam_d = np.zeros(z.shape, float)
    eta_d = np.zeros(z.shape, float)
    p = np.abs(p0)
    gamma = np.asarray(gamma, dtype=float)
    c2 = gamma[0] + 2 * gamma[1]
    c4 = -gamma[1]
    if p == 0:
        one = np.ones(z.shape, float)
        if retall:
            return one, np.ones(z.shape, float), np.zeros(z.shape, float), np.zeros(z.shape, float)
        return one

    four_omega = 1. - gamma[0] / 3. - gamma[1] / 6.
    a = (z - p) * (z - p)
    b = (z + p) * (z + p)
    p2 = p * p
    z2 = z * z
    nine_pi = 9 * np.pi

    pos = p > 0
    i01 = pos * (z >= (1. + p))
    i02 = pos * (z > (.5 + np.abs(p - 0.5))) * (z < (1. + p))
    i03 = pos * (p < 0.5) * (z > p) * (z < (1. - p))
    i04 = pos * (p < 0.5) * (z == (1. - p))
    i05 = pos * (p < 0.5) * (z == p)
    i06 = (p == 0.5) * (z == 0.5)
    i07 = (p > 0.5) * (z == p)
    i08 = (p > 0.5) * (z >= np.abs(1. - p)) * (z < p)
    i09 = pos * (p < 1) * (z > 0) * (z < (0.5 - np.abs(p - 0.5)))
    i10 = pos * (p < 1) * (z == 0)
    i11 = (p > 1) * (z >= 0.) * (z < (p - 1.))

    lam_d[i01] = 0.
    eta_d[i01] = 0.
    if i06.any():
        lam_d[i06] = 1. / 3. - 4. / nine_pi
        eta_d[i06] = 0.09375
    if i11.any():
        lam_d[i11] = 1.
        eta_d[i11] = 0.5

    # Lambda_1: ingress/egress (i02) and large-planet partial (i08)
    g1 = i02 + i08
    q1 = p2 - z2[g1]
    qq = np.sqrt((1. - a[g1]) / (b[g1] - a[g1]))
    pi3 = ellpic_bulirsch(1. / a[g1] - 1., qq)
    ee, kk = ellke(qq)
    lam_d[g1] = (1. / (nine_pi * np.sqrt(p * z[g1]))) * \
        (((1. - b[g1]) * (2 * b[g1] + a[g1] - 3) - 3 * q1 * (b[g1] - 2.)) * kk +
         4 * p * z[g1] * (z2[g1] + 7 * p2 - 4.) * ee -
         3 * (q1 / a[g1]) * pi3)

    # Lambda_2: planet disk fully inside the stellar disk (i03, i09)
    g2 = i03 + i09
    q2 = p2 - z2[g2]
    a2 = a[g2]
    b2 = b[g2]
    oma = 1. - a2
    pi3 = ellpic_bulirsch(b2 / a2 - 1, np.sqrt((b2 - a2) / (oma)))
    ee, kk = ellke(np.sqrt((b2 - a2) / (oma)))
    lam_d[g2] = (2. / (nine_pi * np.sqrt(oma))) * \
        ((1. - 5 * z2[g2] + p2 + q2 * q2) * kk +
         (oma) * (z2[g2] + 7 * p2 - 4.) * ee -
         3 * (q2 / a2) * pi3)

    if i07.any():  # Lambda_3 (z == p > 1/2)
        ee, kk = ellke(0.5 / p)
        lam_d[i07] = 1. / 3. + (16. * p * (2 * p2 - 1.) * ee -
                                (1. - 4 * p2) * (3. - 8 * p2) * kk / p) / nine_pi
    if i05.any():  # Lambda_4 (z == p < 1/2)
        ee, kk = ellke(2. * p)
        lam_d[i05] = 1. / 3. + (2. / nine_pi) * (4 * (2 * p2 - 1.) * ee + (1. - 4 * p2) * kk)
    if i04.any():  # Lambda_5 (z == 1 - p)
        lam_d[i04] = (2. / 3.) * (np.arccos(1. - 2 * p) / np.pi -
                                  (6. / nine_pi) * np.sqrt(p * (1. - p)) *
                                  (3. + 2 * p - 8 * p2) -
                                  float(p > 0.5))
    if i10.any():  # Lambda_6 (z == 0)
        lam_d[i10] = -(2. / 3.) * (1. - p2) ** 1.5

    # eta_1 (arccos arguments are NOT clamped here, unlike occultuniform)
    g3 = g1 + i07
    z2g = z2[g3]
    twoz = 2 * z[g3]
    eta_d[g3] = (0.5 / np.pi) * ((np.arccos((1 - p2 + z2g) / (twoz))) +
                                 (np.arccos((p2 + z2g - 1) / (p * twoz))) * p2 * (p2 + 2 * z2g) -
                                 0.25 * (1. + 5 * p2 + z2g) *
                                 np.sqrt((1. - a[g3]) * (b[g3] - 1.)))
    # eta_2
    g4 = g2 + i04 + i05 + i10
    eta_d[g4] = 0.5 * p2 * (p2 + 2. * z2[g4])

    lam_e = 1. - occultuniform(z, p)
    F = 1. - ((1. - c2) * lam_e +
              c2 * (lam_d + (2. / 3.) * (p > z).astype(float)) -
              c4 * eta_d) / four_omega
    if retall:
        return F, lam_e, lam_d, eta_d
    return F


def fine_time_grid(t, oversamp=10):
    """Sorted supersampling grid of namaste.py:1157-1160 (forward offsets)."""
    t = np.asarray(t, dtype=float)
    cad = np.median(np.diff(t))
    fine = np.hstack(([t + dt for dt in cad * np.arange(0, 1.0, 1 / oversamp)]))
    return np.sort(fine), cad


def get_value(params, t, oversamp=10):
    """MonotransitModel.get_value (namaste.py:1155-1164) for one parameter vector."""
    tcen, b, vel, rprs, ld1, ld2 = [np.float64(v) for v in params]
    fine, _ = fine_time_grid(t, oversamp)
    z = np.sqrt(b ** 2 + (vel * (fine - tcen)) ** 2)
    model = occultquad(z, rprs, np.array((ld1, ld2)))
    return np.average(np.resize(model, (len(t), oversamp)), axis=1)


def mono_ln_prior(params, priors):
    """MonoLnPrior (namaste.py:1314-1332). ``priors`` is an ordered mapping
    name -> [lo, hi] or [lo, hi, kind, mu, sigma]; entry n applies to params[n]."""
    lp = 0
    for n, key in enumerate(priors.keys()):
        row = priors[key]
        if params[n] < row[0] or params[n] > row[1]:
            lp -= 1e20 * (((params[n] - 0.5 * (row[0] + row[1]))) / (row[1] - row[0])) ** 2
        if len(row) > 2:
            if row[2] == 'gaussian':
                lp += stats.norm(row[3], row[4]).pdf(params[n])
            elif row[2] == 'normlim':
                lp += ((1.0 - stats.norm.cdf(params[n], row[3], row[4])) * params[n]) ** 2
            elif row[2] == 'evans':
                lp -= row[3] * np.exp(params[n])
    return lp


def log_like(params, lc, oversamp=10):
    """Gaussian (x -2) log-likelihood of namaste.py:1347."""
    return np.sum(-2 * lc[:, 2] ** -2 * (lc[:, 1] - get_value(params, lc[:, 0], oversamp)) ** 2)


def mono_only_logprob(params, lc, priors, oversamp=10):
    """MonoOnlyLogProb (namaste.py:1345-1349)."""
    return log_like(params, lc, oversamp) + mono_ln_prior(params, priors)


def lnprob_ensemble(params, lc, priors, oversamp=10):
    """Row-wise mono_only_logprob for an ensemble [W, 6]."""
    params = np.atleast_2d(np.asarray(params, dtype=float))
    return np.array([mono_only_logprob(row, lc, priors, oversamp) for row in params])


def calc_tdur(vel, b, p):
    """namaste.py:1469-1472."""
    return (2 * (1 + p) * np.sqrt(1 - (b / (1 + p)) ** 2)) / vel


def anom_cut_diff(flux, thresh=4.2):
    """Single-point outlier mask, namaste.py:1459-1467."""
    d = np.vstack((np.diff(flux[1:]), np.diff(flux[:-1])))
    d /= np.median(abs(d[0, :]))
    keep = ((d[0, :] * d[1, :]) > 0) + (abs(d[0, :]) < thresh) + (abs(d[1, :]) < thresh)
    return np.hstack((True, keep, True))


def prior_table(priors):
    """Flatten an ordered priors mapping into the [npar, 6] float table that the
    C ABI consumes: lo, hi, kind(0 none, 1 gaussian, 2 normlim, 3 evans), mu, sigma, pad."""
    kinds = {'gaussian': 1.0, 'normlim': 2.0, 'evans': 3.0}
    rows = []
    for key in priors.keys():
        row = priors[key]
        if len(row) > 2:
            rows.append([row[0], row[1], kinds[row[2]], row[3], row[4] if len(row) > 4 else 0.0, 0.0])
        else:
            rows.append([row[0], row[1], 0.0, 0.0, 0.0, 0.0])
    return np.array(rows, dtype=float)


# ---- post-run trimming (namaste/namaste.py:597-608), written walker by walker ----------------------
def trim_samples(chain, lnprobability, nsteps):
    """Reference: ncut = min(int(0.25 nsteps), 3000); a walker is kept when the 95th percentile of its
    post-burn-in lnprob exceeds the median over walkers of the per-walker medians; the kept walkers'
    post-burn-in samples are flattened walker-major with the lnprob appended as last column."""
    ncut = min(int(nsteps * 0.25), 3000)
    nw = len(chain)
    med = [float(np.percentile(lnprobability[w, ncut:], 50)) for w in range(nw)]
    p95 = [float(np.percentile(lnprobability[w, ncut:], 95)) for w in range(nw)]
    thresh = float(np.median(med))
    rows = []
    good = []
    for w in range(nw):
        keep = p95[w] > thresh
        good.append(keep)
        if keep:
            for tt in range(ncut, chain.shape[1]):
                rows.append(list(chain[w, tt, :]) + [lnprobability[w, tt]])
    return np.array(rows), np.array(good), ncut
