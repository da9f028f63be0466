"""CPU oracle of the GP noise-model objective ``MonoLogProb`` (test infrastructure only).

Reference call sites: namaste/namaste.py:1334-1339 (MonoLogProb: gp.set_parameter_vector, gp.compute,
gp.log_likelihood + MonoLnPrior), :279-314 (kernels: celerite RealTerm + JitterTerm, or the RotationTerm
of :1355-1379 + JitterTerm), :583 (sampler registration).

**Parity unpinned.**  The solver lives in the third-party package `celerite`, which is neither vendored
in the reference nor installed here (the reference pins no version; the API in use is celerite 0.3.x),
and the reference holds no logged value of this objective for a light curve it ships.  What follows
restates the published algorithm (Foreman-Mackey, Agol, Ambikasaran & Angus 2017, AJ 154, 220: the
semiseparable Cholesky factorisation of their eqs. 44-46 with the pre-conditioned U, V, phi of sec. 5.2)
and the log-likelihood celerite's ``GP.log_likelihood`` forms from it,
    -0.5 * (r^T K^-1 r + log det K + N log 2 pi),   r = y - mean(t),  K = k(t_i - t_j) + diag(yerr^2 + jitter).
It is checked in tests/test_gp_oracle.py against dense linear algebra on the same kernel function,
which is independent of the recurrence.
"""
import numpy as np

from . import namaste_oracle as o

LOG_2PI = np.log(2.0 * np.pi)

KERNELS = {
    # name: full kernel parameter names in celerite's order (term 0, then the jitter term)
    "Real": ("log_a", "log_c", "log_sigma"),
    "quasi": ("log_amp", "log_timescale", "log_period", "log_factor", "log_sigma"),
}


def coefficients(kind, kpar):
    """celerite coefficient arrays (a_real, c_real, a_comp, b_comp, c_comp, d_comp) and the jitter
    variance for the full kernel parameter vector ``kpar``.
    Real : RealTerm.get_real_coefficients = (exp(log_a), exp(log_c));
    quasi: RotationTerm (namaste.py:1355-1379)."""
    kpar = np.asarray(kpar, dtype=np.float64)
    if kind == "Real":
        log_a, log_c, log_sigma = kpar
        return (np.array([np.exp(log_a)]), np.array([np.exp(log_c)]), np.zeros(0), np.zeros(0), np.zeros(0),
                np.zeros(0), np.exp(2 * log_sigma))
    if kind == "quasi":
        log_amp, log_timescale, log_period, log_factor, log_sigma = kpar
        f = np.exp(log_factor)
        ar = np.exp(log_amp) * (1.0 + f) / (2.0 + f)
        cr = np.exp(-log_timescale)
        ac = np.exp(log_amp) / (2.0 + f)
        return (np.array([ar]), np.array([cr]), np.array([ac]), np.array([0.0]), np.array([np.exp(-log_timescale)]),
                np.array([2 * np.pi * np.exp(-log_period)]), np.exp(2 * log_sigma))
    raise ValueError(kind)


def kernel_value(tau, coeffs):
    ar, cr, ac, bc, cc, dc, _ = coeffs
    tau = np.abs(tau)
    k = np.zeros_like(tau)
    for a, c in zip(ar, cr):
        k = k + a * np.exp(-c * tau)
    for a, b, c, d in zip(ac, bc, cc, dc):
        k = k + np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau))
    return k


def dense_loglike(t, resid, yerr, coeffs):
    """Independent check: build K densely and use LAPACK."""
    K = kernel_value(t[:, None] - t[None, :], coeffs)
    K[np.diag_indices_from(K)] += yerr ** 2 + coeffs[6]
    sign, logdet = np.linalg.slogdet(K)
    assert sign > 0
    return -0.5 * (resid @ np.linalg.solve(K, resid) + logdet + len(t) * LOG_2PI)


def celerite_loglike(t, resid, yerr, coeffs):
    """The O(N J^2) semiseparable Cholesky of the celerite paper, one observation at a time."""
    ar, cr, ac, bc, cc, dc, jitter = coeffs
    jr, jc = len(ar), len(ac)
    J = jr + 2 * jc
    n = len(t)
    A = yerr ** 2 + jitter + np.sum(ar) + np.sum(ac)
    U = np.empty((n, J)); V = np.empty((n, J)); c_all = np.empty(J)
    U[:, :jr] = ar[None, :]; V[:, :jr] = 1.0; c_all[:jr] = cr
    for j in range(jc):
        cd, sd = np.cos(dc[j] * t), np.sin(dc[j] * t)
        U[:, jr + 2 * j] = ac[j] * cd + bc[j] * sd
        U[:, jr + 2 * j + 1] = ac[j] * sd - bc[j] * cd
        V[:, jr + 2 * j] = cd
        V[:, jr + 2 * j + 1] = sd
        c_all[jr + 2 * j] = c_all[jr + 2 * j + 1] = cc[j]
    S = np.zeros((J, J))
    D = A[0]
    W = V[0] / D
    f = np.zeros(J)
    z = resid[0]
    quad = z * z / D
    logdet = np.log(D)
    for i in range(1, n):
        phi = np.exp(-c_all * (t[i] - t[i - 1]))
        S = np.outer(phi, phi) * (S + D * np.outer(W, W))
        f = phi * (f + W * z)
        Di = A[i] - U[i] @ S @ U[i]
        W = (V[i] - U[i] @ S) / Di
        D = Di
        z = resid[i] - U[i] @ f
        quad += z * z / D
        logdet += np.log(D)
    return -0.5 * (quad + logdet + n * LOG_2PI)


def split_params(kind, params, kfixed, free_mask):
    """Full kernel vector + the six mean parameters from a walker vector (free kernel parameters in
    kernel order, then tcen, b, vel, RpRs, LD1, LD2 -- celerite's GP.get_parameter_vector order)."""
    params = np.asarray(params, dtype=np.float64)
    kfull = np.array(kfixed, dtype=np.float64).copy()
    nk = int(np.sum(free_mask))
    kfull[np.asarray(free_mask, dtype=bool)] = params[:nk]
    return kfull, params[nk:nk + 6]


def mono_logprob(params, lc, priors, kind, kfixed, free_mask, oversamp=10, dense=False):
    """MonoLogProb (namaste.py:1334-1339) for one walker vector."""
    kfull, mean = split_params(kind, params, kfixed, free_mask)
    coeffs = coefficients(kind, kfull)
    resid = lc[:, 1] - o.get_value(mean, lc[:, 0], oversamp)
    fn = dense_loglike if dense else celerite_loglike
    ll = fn(lc[:, 0], resid, lc[:, 2], coeffs)
    return ll + o.mono_ln_prior(params, priors)
