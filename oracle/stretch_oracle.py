"""CPU oracle of the ensemble sampler: the emcee 2.x stretch move driven by the SAME counter-based
random stream (Philox4x32-10) as the device sampler, with the NumPy oracle as log-probability.

TEST INFRASTRUCTURE ONLY.  emcee is a third-party dependency of the reference that is not vendored
(no pin; the API used at namaste/namaste.py:583-608 is emcee 2.x) and is not installed here, so the
move is restated from the published algorithm (Goodman & Weare 2010; Foreman-Mackey et al. 2013,
Alg. 3) -- "parity unpinned": no reference output fixes chain values, only distributions.  What this
oracle pins is the device implementation: same seed => same proposals, same accept decisions.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011); counters are uint32 arrays, key two ints."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u01_53(hi, lo):
    x = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return x.astype(np.float64) * 1.1102230246251565e-16


def draws(seed, step, half, walkers, cand=0):
    """(u_z, u_j, u_a) for the global walker indices `walkers` of one half-step."""
    walkers = np.asarray(walkers, dtype=np.uint64)
    z = np.zeros_like(walkers)
    c0, c1 = z + np.uint64(step & 0xFFFFFFFF), z + np.uint64((step >> 32) & 0xFFFFFFFF)
    c2 = ((walkers << np.uint64(1)) | np.uint64(half)) & MASK
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    r = philox4x32_10(c0, c1, c2, z + np.uint64((cand << 1) & 0xFFFFFFFF), k0, k1)
    r2 = philox4x32_10(c0, c1, c2, z + np.uint64(((cand << 1) | 1) & 0xFFFFFFFF), k0, k1)
    return u01_53(r[0], r[1]), u01_53(r[2], r[3]), u01_53(r2[0], r2[1])


def run(lnprob_fn, pos0, nsteps, seed, a=2.0, step0=0, lnprob0=None):
    """Stretch-move chain; returns chain [W, nsteps, dim], lnprobability [W, nsteps], naccepted."""
    p = np.array(pos0, dtype=float)
    W, dim = p.shape
    halfk = W // 2
    lnp = np.array(lnprob_fn(p), dtype=float) if lnprob0 is None else np.array(lnprob0, dtype=float)
    if np.any(np.isnan(lnp)):
        raise ValueError("The initial lnprob was NaN.")
    chain = np.zeros((W, nsteps, dim))
    lnchain = np.zeros((W, nsteps))
    nacc = np.zeros(W, dtype=np.int64)
    for it in range(nsteps):
        step = step0 + it
        for half, (S0, S1) in enumerate([(slice(0, halfk), slice(halfk, W)), (slice(halfk, W), slice(0, halfk))]):
            idx = np.arange(W)[S0]
            s, c = p[S0], p[S1]
            uz, uj, ua = draws(seed, step, half, idx)
            zz = ((a - 1.) * uz + 1) ** 2. / a
            j = np.minimum((uj * len(c)).astype(int), len(c) - 1)
            q = c[j] - zz[:, np.newaxis] * (c[j] - s)
            newlnp = np.array(lnprob_fn(q), dtype=float)
            if np.any(np.isnan(newlnp)):
                raise ValueError("lnprob returned NaN.")
            with np.errstate(divide="ignore"):
                lnpdiff = (dim - 1.) * np.log(zz) + newlnp - lnp[S0]
                accept = lnpdiff > np.log(ua)
            p[idx[accept]] = q[accept]
            lnp[idx[accept]] = newlnp[accept]
            nacc[idx[accept]] += 1
        chain[:, it, :] = p
        lnchain[:, it] = lnp
    return chain, lnchain, nacc
