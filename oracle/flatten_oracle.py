"""CPU oracle of the detrending row: ``k2flatten.ReduceNoise`` restated (namaste/k2flatten.py:5-17 dopolyfit,
:19-40 formwindow, :51-84 ReduceNoise), NumPy ``polyfit`` like the reference.

TEST INFRASTRUCTURE ONLY: the checker of tests/ and the CPU baseline leg of bench.py.  Parity pinned:
tests/test_oracle_golden.py compares it with tests/golden/flatten_epic203311200.npz, vectors produced by the
reference's own k2flatten.py (oracle/gen_golden_flatten.py)."""
import warnings

import numpy as np


def clipped_polyfit(win, deg, niter, sigclip):
    """k2flatten.py:5-17: weighted polyfit, then `niter` rounds of clip-and-refit."""
    coef = np.polyfit(win[:, 0], win[:, 1], w=1.0 / np.power(win[:, 2], 2), deg=deg)
    for _ in range(niter):
        offset = np.abs(win[:, 1] - np.polyval(coef, win[:, 0])) / win[:, 2]
        inside = offset < sigclip
        if inside.sum() > int(0.8 * len(win[:, 0])):
            sel = win[inside, :]
        else:
            sel = win[offset < np.average(offset)]
        coef = np.polyfit(sel[:, 0], sel[:, 1], w=1.0 / np.power(sel[:, 2], 2), deg=deg)
    return coef


def form_window(dat, cent, size, boxsize, gapthresh):
    """k2flatten.py:19-40: the fitting window around a step (box excluded), shifted off a gap on one side."""
    t = dat[:, 0]
    wlo, whi = np.searchsorted(t, cent - size / 2.), np.searchsorted(t, cent + size / 2.)
    blo, bhi = np.searchsorted(t, cent - boxsize / 2.), np.searchsorted(t, cent + boxsize / 2.)
    if whi == len(t):
        whi -= 1
    highgap = t[whi] < (cent + size / 2.) - gapthresh
    lowgap = t[wlo] > (cent - size / 2.) + gapthresh
    if highgap and not lowgap:
        wlo = np.searchsorted(t, t[whi] - size)
    elif lowgap and not highgap:
        whi = np.searchsorted(t, t[wlo] + size)
    return np.concatenate((dat[wlo:blo, :], dat[bhi:whi, :])), blo, bhi


def reduce_noise(lc2, winsize=2, stepsize=0.2, polydegree=3, niter=20, sigmaclip=3., gapthreshold=1.0):
    """k2flatten.py:51-84."""
    lc = np.copy(lc2)
    out = np.column_stack((lc[:, 0], np.zeros(len(lc[:, 0])), lc[:, 2]))
    lc[:, 0] = lc[:, 0] - lc[0, 0]
    lenlc = lc[-1, 0]
    base = np.median(lc[:, 1])
    lc[:, 1] /= base
    lc[:, 2] = lc[:, 2] / base
    nsteps = np.ceil(lenlc / stepsize).astype('int')
    cents = np.arange(nsteps) / float(nsteps) * lenlc + stepsize / 2.
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # polyfit's RankWarning on short windows, as in the reference
        for s in range(nsteps):
            win, blo, bhi = form_window(lc, cents[s], winsize, stepsize, gapthreshold)
            coef = clipped_polyfit(win, polydegree, niter, sigmaclip)
            out[blo:bhi, 1] = lc[blo:bhi, 1] - np.polyval(coef, lc[blo:bhi, 0]) + 1
    return out
