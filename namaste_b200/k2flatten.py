"""``k2flatten.ReduceNoise`` on the GPU (namaste/k2flatten.py:51-84; windows :19-40, fit :5-17).

Namaste detrends every light curve with it before fitting (namaste.py:442, Example.ipynb cell 45:
``ReduceNoise(lc, winsize=4.5, stepsize=0.125)``).  The window bookkeeping (a handful of
``searchsorted`` calls per step) stays on the host; the clipped polynomial fits -- one per step, 21
weighted least-squares solves each -- run one thread block per step in csrc/flatten_kernels.cuh.
"""
import ctypes

import numpy as np

from ._lib import as_f64, check, load, ptr

__all__ = ["ReduceNoise", "step_windows"]


def step_windows(t, winsize, stepsize, gapthreshold):
    """Per detrending step: window = [wlo, blo) U [bhi, whi) and box = [blo, bhi) as index ranges into
    the (zero-based) time array, following formwindow (k2flatten.py:19-40): the window is winsize wide
    around the step centre; if it runs into a gap wider than gapthreshold on one side only it is
    shifted to start (or end) at the gap, keeping its width."""
    n = len(t)
    lenlc = t[-1]
    nsteps = int(np.ceil(lenlc / stepsize))
    cents = np.arange(nsteps) / float(nsteps) * lenlc + stepsize / 2.
    rows = np.empty((nsteps, 4), dtype=np.int64)
    for s, cent in enumerate(cents):
        lo_edge, hi_edge = cent - winsize / 2., cent + winsize / 2.
        wlo, whi = np.searchsorted(t, lo_edge), np.searchsorted(t, hi_edge)
        blo, bhi = np.searchsorted(t, cent - stepsize / 2.), np.searchsorted(t, cent + stepsize / 2.)
        if whi == n:
            whi -= 1
        gap_above = t[whi] < hi_edge - gapthreshold
        gap_below = t[wlo] > lo_edge + gapthreshold
        if gap_above and not gap_below:
            wlo = np.searchsorted(t, t[whi] - winsize)
        elif gap_below and not gap_above:
            whi = np.searchsorted(t, t[wlo] + winsize)
        rows[s] = (wlo, blo, bhi, whi)
    return rows


def _owned(rows):
    """Part of each box that survives the reference's sequential loop: where boxes overlap, the later
    step overwrites the earlier one."""
    nsteps = len(rows)
    own = np.empty((nsteps, 2), dtype=np.int64)
    nxt = np.iinfo(np.int64).max
    # boxes are ordered (blo and bhi non-decreasing), so a point of box s is overwritten iff it lies
    # at or beyond the smallest blo among the later, non-empty boxes (and inside that box, which the
    # ordering guarantees up to that box's bhi; anything beyond is handled by even later boxes)
    for s in range(nsteps - 1, -1, -1):
        blo, bhi = rows[s, 1], rows[s, 2]
        own[s] = (blo, max(blo, min(bhi, nxt)))
        if bhi > blo:
            nxt = min(nxt, blo)
    return own


def ReduceNoise(lc2, winsize=2, stepsize=0.2, polydegree=3, niter=20, sigmaclip=3., gapthreshold=1.0):
    """Detrended copy of ``lc2`` ([N, 3] time, flux, flux_err): each ``stepsize`` box has the clipped
    degree-``polydegree`` polynomial fitted to its surrounding ``winsize`` window subtracted and 1 added.
    Same signature, defaults and output layout (time, detrended flux, original flux_err) as the
    reference; points that no box covers stay 0, as there."""
    lc = np.array(lc2, dtype=np.float64, copy=True)
    if lc.ndim != 2 or lc.shape[1] < 3:
        raise ValueError("lc must have shape [N, 3] (time, flux, flux_err)")
    out = np.column_stack((lc[:, 0], np.zeros(len(lc)), lc[:, 2]))
    t = lc[:, 0] - lc[0, 0]
    lcbase = np.median(lc[:, 1])
    f = lc[:, 1] / lcbase
    e = lc[:, 2] / lcbase
    rows = step_windows(t, winsize, stepsize, gapthreshold)
    own = _owned(rows)
    steps = np.ascontiguousarray(np.hstack([rows, own]), dtype=np.int32)
    flux = np.zeros(len(lc))
    check(load().nmb_flatten(ptr(as_f64(t)), ptr(as_f64(f)), ptr(as_f64(e)), len(t), ptr(steps), len(steps),
                             float(winsize), int(polydegree), int(niter), float(sigmaclip), ptr(flux)), "ReduceNoise")
    out[:, 1] = flux
    return out
