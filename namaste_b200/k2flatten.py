"""``k2flatten.ReduceNoise`` on the GPU (namaste/k2flatten.py:51-84; windows :19-40, fit :5-17), for one light curve
or a batch of candidates.

Namaste detrends every light curve with it before fitting (namaste.py:1143, Example.ipynb cell 45:
``ReduceNoise(lc, winsize=4.5, stepsize=0.125)``, twice).  Everything per point and per step runs on the device:
the window / box bookkeeping of ``formwindow`` (one thread per detrending step) and the clipped polynomial fits
(one thread block per step, 21 weighted least-squares solves each) -- csrc/flatten_kernels.cuh.  The host only
shifts the times, divides by the median flux (one ``np.median`` per light curve, as the reference does once per
star) and pads ragged candidates.  The device buffers live in one workspace handle that is kept between calls.
"""
import ctypes

import numpy as np

from ._lib import as_f64, check, load, ptr

__all__ = ["ReduceNoise", "ReduceNoiseBatch", "AnomCutDiffBatch", "window_mask_batch", "Workspace"]


class Workspace(object):
    """Device buffers and stream of the batched pre-fit steps (nmb_flat handle), grown on demand."""

    def __init__(self):
        h = ctypes.c_void_p()
        check(load().nmb_flat_create(ctypes.byref(h)), "Workspace")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            load().nmb_flat_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ws = None


def _workspace(ws):
    global _default_ws
    if ws is not None:
        return ws
    if _default_ws is None or not _default_ws._h:
        _default_ws = Workspace()
    return _default_ws


def _pad(columns, fill):
    """List of 1-D arrays -> ([ncand, stride] array, lengths)."""
    n = np.array([len(c) for c in columns], dtype=np.int64)
    out = np.full((len(columns), int(n.max())), fill, dtype=np.float64)
    for i, c in enumerate(columns):
        out[i, :n[i]] = c
    return out, n


def ReduceNoiseBatch(lcs, winsize=2, stepsize=0.2, polydegree=3, niter=20, sigmaclip=3., gapthreshold=1.0, workspace=None):
    """``ReduceNoise`` for a list of light curves ([N_c, 3] each, ragged lengths allowed) in ONE device call.
    Returns the list of detrended copies (time, detrended flux, original flux_err); candidate c of the batch has
    the bits ``ReduceNoise(lcs[c], ...)`` gives on its own."""
    lcs = [np.asarray(lc, dtype=np.float64) for lc in lcs]
    for lc in lcs:
        if lc.ndim != 2 or lc.shape[1] < 3:
            raise ValueError("every light curve must have shape [N, 3] (time, flux, flux_err)")
    base = [np.median(lc[:, 1]) for lc in lcs]                      # lcbase (k2flatten.py:62)
    t, n = _pad([lc[:, 0] - lc[0, 0] for lc in lcs], 0.0)           # JD0 shift (:59-60)
    f, _ = _pad([lc[:, 1] / b for lc, b in zip(lcs, base)], 1.0)
    e, _ = _pad([lc[:, 2] / b for lc, b in zip(lcs, base)], 1.0)
    out = np.zeros_like(t)
    ws = _workspace(workspace)
    check(load().nmb_flatten_batch(ws._h, ptr(t), ptr(f), ptr(e), ptr(n), len(lcs), t.shape[1], float(winsize),
                                   float(stepsize), float(gapthreshold), int(polydegree), int(niter), float(sigmaclip),
                                   ptr(out)), "ReduceNoise")
    return [np.column_stack((lc[:, 0], out[i, :n[i]], lc[:, 2])) for i, lc in enumerate(lcs)]


def ReduceNoise(lc2, winsize=2, stepsize=0.2, polydegree=3, niter=20, sigmaclip=3., gapthreshold=1.0):
    """Detrended copy of ``lc2`` ([N, 3] time, flux, flux_err): each ``stepsize`` box has the clipped
    degree-``polydegree`` polynomial fitted to its surrounding ``winsize`` window subtracted and 1 added.
    Same signature, defaults and output layout (time, detrended flux, original flux_err) as the
    reference; points that no box covers stay 0, as there."""
    lc = np.asarray(lc2, dtype=np.float64)
    if lc.ndim != 2 or lc.shape[1] < 3:
        raise ValueError("lc must have shape [N, 3] (time, flux, flux_err)")
    return ReduceNoiseBatch([lc], winsize, stepsize, polydegree, niter, sigmaclip, gapthreshold)[0]


def AnomCutDiffBatch(fluxes, thresh=4.2, workspace=None):
    """``AnomCutDiff`` (namaste.py:1459-1467) for a list of NaN-free flux series in one device call: list of boolean
    masks of the points to KEEP."""
    cols = [as_f64(f).reshape(-1) for f in fluxes]
    flux, n = _pad(cols, 1.0)
    keep = np.zeros(flux.shape, dtype=np.uint8)
    ws = _workspace(workspace)
    check(load().nmb_anomcut_batch(ws._h, ptr(flux), ptr(n), len(cols), flux.shape[1], float(thresh), ptr(keep)), "AnomCutDiff")
    return [keep[i, :n[i]].astype(bool) for i in range(len(cols))]


def window_mask_batch(times, tcen, width, two_sided=False, workspace=None):
    """The fit window of ``RunMCMC`` (namaste.py:550-556) for a list of time arrays: ``abs(t - tcen < width)`` -- the
    reference's quirk, which keeps everything before ``tcen + width`` -- or ``|t - tcen| < width``."""
    cols = [as_f64(t).reshape(-1) for t in times]
    t, n = _pad(cols, 0.0)
    keep = np.zeros(t.shape, dtype=np.uint8)
    ws = _workspace(workspace)
    tc, wd = as_f64(np.broadcast_to(tcen, len(cols))), as_f64(np.broadcast_to(width, len(cols)))
    check(load().nmb_window_mask_batch(ws._h, ptr(t), ptr(n), len(cols), t.shape[1], ptr(tc), ptr(wd),
                                       1 if two_sided else 0, ptr(keep)), "window_mask")
    return [keep[i, :n[i]].astype(bool) for i in range(len(cols))]
