"""Thin Python objects over the C ABI: staged light curves and ensemble evaluations.

Everything numerical happens in namaste_b200/csrc (CUDA, sm_100a).  This module only
converts argument layouts; it never evaluates the model on the CPU."""
import ctypes

import numpy as np

from . import _lib
from ._lib import MODES, F64, F32, F64_STRICT, F32_PURE, as_f64, check, load, ptr

DTYPES = {"f64": F64, "f32": F32, F64: F64, F32: F32, "float64": F64, "float32": F32,
          "f64_strict": F64_STRICT, "strict": F64_STRICT, F64_STRICT: F64_STRICT,
          "f32_pure": F32_PURE, F32_PURE: F32_PURE}

PARAM_NAMES = ("tcen", "b", "vel", "RpRs", "LD1", "LD2")  # namaste/namaste.py:1153
_PRIOR_KINDS = {"gaussian": 1.0, "normlim": 2.0, "evans": 3.0}


def prior_table(priors, npar=6):
    """Ordered priors mapping (namaste.py:1062-1066, :479-480) -> [npar, 6] float table
    (lo, hi, kind, mu, sigma, pad).  Entry n applies to params[n] (namaste.py:1316).
    ``None`` or an empty mapping mean "no prior term" (quirk Q4 of SURVEY.md)."""
    if priors is None:
        return None
    if isinstance(priors, np.ndarray):
        tab = as_f64(priors)
        if tab.shape[-2:] != (npar, 6):
            raise ValueError("prior table must have shape [..., %d, 6]" % npar)
        return tab
    if len(priors) == 0:
        return None
    if len(priors) != npar:
        raise ValueError("priors has %d entries, expected %d (one per parameter, in order)" % (len(priors), npar))
    rows = []
    for key in priors.keys():
        row = priors[key]
        if len(row) > 2:
            if row[2] not in _PRIOR_KINDS:
                raise ValueError("unknown prior kind %r for %r" % (row[2], key))
            rows.append([row[0], row[1], _PRIOR_KINDS[row[2]], row[3], row[4] if len(row) > 4 else 0.0, 0.0])
        else:
            rows.append([row[0], row[1], 0.0, 0.0, 0.0, 0.0])
    return np.array(rows, dtype=np.float64)


class LightCurve:
    """One or more light curves staged on the current CUDA device (nmb_lc handle).

    ``lc`` is the reference's ``float64 [N, 3]`` array of (time, flux, flux_err)
    (namaste.py:1345 argument ``lc``), or a list of such arrays for a batch of
    independent candidates (ragged lengths allowed)."""

    def __init__(self, lc, oversamp=10):
        lib = load()
        single = isinstance(lc, np.ndarray) and lc.ndim == 2
        curves = [lc] if single else list(lc)
        if not curves:
            raise ValueError("no light curves given")
        curves = [as_f64(c) for c in curves]
        for c in curves:
            if c.ndim != 2 or c.shape[1] != 3:
                raise ValueError("each light curve must have shape [N, 3] (time, flux, flux_err)")
        n = np.array([len(c) for c in curves], dtype=np.int64)
        stride = int(n.max())
        t = np.zeros((len(curves), stride))
        f = np.ones((len(curves), stride))
        e = np.ones((len(curves), stride))
        for i, c in enumerate(curves):
            t[i, :n[i]], f[i, :n[i]], e[i, :n[i]] = c[:, 0], c[:, 1], c[:, 2]
        handle = ctypes.c_void_p()
        check(lib.nmb_lc_create_batch(ptr(t), ptr(f), ptr(e), ptr(n), len(curves), stride, int(oversamp),
                                      ctypes.byref(handle)), "LightCurve")
        self._h = handle
        self.ncand = len(curves)
        self.npoints = n
        self.oversamp = int(oversamp)
        self.cadence = np.array([lib.nmb_lc_cadence(handle, i) for i in range(self.ncand)])

    def close(self):
        if getattr(self, "_h", None):
            load().nmb_lc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer path (the call a reference user makes) --------------------------------
    def lnprob(self, params, priors=None, mode="windowed", return_loglike=False, dtype="f64", out=None):
        """Ensemble log-probability.  ``params`` is [W, 6] (single light curve) or
        [ncand, W, 6]; returns [W] / [ncand, W] (and the bare log-likelihood if asked).

        ``out``: optional float64 C-contiguous array of the result's shape to write into.  Buffers in page-locked
        host memory (e.g. ``torch.empty(..., pin_memory=True).numpy()``) -- ``params`` and / or ``out`` -- are copied
        to and from the device directly; ordinary arrays pass through the handle's pinned staging buffer."""
        lib = load()
        p = as_f64(params)
        squeeze = p.ndim == 2
        if squeeze:
            if self.ncand != 1:
                raise ValueError("params must be [ncand, W, 6] for a batched LightCurve")
            p = p[None]
        if p.ndim != 3 or p.shape[0] != self.ncand or p.shape[2] != 6:
            raise ValueError("params must have shape [W, 6] or [ncand, W, 6]")
        W = p.shape[1]
        tab = prior_table(priors)
        if tab is not None:
            if tab.size == self.ncand * 36:          # already one table per candidate: no copy
                tab = tab.reshape(self.ncand, 6, 6)
            else:
                tab = np.ascontiguousarray(np.broadcast_to(tab, (self.ncand, 6, 6)))
        if out is None:
            out = np.empty((self.ncand, W))
        else:
            if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous
                    and out.size == self.ncand * W):
                raise ValueError("out must be a C-contiguous float64 array with ncand * W elements")
            out = out.reshape(self.ncand, W)
        ll = np.empty((self.ncand, W)) if return_loglike else None
        check(lib.nmb_lnprob_host(self._h, ptr(p), W, ptr(tab), ptr(out), ptr(ll), MODES[mode], DTYPES[dtype]), "lnprob")
        if squeeze:
            out = out[0]
            ll = ll[0] if ll is not None else None
        return (out, ll) if return_loglike else out

    # ---- device-buffer path (torch tensors stay on the GPU) ---------------------------------
    def lnprob_device(self, params, priors=None, out=None, loglike=None, mode="windowed", stream=None,
                      dtype="f64"):
        """Asynchronous evaluation on torch CUDA tensors: params [ncand*W*6] float64,
        priors [ncand, 6, 6] float64 tensor or None.  Returns the output tensor."""
        import torch
        lib = load()
        if not (params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()):
            raise ValueError("params must be a contiguous float64 CUDA tensor")
        W = params.numel() // (6 * self.ncand)
        if out is None:
            out = torch.empty((self.ncand, W), dtype=torch.float64, device=params.device)
        if stream is None:
            stream = torch.cuda.current_stream(params.device).cuda_stream
        check(lib.nmb_lnprob(self._h, ptr(params), W, ptr(priors), ptr(out), ptr(loglike), MODES[mode], DTYPES[dtype],
                             ctypes.c_void_p(stream)), "lnprob_device")
        return out

    def last_kernels(self):
        """Names of the (stage-1, stage-2) kernels the last evaluation launched."""
        a, b = ctypes.c_char_p(), ctypes.c_char_p()
        check(load().nmb_last_kernels(self._h, ctypes.byref(a), ctypes.byref(b)), "last_kernels")
        return a.value.decode(), b.value.decode()

    def stage_timing(self, enable=True):
        """Switch per-kernel device timing of the two pipeline stages on or off (resets the means)."""
        check(load().nmb_stage_timing(self._h, 1 if enable else 0), "stage_timing")

    def stage_times(self):
        """(stage-1 ms, stage-2 ms, evaluations) averaged since the last call; needs stage_timing()."""
        a, b, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        check(load().nmb_stage_times(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)), "stage_times")
        return a.value, b.value, int(n.value)

    def model(self, params, cand=0):
        """MonotransitModel.get_value for W parameter rows -> [W, N] (host arrays)."""
        import torch
        lib = load()
        p = np.atleast_2d(as_f64(params))
        W, n = p.shape[0], int(self.npoints[cand])
        dp = torch.from_numpy(p).cuda()
        dm = torch.empty((W, n), dtype=torch.float64, device="cuda")
        check(lib.nmb_model(self._h, cand, ptr(dp), W, ptr(dm), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "model")
        return dm.cpu().numpy()


def occultquad(z, p0, gamma, retall=False):
    """Crossfield_transit.occultquad(z, p0, gamma, retall) evaluated on the GPU."""
    import torch
    lib = load()
    z = as_f64(z)
    shape = z.shape
    zf = z.reshape(-1)
    gamma = np.ravel(np.asarray(gamma, dtype=float))
    g1, g2 = (float(gamma[0]), float(gamma[1])) if gamma.size >= 2 else (float(gamma[0]), 0.0)
    dz = torch.from_numpy(zf).cuda()
    out = torch.empty((4, zf.size), dtype=torch.float64, device="cuda")
    check(lib.nmb_occultquad(ptr(dz), zf.size, float(p0), g1, g2, ptr(out),
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "occultquad")
    res = out.cpu().numpy()
    if retall:
        return tuple(res[i].reshape(shape) for i in range(4))
    return res[0].reshape(shape)


def lnprior(params, priors):
    """MonoLnPrior for rows of ``params`` ([npar] or [W, npar]) on the GPU."""
    import torch
    lib = load()
    p = as_f64(params)
    scalar = p.ndim == 1
    p = np.atleast_2d(p)
    tab = prior_table(priors, npar=p.shape[1])
    if tab is None:
        return 0.0 if scalar else np.zeros(len(p))
    dp, dt = torch.from_numpy(p).cuda(), torch.from_numpy(tab).cuda()
    out = torch.empty(len(p), dtype=torch.float64, device="cuda")
    check(lib.nmb_lnprior(ptr(dp), len(p), p.shape[1], ptr(dt), ptr(out),
                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "lnprior")
    res = out.cpu().numpy()
    return float(res[0]) if scalar else res


def measure_fp64_peak():
    """(TFLOP/s, ms) of a pure DFMA kernel on the current device."""
    lib = load()
    tf, ms = ctypes.c_double(), ctypes.c_double()
    check(lib.nmb_measure_fp64_peak(ctypes.byref(tf), ctypes.byref(ms)), "measure_fp64_peak")
    return tf.value, ms.value


def launch_count():
    return int(load().nmb_launch_count())
