"""Host-side mirror of the reference's interface for the likelihood hot path.

Same names, argument meaning and error behaviour as ``namaste/namaste.py`` so that a
Namaste user can switch imports; all arithmetic runs in the CUDA library.

  MonotransitModel   <- namaste/namaste.py:1152-1164 (celerite.modeling.Model duck type)
  MonoLnPrior        <- namaste/namaste.py:1314-1332
  MonoOnlyLogProb    <- namaste/namaste.py:1345-1349   (also accepts an ensemble [W, 6])
  MonoOnlyNegLogProb <- namaste/namaste.py:1351-1352
  CalcTdur / CalcVel <- namaste/namaste.py:1469-1477
"""
from collections import OrderedDict

import numpy as np

from . import core

__all__ = ["MonotransitModel", "MonoLnPrior", "MonoOnlyLogProb", "MonoOnlyNegLogProb", "CalcTdur", "CalcVel",
           "stage_lightcurve", "clear_cache"]

_LC_CACHE = OrderedDict()
_LC_CACHE_MAX = 8


def _fingerprint(a):
    """Cheap identity + content probe of an array (address, layout, strided sample)."""
    flat = a.reshape(-1)
    step = max(1, flat.size // 97)
    return (a.__array_interface__["data"][0], a.shape, a.strides, float(flat[::step].sum()), float(flat[-1]))


def stage_lightcurve(lc, oversamp=10):
    """Return the device-staged LightCurve for ``lc`` (float64 [N, 3]), staging it once.

    The reference passes the same ``lc`` array to every log-prob call (emcee ``args=``,
    namaste.py:590); here it is copied to HBM on first use and found again by identity."""
    lc = np.asarray(lc)
    key = (_fingerprint(lc), int(oversamp))
    hit = _LC_CACHE.get(key)
    if hit is not None:
        _LC_CACHE.move_to_end(key)
        return hit
    staged = core.LightCurve(np.ascontiguousarray(lc, dtype=np.float64), oversamp=oversamp)
    _LC_CACHE[key] = staged
    while len(_LC_CACHE) > _LC_CACHE_MAX:
        _, old = _LC_CACHE.popitem(last=False)
        old.close()
    return staged


def clear_cache():
    while _LC_CACHE:
        _, old = _LC_CACHE.popitem()
        old.close()


class MonotransitModel(object):
    """Six-parameter single-transit model with the surface of celerite.modeling.Model
    that Namaste uses (namaste.py:486-491, :550-553, :1016-1021, :1346, :1482-1505)."""
    parameter_names = ("tcen", "b", "vel", "RpRs", "LD1", "LD2")
    oversamp = 10  # hard-coded in the reference (namaste.py:1158)

    def __init__(self, *args, **kwargs):
        names = self.parameter_names
        if args and kwargs:
            raise ValueError("parameters must be given either positionally or by keyword, not both")
        if args:
            if len(args) != len(names):
                raise ValueError("expected %d parameters" % len(names))
            vec = list(args)
        else:
            missing = [nm for nm in names if nm not in kwargs]
            if missing:
                raise ValueError("missing parameter(s): %s" % ", ".join(missing))
            vec = [kwargs[nm] for nm in names]
        object.__setattr__(self, "parameter_vector", np.array(vec, dtype=np.float64))

    # attribute access to parameters, like celerite's Model.__getattr__
    def __getattr__(self, name):
        names = object.__getattribute__(self, "parameter_names")
        if name in names:
            return object.__getattribute__(self, "parameter_vector")[names.index(name)]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self.parameter_names:
            self.parameter_vector[self.parameter_names.index(name)] = value
        else:
            object.__setattr__(self, name, value)

    def __len__(self):
        return len(self.parameter_names)

    @property
    def full_size(self):
        return len(self.parameter_names)

    @property
    def vector_size(self):
        return len(self.parameter_names)

    def get_parameter_names(self, include_frozen=False):
        return tuple(self.parameter_names)

    def get_parameter_vector(self, include_frozen=False):
        return self.parameter_vector.copy()

    def set_parameter_vector(self, vector, include_frozen=False):
        v = np.asarray(vector, dtype=np.float64)
        if v.shape != self.parameter_vector.shape:
            raise ValueError("dimension mismatch")
        self.parameter_vector[:] = v

    def get_parameter_dict(self, include_frozen=False):
        return OrderedDict(zip(self.parameter_names, self.parameter_vector))

    def get_parameter(self, name):
        return self.parameter_vector[self.parameter_names.index(name)]

    def set_parameter(self, name, value):
        self.parameter_vector[self.parameter_names.index(name)] = value

    def get_value(self, t):
        """Binned, supersampled transit model at times ``t`` (namaste.py:1155-1164)."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        lc = np.column_stack((t, np.ones_like(t), np.ones_like(t)))
        staged = core.LightCurve(lc, oversamp=self.oversamp)
        try:
            return staged.model(self.parameter_vector[None, :])[0]
        finally:
            staged.close()


def MonoLnPrior(params, priors):
    """Prior term of the log-probability (namaste.py:1314-1332), evaluated on the GPU."""
    return core.lnprior(params, priors)


def MonoOnlyLogProb(params, lc, priors, monomodel=None, mode="windowed", dtype="f64"):
    """Gaussian (no-GP) log-probability of the single-transit model (namaste.py:1345-1349).

    ``params`` may be one vector [6] (returns a float, exactly the reference's contract) or
    an ensemble [W, 6] (returns [W]).  NaN propagates; out-of-bounds parameters give a
    hugely negative value, never -inf (soft walls, namaste.py:1318-1320)."""
    p = np.asarray(params, dtype=np.float64)
    oversamp = getattr(monomodel, "oversamp", 10) if monomodel is not None else 10
    staged = stage_lightcurve(lc, oversamp)
    if p.ndim == 1:
        if monomodel is not None:
            monomodel.set_parameter_vector(p)  # the reference mutates the model (namaste.py:1346)
        return float(staged.lnprob(p[None, :], priors, mode=mode, dtype=dtype)[0])
    return staged.lnprob(p, priors, mode=mode, dtype=dtype)


def MonoOnlyNegLogProb(params, lc, priors, monomodel=None, mode="windowed", dtype="f64"):
    return -1 * MonoOnlyLogProb(params, lc, priors, monomodel, mode=mode, dtype=dtype)


def CalcTdur(vel, b, p):
    """Transit duration from velocity [Rs/day], impact parameter, radius ratio (namaste.py:1469-1472)."""
    return (2 * (1 + p) * np.sqrt(1 - (b / (1 + p)) ** 2)) / vel


def CalcVel(Tdur, b, p):
    """Velocity [Rs/day] from transit duration (namaste.py:1474-1477)."""
    return (2 * (1 + p) * np.sqrt(1 - (b / (1 + p)) ** 2)) / Tdur
