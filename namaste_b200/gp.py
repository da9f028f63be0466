"""GP noise-model objective of Namaste on the GPU: ``MonoLogProb`` and the celerite duck types it needs.

Mirrors the reference's use of celerite (namaste/namaste.py:279-314 kernels, :583 sampler registration,
:1334-1342 MonoLogProb / MonoNegLogProb, :1355-1379 RotationTerm):

    kern = RealTerm(log_a=..., log_c=..., bounds=...) + JitterTerm(log_sigma, bounds=...)
    kern = RotationTerm(log_amp, log_timescale, log_period, log_factor, bounds=...) + JitterTerm(...)
    kern.freeze_parameter('terms[1]:log_sigma')
    gp = GP(kernel=kern, mean=MonotransitModel(...), fit_mean=True)
    MonoLogProb(params, lc, priors, gp)        # params = gp.get_parameter_vector() layout

Only these two kernels exist in Namaste, and only they are implemented (anything else raises
``TypeError``).  The solver is celerite's published semiseparable Cholesky, run one walker per thread
(csrc/gp_kernels.cuh); celerite itself is absent from the reference tree, see oracle/gp_oracle.py.
"""
from collections import OrderedDict

import numpy as np

from . import core
from ._lib import as_f64, check, load, ptr

__all__ = ["RealTerm", "JitterTerm", "RotationTerm", "GP", "MonoLogProb", "MonoNegLogProb"]

NMB_GP_REAL, NMB_GP_QUASI = 0, 1


class _Parametrised(object):
    parameter_names = ()

    def _init_params(self, values, bounds):
        self._names = list(self.parameter_names)
        self._vector = np.array(values, dtype=np.float64)
        self._frozen = np.zeros(len(self._names), dtype=bool)
        b = bounds or {}
        if isinstance(b, dict):
            self._bounds = [tuple(b.get(nm, (None, None))) for nm in self._names]
        else:
            self._bounds = [tuple(x) for x in b]

    # celerite.modeling.Model surface used by Namaste
    def get_parameter_names(self, include_frozen=False):
        return tuple(nm for nm, fz in zip(self._names, self._frozen) if include_frozen or not fz)

    def get_parameter_vector(self, include_frozen=False):
        return self._vector.copy() if include_frozen else self._vector[~self._frozen].copy()

    def set_parameter_vector(self, vector, include_frozen=False):
        v = np.asarray(vector, dtype=np.float64)
        if include_frozen:
            if v.shape != self._vector.shape:
                raise ValueError("dimension mismatch")
            self._vector[:] = v
        else:
            if v.shape != (int((~self._frozen).sum()),):
                raise ValueError("dimension mismatch")
            self._vector[~self._frozen] = v

    def get_parameter_bounds(self, include_frozen=False):
        return [b for b, fz in zip(self._bounds, self._frozen) if include_frozen or not fz]

    def get_parameter_dict(self, include_frozen=False):
        return OrderedDict(zip(self.get_parameter_names(include_frozen), self.get_parameter_vector(include_frozen)))

    def get_parameter(self, name):
        return self._vector[self._names.index(name)]

    def set_parameter(self, name, value):
        self._vector[self._names.index(name)] = value

    def freeze_parameter(self, name):
        self._frozen[self._names.index(name)] = True

    def thaw_parameter(self, name):
        self._frozen[self._names.index(name)] = False

    def __len__(self):
        return int((~self._frozen).sum())


class Term(_Parametrised):
    def __init__(self, *args, **kwargs):
        bounds = kwargs.pop("bounds", None)
        names = self.parameter_names
        if args:
            if len(args) != len(names):
                raise ValueError("expected %d parameters" % len(names))
            vals = list(args)
        else:
            vals = [kwargs[nm] for nm in names]
        self._init_params(vals, bounds)

    def __add__(self, other):
        return TermSum(self, other)


class RealTerm(Term):
    """k(tau) = exp(log_a) exp(-exp(log_c) tau)   (celerite.terms.RealTerm)"""
    parameter_names = ("log_a", "log_c")


class JitterTerm(Term):
    """k(tau) = exp(2 log_sigma) delta(tau)   (celerite.terms.JitterTerm)"""
    parameter_names = ("log_sigma",)


class RotationTerm(Term):
    """Namaste's quasi-periodic term (namaste.py:1355-1379): one real and one complex celerite term."""
    parameter_names = ("log_amp", "log_timescale", "log_period", "log_factor")


class TermSum(_Parametrised):
    def __init__(self, *terms):
        self.terms = list(terms)
        self._names = ["terms[%d]:%s" % (i, nm) for i, t in enumerate(self.terms) for nm in t._names]

    # views onto the member terms
    @property
    def _vector(self):
        return _Joined(self.terms, "_vector")

    @property
    def _frozen(self):
        return _Joined(self.terms, "_frozen")

    @property
    def _bounds(self):
        return [b for t in self.terms for b in t._bounds]

    def get_parameter_vector(self, include_frozen=False):
        v = np.concatenate([t._vector for t in self.terms])
        fz = np.concatenate([t._frozen for t in self.terms])
        return v.copy() if include_frozen else v[~fz]

    def set_parameter_vector(self, vector, include_frozen=False):
        v = np.asarray(vector, dtype=np.float64)
        at = 0
        for t in self.terms:
            n = len(t._names) if include_frozen else len(t)
            t.set_parameter_vector(v[at:at + n], include_frozen)
            at += n
        if at != len(v):
            raise ValueError("dimension mismatch")

    def get_parameter_names(self, include_frozen=False):
        fz = np.concatenate([t._frozen for t in self.terms])
        return tuple(nm for nm, f in zip(self._names, fz) if include_frozen or not f)

    def _locate(self, name):
        i = self._names.index(name)
        for t in self.terms:
            if i < len(t._names):
                return t, t._names[i]
            i -= len(t._names)
        raise ValueError(name)

    def get_parameter(self, name):
        t, nm = self._locate(name)
        return t.get_parameter(nm)

    def set_parameter(self, name, value):
        t, nm = self._locate(name)
        t.set_parameter(nm, value)

    def freeze_parameter(self, name):
        t, nm = self._locate(name)
        t.freeze_parameter(nm)

    def thaw_parameter(self, name):
        t, nm = self._locate(name)
        t.thaw_parameter(nm)

    def __len__(self):
        return sum(len(t) for t in self.terms)

    def __add__(self, other):
        return TermSum(*(self.terms + [other]))


class _Joined(object):  # only what _Parametrised's generic getters need from a concatenated view
    def __init__(self, terms, attr):
        self._arr = np.concatenate([getattr(t, attr) for t in terms])

    def __getitem__(self, k):
        return self._arr[k]

    def copy(self):
        return self._arr.copy()


def _kernel_spec(kernel):
    """(kind, full kernel vector, free mask) for the two kernels Namaste builds."""
    if isinstance(kernel, TermSum) and len(kernel.terms) == 2 and isinstance(kernel.terms[1], JitterTerm):
        head = kernel.terms[0]
        if isinstance(head, RealTerm):
            kind = NMB_GP_REAL
        elif isinstance(head, RotationTerm):
            kind = NMB_GP_QUASI
        else:
            kind = None
        if kind is not None:
            full = kernel.get_parameter_vector(include_frozen=True)
            frozen = np.concatenate([t._frozen for t in kernel.terms])
            mask = 0
            for i, fz in enumerate(frozen):
                if not fz:
                    mask |= 1 << i
            return kind, full, mask
    raise TypeError("namaste_b200 implements the two GP kernels Namaste uses: RealTerm + JitterTerm and "
                    "RotationTerm + JitterTerm (namaste.py:279-314)")


class GP(object):
    """celerite.GP as Namaste drives it: parameter vector = free kernel parameters, then (with
    ``fit_mean=True``) the mean model's; ``compute(t, yerr)`` + ``log_likelihood(y)``."""

    def __init__(self, kernel, mean=None, fit_mean=False):
        _kernel_spec(kernel)
        self.kernel, self.mean, self.fit_mean = kernel, mean, fit_mean
        self._t = self._yerr = None

    def get_parameter_names(self, include_frozen=False):
        names = tuple("kernel:" + nm for nm in self.kernel.get_parameter_names(include_frozen))
        if self.fit_mean or include_frozen:
            names += tuple("mean:" + nm for nm in self.mean.get_parameter_names())
        return names

    def get_parameter_vector(self, include_frozen=False):
        v = self.kernel.get_parameter_vector(include_frozen)
        if self.fit_mean or include_frozen:
            v = np.concatenate([v, self.mean.get_parameter_vector()])
        return v

    def set_parameter_vector(self, vector, include_frozen=False):
        v = np.asarray(vector, dtype=np.float64)
        nk = len(self.kernel.get_parameter_vector(include_frozen))
        self.kernel.set_parameter_vector(v[:nk], include_frozen)
        if self.fit_mean or include_frozen:
            self.mean.set_parameter_vector(v[nk:])
        elif len(v) != nk:
            raise ValueError("dimension mismatch")

    def get_parameter(self, name):
        part, nm = name.split(":", 1)
        return (self.kernel if part == "kernel" else self.mean).get_parameter(nm)

    def set_parameter(self, name, value):
        part, nm = name.split(":", 1)
        (self.kernel if part == "kernel" else self.mean).set_parameter(nm, value)

    def freeze_parameter(self, name):
        part, nm = name.split(":", 1)
        if part != "kernel":
            raise ValueError("only kernel parameters can be frozen here")
        self.kernel.freeze_parameter(nm)

    def get_parameter_dict(self, include_frozen=False):
        return OrderedDict(zip(self.get_parameter_names(include_frozen), self.get_parameter_vector(include_frozen)))

    def compute(self, t, yerr=1.123e-12, **kwargs):
        t = np.ascontiguousarray(t, dtype=np.float64)
        self._t = t
        self._yerr = np.broadcast_to(np.asarray(yerr, dtype=np.float64), t.shape).copy()

    def log_likelihood(self, y, **kwargs):
        if self._t is None:
            raise RuntimeError("You need to compute the model first")
        lc = np.column_stack((self._t, np.asarray(y, dtype=np.float64), self._yerr))
        vec = np.concatenate([self.kernel.get_parameter_vector(), self.mean.get_parameter_vector()])
        return float(_gp_eval(vec[None, :], lc, None, self, want_loglike=True)[1][0])


def _gp_eval(params, lc, priors, gp, want_loglike=False):
    from . import namaste as _nm
    kind, kfull, mask = _kernel_spec(gp.kernel)
    nfree = bin(mask).count("1")
    ndim = nfree + 6
    p = as_f64(params)
    if p.ndim != 2 or p.shape[1] != ndim:
        raise ValueError("walker vectors must have %d entries (%d free kernel parameters + 6 transit parameters)"
                         % (ndim, nfree))
    oversamp = getattr(gp.mean, "oversamp", 10)
    staged = lc if isinstance(lc, core.LightCurve) else _nm.stage_lightcurve(lc, oversamp)
    if staged.ncand != 1:
        raise ValueError("the GP objective takes one light curve")
    tab = core.prior_table(priors, npar=ndim)
    out = np.empty(len(p))
    ll = np.empty(len(p)) if want_loglike else None
    kf = as_f64(kfull)          # keep the converted array alive across the call
    check(load().nmb_gp_lnprob_host(staged._h, ptr(p), len(p), kind, ptr(kf), mask, ptr(tab), ptr(out), ptr(ll)),
          "MonoLogProb")
    return out, ll


def MonoLogProb(params, lc, priors, gp):
    """GP log-probability (namaste.py:1334-1339).  ``params`` is one vector in ``gp.get_parameter_vector()``
    layout (returns a float and, like the reference, leaves the vector set on ``gp``) or an ensemble
    [W, ndim] (returns [W])."""
    p = np.asarray(params, dtype=np.float64)
    if p.ndim == 1:
        gp.set_parameter_vector(p)
        return float(_gp_eval(p[None, :], lc, priors, gp)[0][0])
    return _gp_eval(p, lc, priors, gp)[0]


def MonoNegLogProb(params, lc, priors, gp):
    return -1 * MonoLogProb(params, lc, priors, gp)
