"""emcee-2.x-shaped front end of the device-resident ensemble sampler.

Mirrors the surface Namaste uses (namaste/namaste.py:583-608; Example.ipynb cells 30-35):
``EnsembleSampler(nwalkers, dim, lnpostfn, args=(), threads=1)``, ``run_mcmc(pos0, N, rstate0=None)``
(successive calls append), ``.chain [W, T, dim]``, ``.lnprobability [W, T]``, ``.flatchain``,
``.flatlnprobability``, ``.acceptance_fraction``, ``.naccepted``, ``.iterations``, ``.reset()``.

``lnpostfn`` must be this package's ``MonoOnlyLogProb`` (args = (lc, priors[, monomodel]),
namaste.py:590) or ``MonoLogProb`` (GP noise model, args = (lc, priors, gp), namaste.py:583): proposal,
likelihood and accept/reject then run on the GPU with no per-step host round trip.  Any other callable
raises ``TypeError`` (there is no CPU fallback).  Random numbers are Philox4x32-10 (counter based), not NumPy's Mersenne Twister:
chains agree with emcee's in distribution, not draw by draw; ``rstate0`` only seeds the stream.
"""
import ctypes

import numpy as np

from . import core, gp as _gp, namaste as _nm
from ._lib import MODES, check, load, ptr, as_f64

__all__ = ["EnsembleSampler"]


def _seed_from(rstate0, seed):
    if seed is not None:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    if rstate0 is None:
        return int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
    if isinstance(rstate0, (int, np.integer)):
        return int(rstate0) & 0xFFFFFFFFFFFFFFFF
    try:  # np.random.get_state() tuple: fold the key words into 64 bits
        key = np.asarray(rstate0[1], dtype=np.uint64)
        h = np.uint64(1469598103934665603)
        for w in key[:16]:
            h = np.uint64((int(h) ^ int(w)) * 1099511628211 & 0xFFFFFFFFFFFFFFFF)
        return int(h) ^ int(rstate0[2])
    except Exception:
        raise ValueError("rstate0 must be None, an integer seed or a numpy.random.get_state() tuple")


class EnsembleSampler(object):
    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, args=(), kwargs=None, threads=1, pool=None,
                 live_dangerously=False, mode="windowed", seed=None):
        if lnpostfn not in (_nm.MonoOnlyLogProb, _gp.MonoLogProb):
            raise TypeError("namaste_b200.EnsembleSampler runs the log-probability on the GPU and only accepts "
                            "namaste_b200.MonoOnlyLogProb or namaste_b200.MonoLogProb as lnpostfn (no CPU fallback)")
        self._gp = None
        if lnpostfn is _gp.MonoLogProb:
            if len(args) < 3 or not isinstance(args[2], _gp.GP):
                raise ValueError("args must be (lc, priors, gp) as in namaste.py:583")
            self._gp = args[2]
            self._gpspec = _gp._kernel_spec(self._gp.kernel)
            want = bin(self._gpspec[2]).count("1") + 6
        else:
            want = 6
        if dim != want:
            raise ValueError("this objective has %d parameters; got dim=%r" % (want, dim))
        if nwalkers % 2 != 0:
            raise AssertionError("The number of walkers must be even.")
        if nwalkers < 2 * dim:
            raise AssertionError("The number of walkers needs to be more than twice the dimension of your "
                                 "parameter space.")
        if len(args) < 2:
            raise ValueError("args must be (lc, priors[, monomodel]) as in namaste.py:590")
        self.k, self.dim, self.a = int(nwalkers), int(dim), float(a)
        self.args, self.threads = args, threads  # threads is accepted and ignored (one GPU does the work)
        lc, priors = args[0], args[1]
        model = self._gp.mean if self._gp is not None else (args[2] if len(args) > 2 else None)
        self._oversamp = getattr(model, "oversamp", 10) if model is not None else 10
        self._lc = lc if isinstance(lc, core.LightCurve) else _nm.stage_lightcurve(lc, self._oversamp)
        self._ptab = core.prior_table(priors, npar=self.dim)
        self._mode, self._seed0 = mode, seed
        self._h = None
        self._have_state = False
        self.iterations = 0
        self._last = None
        self._cand0 = 0   # index of the batch's first candidate in the random stream (CandidateBatchSampler)

    # ---- handle management --------------------------------------------------------------------
    def _ensure(self, rstate0):
        if self._h is not None:
            return
        lib = load()
        tab = None
        if self._ptab is not None:
            tab = np.ascontiguousarray(np.broadcast_to(self._ptab, (self._lc.ncand, self.dim, 6)))
        h = ctypes.c_void_p()
        seed = ctypes.c_uint64(_seed_from(rstate0, self._seed0))
        if self._gp is not None:
            kind, kfull, mask = self._gpspec
            kf = as_f64(kfull)  # keep the converted array alive across the call
            check(lib.nmb_sampler_create_gp(self._lc._h, self.k, kind, ptr(kf), mask, ptr(tab), seed, self.a,
                                            ctypes.byref(h)), "EnsembleSampler")
        else:
            check(lib.nmb_sampler_create(self._lc._h, self.k, ptr(tab), seed, self.a, MODES[self._mode],
                                         ctypes.byref(h)), "EnsembleSampler")
        self._h = h
        if self._cand0:
            check(lib.nmb_sampler_set_candidate_offset(h, int(self._cand0)), "EnsembleSampler")

    def __del__(self):
        try:
            if self._h is not None:
                load().nmb_sampler_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- emcee surface ---------------------------------------------------------------------------
    def run_mcmc(self, pos0, N, rstate0=None, lnprob0=None, thin=1, storechain=True, **kwargs):
        """Advance the ensemble N steps from pos0 (or from where the last call stopped if None).
        Returns (pos, lnprob, rstate) like emcee 2.x; rstate is the (seed, step) pair of the stream."""
        lib = load()
        if not getattr(self._lc, "_h", None):
            raise ValueError("the LightCurve of this sampler has been closed")
        self._ensure(rstate0)
        if pos0 is None:
            if not self._have_state:
                raise ValueError("Cannot have pos0=None if run_mcmc has never been called.")
        else:
            p = as_f64(pos0)
            want = (self.k, self.dim) if self._lc.ncand == 1 else (self._lc.ncand, self.k, self.dim)
            if p.shape != want:
                raise ValueError("pos0 must have shape %r" % (want,))
            l0 = None if lnprob0 is None else as_f64(lnprob0)
            check(lib.nmb_sampler_set_state(self._h, ptr(p), ptr(l0)), "run_mcmc")
            self._have_state = True
        check(lib.nmb_sampler_run(self._h, int(N), int(thin), 1 if storechain else 0), "run_mcmc")
        self.iterations += int(N)
        pos, lnp = self._get(0), self._get(1)
        self._last = (pos, lnp)
        return pos, lnp, ("philox4x32-10", int(lib.nmb_sampler_steps(self._h)))

    def sample(self, p0, lnprob0=None, rstate0=None, iterations=1, thin=1, storechain=True, **kwargs):
        """Generator form: yields (pos, lnprob, rstate) after every step."""
        first = True
        for _ in range(int(iterations)):
            yield self.run_mcmc(p0 if first else None, 1, rstate0=rstate0, lnprob0=lnprob0 if first else None,
                                thin=1, storechain=storechain)
            first = False

    def _get(self, what):
        lib = load()
        C, W = self._lc.ncand, self.k
        T = int(lib.nmb_sampler_len(self._h)) if self._h is not None else 0
        shape = {0: (C, W, self.dim), 1: (C, W), 2: (C, W, T, self.dim), 3: (C, W, T), 4: (C, W)}[what]
        out = np.zeros(shape, dtype=np.int64 if what == 4 else np.float64)
        if self._h is not None and out.size:
            check(lib.nmb_sampler_get(self._h, what, ptr(out)), "EnsembleSampler")
        return out[0] if C == 1 else out

    @property
    def chain(self):
        """Markov chain, shape [nwalkers, iterations, dim] (emcee layout)."""
        return self._get(2)

    @property
    def lnprobability(self):
        return self._get(3)

    @property
    def flatchain(self):
        c = self.chain
        return c.reshape(c.shape[:-3] + (c.shape[-3] * c.shape[-2], self.dim))

    @property
    def flatlnprobability(self):
        l = self.lnprobability
        return l.reshape(l.shape[:-2] + (l.shape[-2] * l.shape[-1],))

    @property
    def naccepted(self):
        return self._get(4)

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(max(self.iterations, 1))

    def summary(self, nsteps=None, percentiles=(15.865525393145707, 50.0, 84.13447460685429)):
        """RunMCMC's trimming and percentile summaries (namaste.py:597-608, :643-652) computed on the
        device-resident chain: returns a dict with ``ncut``, ``good_wlkrs`` [W] (failed-walker filter),
        ``nsamples`` and ``percentiles`` [len(percentiles), ndim + 1] of the kept, post-burn-in samples
        (last column: log-probability), i.e. ``np.percentile(samples[:, d], q)`` of
        ``namaste_b200.trim_samples(chain, lnprobability, nsteps)`` without copying the chain."""
        lib = load()
        if self._h is None:
            raise ValueError("run_mcmc has never been called")
        nsteps = self.iterations if nsteps is None else int(nsteps)
        q = as_f64(percentiles)
        C, W = self._lc.ncand, self.k
        pct = np.empty((C, len(q), self.dim + 1))
        good = np.zeros((C, W), dtype=np.uint8)
        ncut, nsamp = ctypes.c_int64(), ctypes.c_int64()
        check(lib.nmb_sampler_summary(self._h, nsteps, ptr(q), len(q), ptr(pct), ptr(good), ctypes.byref(ncut),
                                      ctypes.byref(nsamp)), "summary")
        good = good.astype(bool)
        return {"ncut": int(ncut.value), "nsamples": int(nsamp.value), "good_wlkrs": good[0] if C == 1 else good,
                "percentiles": pct[0] if C == 1 else pct}

    def reset(self):
        """Clear chain, lnprobability and acceptance counters (emcee: clear_chain/reset)."""
        if self._h is not None:
            check(load().nmb_sampler_reset(self._h), "reset")
        self.iterations = 0

    clear_chain = reset
