"""Host-side mirror of the reference's interface for the likelihood hot path.

Same names, argument meaning and error behaviour as ``namaste/namaste.py`` so that a
Namaste user can switch imports; all arithmetic runs in the CUDA library.

  MonotransitModel   <- namaste/namaste.py:1152-1164 (celerite.modeling.Model duck type)
  MonoLnPrior        <- namaste/namaste.py:1314-1332
  MonoOnlyLogProb    <- namaste/namaste.py:1345-1349   (also accepts an ensemble [W, 6])
  MonoOnlyNegLogProb <- namaste/namaste.py:1351-1352
  CalcTdur / CalcVel <- namaste/namaste.py:1469-1477
  calcminp / calcmaxvel / monoPriors / initLD / BuildMeanPriors
                     <- namaste/namaste.py:947-955, :938-945, :1060-1067, :475-480 (prior construction)
  transit_window_mask / initial_walkers / trim_samples
                     <- namaste/namaste.py:550-556, :537-547, :597-608 (RunMCMC set-up and trimming)
  Optimize_mono      <- namaste/namaste.py:1010-1057 (multi-start L-BFGS-B; the stencils of all starts
                        are evaluated as one ensemble per round)
"""
from collections import OrderedDict

import numpy as np

from . import core, _lib

__all__ = ["MonotransitModel", "MonoLnPrior", "MonoOnlyLogProb", "MonoOnlyNegLogProb", "CalcTdur", "CalcVel",
           "stage_lightcurve", "clear_cache", "calcminp", "calcmaxvel", "monoPriors", "initLD", "BuildMeanPriors",
           "transit_window_mask", "nonan", "AnomCutDiff", "initial_walkers", "trim_samples", "RunMCMC", "Optimize_mono", "VelToOrbit", "PlanetRtoM", "MonoFinalPars"]

_LC_CACHE = OrderedDict()
_LC_CACHE_MAX = 8


def _fingerprint(a):
    """Identity + content key of a light-curve array: address and layout, plus a position-sensitive checksum
    over EVERY 64-bit word of its float64 image (nmb_host_checksum: one pass at memory speed, ~2 us for a
    4000-point light curve, ~1 ms for 500 000 points).  A sampled probe is not enough: a light curve
    re-detrended in place, or a new array that lands on a freed address with the same times and errors,
    differs from the staged one only in some flux values.  An array that cannot change under the caller --
    ``writeable`` False on it and on its base -- is keyed by identity alone (``lc.setflags(write=False)`` is
    the cheap way to call MonoOnlyLogProb with a raw array in a loop; passing the staged LightCurve is free)."""
    ident = (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str)
    base = a.base
    if not a.flags.writeable and (base is None or not getattr(getattr(base, "flags", None), "writeable", True)):
        return ident
    img = a if (a.dtype == np.float64 and a.flags.c_contiguous) else np.ascontiguousarray(a, dtype=np.float64)
    out = np.zeros(2, dtype=np.uint64)
    lib = _lib.load()
    _lib.check(lib.nmb_host_checksum(_lib.ptr(img), img.size, _lib.ptr(out)), "stage_lightcurve")
    return ident + (int(out[0]), int(out[1]))


def stage_lightcurve(lc, oversamp=10):
    """Return the device-staged LightCurve for ``lc`` (float64 [N, 3]), staging it once.

    The reference passes the same ``lc`` array to every log-prob call (emcee ``args=``,
    namaste.py:590); here it is copied to HBM on first use and found again by identity."""
    if isinstance(lc, core.LightCurve):  # already staged by the caller: no lookup, no checksum
        return lc
    lc = np.asarray(lc)
    key = (_fingerprint(lc), int(oversamp))
    hit = _LC_CACHE.get(key)
    if hit is not None:
        _LC_CACHE.move_to_end(key)
        return hit
    staged = core.LightCurve(np.ascontiguousarray(lc, dtype=np.float64), oversamp=oversamp)
    _LC_CACHE[key] = staged
    # Eviction only drops the cache's reference: a live EnsembleSampler keeps the raw handle of its
    # LightCurve, so the device memory must stay until the last Python reference goes (LightCurve.__del__).
    while len(_LC_CACHE) > _LC_CACHE_MAX:
        _LC_CACHE.popitem(last=False)
    return staged


def clear_cache():
    """Forget every staged light curve (their device memory is freed once nothing else refers to them)."""
    _LC_CACHE.clear()


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


# ---- prior construction feeding MonoLnPrior (once per star, host side) ----------------------------
def calcminp(lc, tcen, tdur):
    """Shortest period compatible with seeing a single transit (Monotransit.calcminp,
    namaste.py:947-955): the distance from ``tcen`` to the first ``tdur``-wide jump in |t - tcen|
    taken in array order (or to the farthest point if there is none) plus a third of the duration."""
    dist = abs(np.asarray(lc)[:, 0] - tcen)
    dur_jumps = np.where(np.diff(dist) > tdur)[0]
    if len(dur_jumps) == 0:
        return np.max(dist) + tdur * 0.33
    return dist[dur_jumps[0]] + tdur * 0.33


def calcmaxvel(lc, sdenss, tcen, tdur):
    """Maximum transverse velocity [Rs/day] from the stellar-density PDF (Monotransit.calcmaxvel,
    namaste.py:938-945).  Returns (maxvel, maxveluerr, maxvelderr, minp)."""
    minp = calcminp(lc, tcen, tdur)
    maxvels = ((18226 * abs(np.asarray(sdenss, dtype=np.float64))) / minp) ** (1 / 3.0)
    prcnts = np.percentile(maxvels, [15.865525393145707, 50.0, 84.13447460685429])
    return prcnts[1], prcnts[2] - prcnts[1], prcnts[1] - prcnts[0], minp


def monoPriors(tcen, tdur, maxvel, maxveluerr, name="mean"):
    """Box + density prior rows of one transit (Monotransit.monoPriors, namaste.py:1060-1067)."""
    return OrderedDict([
        (name + ":tcen", [tcen - tdur * 0.3, tcen + tdur * 0.3]),
        (name + ":b", [0.0, 1.25]),
        (name + ":vel", [0, maxvel + 5 * maxveluerr, "normlim", maxvel, maxveluerr]),
        (name + ":RpRs", [0.02, 0.25]),
    ])


def initLD(LD1s, LD2s):
    """Limb-darkening prior rows from the LD coefficient PDFs (Star.initLD, namaste.py:475-480)."""
    return OrderedDict([
        ("LD1", [0, 1.0, "gaussian", np.median(LD1s), np.std(LD1s)]),
        ("LD2", [0, 1.0, "gaussian", np.median(LD2s), np.std(LD2s)]),
    ])


def BuildMeanPriors(lc, tcen, tdur, sdenss, LD1s, LD2s, name="mean"):
    """The ordered six-row priors mapping MonoOnlyLogProb expects (namaste.py:1316: entry n applies
    to params[n]), assembled the way Star.BuildMeanPriors does: the transit rows, then the LD rows
    under the same ``name:`` prefix."""
    maxvel, maxveluerr, _, _ = calcmaxvel(lc, sdenss, tcen, tdur)
    priors = monoPriors(tcen, tdur, maxvel, maxveluerr, name)
    for key, row in initLD(LD1s, LD2s).items():
        priors[name + ":" + key] = row
    return priors


def nonan(lc):
    """Rows of a light curve that hold no NaN in any column (namaste.py:1456-1457)."""
    return ~np.isnan(np.sum(np.asarray(lc, dtype=np.float64), axis=1))


def AnomCutDiff(flux, thresh=4.2):
    """Mask of the points to KEEP after the single-point outlier cut (namaste.py:1459-1467; Namaste calls
    it with 3.75 before detrending, :64 / :1109).  An interior point is dropped only when it is a spike:
    it differs from BOTH neighbours by at least ``thresh`` times the median absolute point-to-point
    difference and the two steps have opposite signs.  Two outliers in a row, ramps and the end points
    stay.  ``flux`` must be NaN-free (``nonan``).  Host NumPy like the reference: once per star."""
    f = np.asarray(flux, dtype=np.float64)
    step = np.diff(f)
    # the reference scales by the median of |diff(flux[1:])|, i.e. without the first step
    scale = np.median(np.abs(step[1:]))
    ahead, behind = step[1:] / scale, step[:-1] / scale      # f[i+1] - f[i] and f[i] - f[i-1] for i = 1 .. n-2
    keep = (ahead * behind > 0) | (np.abs(ahead) < thresh) | (np.abs(behind) < thresh)
    return np.concatenate(([True], keep, [True]))


def transit_window_mask(t, tcen, vel, b, RpRs, ndurswindow=6, two_sided=False):
    """Time window RunMCMC fits (namaste.py:550-556).  The reference takes ``abs()`` of a boolean,
    which keeps everything *before* ``tcen + ndurswindow * Tdur`` (quirk Q6 of SURVEY.md);
    ``two_sided=True`` gives the intended |t - tcen| < ndurswindow * Tdur."""
    t = np.asarray(t, dtype=np.float64)
    width = ndurswindow * CalcTdur(vel, b, RpRs)
    if two_sided:
        return np.abs(t - tcen) < width
    return abs(t - tcen < width).astype(bool)


def initial_walkers(pdfs, nwalkers, rng=None):
    """Initial ensemble drawn from the parameter PDFs (namaste.py:537-547): ``nwalkers`` distinct rows
    of the [npdf, 6] sample table (tcen, b, vel, RpRs, LD1, LD2); NaNs are replaced by
    ``nanmedian(isnan(column))`` = 0.0 as the reference does (quirk Q5)."""
    pdfs = np.asarray(pdfs, dtype=np.float64)
    rng = np.random if rng is None else rng
    chx = rng.choice(len(pdfs), nwalkers, replace=False)
    pos = pdfs[chx].copy()
    for col in range(pos.shape[1]):
        bad = np.isnan(pos[:, col])
        pos[bad, col] = np.nanmedian(bad)
    return pos


# ---- post-run trimming (namaste.py:597-608) --------------------------------------------------------
def trim_samples(chain, lnprobability, nsteps):
    """Burn-in cut and failed-walker filter of RunMCMC.  ``chain`` is [W, T, ndim], ``lnprobability``
    [W, T].  Returns (samples [n, ndim + 1] with the log-probability as last column, good_wlkrs, ncut)."""
    chain = np.asarray(chain)
    lnprobability = np.asarray(lnprobability)
    ncut = np.min([int(nsteps * 0.25), 3000])
    prcnt = np.percentile(lnprobability[:, ncut:], [50, 95], axis=1)
    # "failed" walkers: 95th percentile below the median of the walkers' medians
    good_wlkrs = (prcnt[1] > np.median(prcnt[0]))
    ndim = chain.shape[-1]
    samples = chain[good_wlkrs, ncut:, :].reshape((-1, ndim))
    samples = np.column_stack((samples, lnprobability[good_wlkrs, ncut:].reshape(-1)))
    return samples, good_wlkrs, ncut


# ---- the sampling run of one star (Star.RunMCMC, namaste.py:537-608, no-GP branch) -------------------
def RunMCMC(lc, priors, pdfs, meanmodel, nwalkers=24, nsteps=12500, ndurswindow=6, rng=None, seed=None,
            mode="windowed", two_sided_window=False):
    """What ``Star.RunMCMC`` does between building the priors and saving the samples, for the transit-only
    objective (``Settings.GP = False``), with the reference's defaults (24 walkers, 12 500 steps, window of
    6 durations; namaste.py:57-61):

    1. ``nwalkers`` distinct rows of the PDF table ``pdfs`` [npdf, 6] become the initial ensemble, NaNs
       replaced as the reference replaces them (:537-547, ``initial_walkers``);
    2. the light curve is cut to the window around ``meanmodel``'s transit (:550-556, including the
       reference's one-sided quirk unless ``two_sided_window``);
    3. ``EnsembleSampler(nwalkers, 6, MonoOnlyLogProb, args=(lc[mask], priors, meanmodel))`` runs one step
       and then ``nsteps`` steps, both from the initial ensemble, the second call appending to the chain
       (:590-594) -- proposal, likelihood and accept/reject stay on the GPU for the whole run;
    4. burn-in cut ``min(nsteps/4, 3000)`` and failed-walker filter (:597-608, ``trim_samples``).

    Returns a dict with ``samples`` [n, 7] (six parameters + ``logprob``), ``sampleheaders``, the
    ``sampler`` (chain and lnprobability stay in HBM until read), ``mask``, ``init_mcmc_params``,
    ``good_wlkrs`` and ``ncut``.  Plotting, logging and file output of the reference are not part of it."""
    from .sampler import EnsembleSampler
    lc = np.asarray(lc, dtype=np.float64)
    init = initial_walkers(pdfs, nwalkers, rng=rng)
    par = meanmodel.get_parameter_dict()
    mask = transit_window_mask(lc[:, 0], par["tcen"], par["vel"], par["b"], par["RpRs"], ndurswindow,
                               two_sided=two_sided_window)
    fit_lc = np.ascontiguousarray(lc[mask, :])
    sampler = EnsembleSampler(nwalkers, init.shape[1], MonoOnlyLogProb, args=(fit_lc, priors, meanmodel),
                              mode=mode, seed=seed)
    sampler.run_mcmc(init, 1)
    sampler.run_mcmc(init, nsteps)
    samples, good, ncut = trim_samples(sampler.chain, sampler.lnprobability, nsteps)
    headers = ["mean:" + nm for nm in MonotransitModel.parameter_names] + ["logprob"]
    return {"samples": samples, "sampleheaders": headers, "sampler": sampler, "mask": mask,
            "init_mcmc_params": init, "good_wlkrs": good, "ncut": ncut}


# ---- multi-start optimisation of the transit parameters (namaste.py:1010-1057) -----------------
class _BatchedObjective(object):
    """Lets several scipy optimisers, each running in its own thread, share one GPU call per round:
    every optimiser posts the rows it needs (its point and the forward-difference stencil), the last
    one to arrive evaluates all posted rows as a single ensemble and wakes the others."""

    def __init__(self, evaluate, nclients):
        import threading
        self.evaluate = evaluate
        self.cv = threading.Condition()
        self.active = nclients
        self.pending = {}
        self.results = {}
        self.generation = 0
        self.ncalls = 0
        self.error = None

    def _flush(self):
        keys = list(self.pending.keys())
        try:
            rows = np.concatenate([self.pending[k] for k in keys], axis=0)
            vals = self.evaluate(rows)
            self.ncalls += 1
            at = 0
            for k in keys:
                n = len(self.pending[k])
                self.results[k] = vals[at:at + n]
                at += n
        except BaseException as exc:  # a failed GPU call must wake every waiting optimiser, not strand it
            self.error = exc
        finally:
            self.pending.clear()
            self.generation += 1
            self.cv.notify_all()

    def call(self, key, rows):
        with self.cv:
            if self.error is not None:
                raise self.error
            self.pending[key] = rows
            if len(self.pending) == self.active:
                self._flush()
            else:
                gen = self.generation
                while self.generation == gen:
                    self.cv.wait()
            if self.error is not None:
                raise self.error
            return self.results.pop(key)

    def leave(self):
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.pending) == self.active:
                self._flush()


def Optimize_mono(flatlc, priors, starts, tcen=None, tdur=None, monomodel=None, eps=1e-8, mode="windowed",
                  options=None):
    """Multi-start L-BFGS-B fit of (tcen, b, vel, RpRs, LD1, LD2) as Monotransit.Optimize_mono does
    (namaste.py:1010-1057): the light curve is cut to |t - tcen| < 5 tdur (when both are given), every
    row of ``starts`` seeds one ``scipy.optimize.minimize(MonoOnlyNegLogProb, x0, method="L-BFGS-B")``
    and the successful, non-NaN results are ranked by negative log-probability.

    The optimiser is scipy's own; what changes is where the objective runs: each optimiser asks for
    its point and the six forward-difference neighbours (scipy's default gradient for L-BFGS-B, step
    ``eps`` = 1e-8), and the requests of all starts of a round are evaluated as ONE ensemble on the GPU.
    Returns (bestres [7] = best parameters + its negative log-probability, suc_res [n_success, 7])."""
    import threading
    import scipy.optimize as opt
    flatlc = np.asarray(flatlc, dtype=np.float64)
    if tcen is not None and tdur is not None:
        flatlc = flatlc[abs(flatlc[:, 0] - tcen) < 5 * tdur]
    flatlc = np.ascontiguousarray(flatlc)
    starts = np.atleast_2d(np.asarray(starts, dtype=np.float64))
    oversamp = getattr(monomodel, "oversamp", 10) if monomodel is not None else 10
    staged = stage_lightcurve(flatlc, oversamp)
    # forward differences with a 1e-8 step see the rounding of the objective: evaluate it in the reference's
    # own rounding sequence (dtype "strict"), so the optimisers walk the surface the reference's would
    batch = _BatchedObjective(lambda rows: -1 * staged.lnprob(rows, priors, mode=mode, dtype="strict"), len(starts))
    results = [None] * len(starts)

    def run(i):
        def fun_and_grad(x):
            rows = np.repeat(x[None, :], 7, axis=0)
            rows[1:] += eps * np.eye(6)
            f = batch.call(i, rows)
            return f[0], (f[1:] - f[0]) / eps
        try:
            results[i] = opt.minimize(fun_and_grad, starts[i], jac=True, method="L-BFGS-B", options=options)
        finally:
            batch.leave()

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(starts))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if batch.error is not None:
        raise batch.error
    suc = [np.hstack((r.x, r.fun)) for r in results if r is not None and r.success]
    if len(suc) == 0:
        raise ValueError("No successful Monotransit minimizations")
    suc_res = np.vstack(suc)
    suc_res = suc_res[~np.isnan(suc_res[:, -1]), :]
    bestres = suc_res[np.argmin(suc_res[:, -1])]
    if monomodel is not None:
        monomodel.set_parameter_vector(bestres[:-1])
    Optimize_mono.last_gpu_calls = batch.ncalls
    return bestres, suc_res


# ---- derived planet parameters (MonoFinalPars, namaste.py:629-652; planetlib.py:1346-1373) -------------
def VelToOrbit(vel, Rs, Ms, dens=None, ecc=0, omega=0, timebin=86400.0):
    """Semi-major axis [AU] and period [days] of a circular orbit from the transverse velocity in
    stellar radii per day (planetlib.VelToOrbit, planetlib.py:1346-1355)."""
    if dens is None:
        dens = Ms / Rs ** 3
    Per = 18226 * dens / (vel ** 3)
    SMA = 13.7375 * (Ms * (Rs / vel) ** 2)
    return SMA, Per


def PlanetRtoM(Rad):
    """Planet mass [Mearth] from radius [Rearth] (planetlib.PlanetRtoM, planetlib.py:1367+): rocky
    density law below 1.5, Weiss & Marcy power law to 4, a tabulated giant branch to 16, stellar above."""
    Rearth, Mearth = 6.137e6, 5.96e24
    Rad = np.atleast_1d(np.asarray(Rad, dtype=np.float64))
    M = np.zeros_like(Rad)
    small = Rad <= 1.5
    M[small] = ((2430 + 3390 * Rad[small]) * (1.3333 * np.pi * (Rad[small] * Rearth) ** 3)) / Mearth
    mid = (Rad > 1.5) * (Rad <= 4.0)
    M[mid] = 2.69 * (Rad[mid] ** 0.93)
    giants = (Rad > 4.) * (Rad <= 16.0)
    if giants.sum() > 0:
        xr = np.array([3.99, 9.0, 10.5, 12.0, 13.0, 14.0, 14.5, 15.5, 16.0])
        yr = np.array([10.0, 100.0, 300.0, 1000.0, 1100.0, 1200.0, 1350.0, 5000.0, 10000.0])
        M[giants] = np.interp(Rad[giants], xr, yr)
    big = Rad > 16.0
    M[big] = 333000 * (Rad[big] / 109.1) ** 1.5
    return M


def MonoFinalPars(samples, srads, smasss, sdenss, rows=None, sigs=(15.865525393145707, 50.0, 84.13447460685429),
                  rng=None):
    """Derived planet quantities of Star.MonoFinalPars (namaste.py:629-652) for sample rows ``rows`` of
    ``samples`` [n, >=4] (tcen, b, vel, RpRs, ...) paired with the stellar PDFs: planet radius [Rearth],
    semi-major axis [AU], period [d], mass [Mearth], RV semi-amplitude [m/s]; returns the PDFs and
    (median, +err, -err) of each.  By default the rows are drawn like the reference draws them,
    ``choice(len(samples), npdf, replace=False)`` (namaste.py:636) from ``rng`` (``np.random`` if None):
    samples are walker-major, so the first npdf rows would be one autocorrelated stretch of one chain."""
    samples = np.asarray(samples, dtype=np.float64)
    if rows is None:
        rows = (np.random if rng is None else rng).choice(len(samples), len(srads), replace=False)
    rows = np.asarray(rows)
    out = {}
    out["Rps"] = (samples[rows, 3] * 695500000 * srads) / 6.371e6
    out["Prob_pl"] = len(out["Rps"][out["Rps"] < (1.5 * 11.2)]) / len(out["Rps"])
    out["smas"], out["Ps"] = VelToOrbit(samples[rows, 2], srads, smasss, sdenss)
    out["Mps"] = PlanetRtoM(out["Rps"])
    out["Krvs"] = ((2. * np.pi * 6.67e-11) / (out["Ps"] * 86400)) ** (1. / 3.) * (out["Mps"] * 5.96e24 / ((1.96e30 * smasss) ** (2. / 3.)))
    for val in ["Rps", "smas", "Ps", "Mps", "Krvs"]:
        p = np.percentile(np.array(out[val]), sigs)
        out[val[:-1]] = p[1]
        out[val[:-1] + "uerr"] = p[2] - p[1]
        out[val[:-1] + "derr"] = p[1] - p[0]
    return out
