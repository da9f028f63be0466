"""Multi-GPU modes of the likelihood path (one process per GPU, torch.distributed).

(a) Independent candidates (BASELINE config 4): every candidate has its own light curve, priors,
    ensemble and RNG stream, so candidates are dealt to ranks round-robin and there is NO
    data-path collective -- :func:`shard_candidates`.
(b) One large ensemble split by walker (BASELINE config 5): every rank holds the full light curve
    and a replica of the ensemble, updates its slice of the active half-ensemble, and the slices
    are all-gathered (NCCL over NVLink) after each half-step -- :class:`WalkerSplit`,
    :func:`allgather_half`, :class:`SplitEnsembleSampler`.  Because the random stream is keyed by
    (seed, step, half, walker), the chain is bit-identical for any number of ranks.

The reference has no collective at all (its only parallelism is emcee's process pool,
namaste/namaste.py:583/590); this module is the B200-native counterpart of that pool.
"""
import ctypes

import numpy as np

from ._lib import MODES, check, load, ptr, as_f64


def shard_candidates(ncand, rank, world):
    """Indices of the candidates rank `rank` of `world` owns (candidate k -> rank k mod world)."""
    return list(range(rank, ncand, world))


class WalkerSplit(object):
    """Slice bookkeeping for splitting each half-ensemble across `world` ranks."""

    def __init__(self, nwalkers, world, rank):
        if nwalkers % 2:
            raise ValueError("The number of walkers must be even.")
        self.W, self.world, self.rank = int(nwalkers), int(world), int(rank)
        self.halfk = self.W // 2
        if self.halfk % self.world:
            raise ValueError("half the ensemble (%d walkers) must divide evenly over %d ranks" % (self.halfk, world))
        self.nk = self.halfk // self.world
        self.k0 = self.rank * self.nk

    def rows(self, half, rank=None):
        """(first global walker row, count) updated by `rank` in half-step `half`."""
        r = self.rank if rank is None else rank
        return half * self.halfk + r * self.nk, self.nk


def allgather_half(pos, lnp, half, split, group=None):
    """All-gather the slices of half-ensemble `half` that each rank just updated, in place.
    `pos` [W, ndim] and `lnp` [W] are torch tensors (CUDA with NCCL, CPU with gloo)."""
    import torch.distributed as dist
    if split.world == 1:
        return
    r0, nk = split.rows(half)
    for full, trailing in ((pos, pos.shape[1:]), (lnp, ())):
        mine = full[r0:r0 + nk].clone()
        outs = [full[split.rows(half, r)[0]:split.rows(half, r)[0] + nk] for r in range(split.world)]
        dist.all_gather(outs, mine, group=group)


class _DevArray(object):
    """Minimal __cuda_array_interface__ view of library-owned device memory (zero copy)."""

    def __init__(self, addr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(addr), False),
                                         "version": 2, "strides": None}


class SplitEnsembleSampler(object):
    """Walker-split ensemble sampler: same chain as :class:`namaste_b200.EnsembleSampler` with the
    same seed, computed by `world` GPUs.  Requires an initialised torch.distributed process group."""

    def __init__(self, lightcurve, priors, nwalkers, seed, mode="windowed", a=2.0, group=None):
        import torch
        import torch.distributed as dist
        from . import core
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.split = WalkerSplit(nwalkers, self.world, self.rank)
        if lightcurve.ncand != 1:
            raise ValueError("walker split works on a single light curve")
        self._lc = lightcurve
        lib = load()
        tab = core.prior_table(priors)
        h = ctypes.c_void_p()
        check(lib.nmb_sampler_create(lightcurve._h, int(nwalkers), ptr(tab), ctypes.c_uint64(int(seed)), float(a),
                                     MODES[mode], ctypes.byref(h)), "SplitEnsembleSampler")
        self._h = h
        self.W = int(nwalkers)
        self.pos = torch.as_tensor(_DevArray(lib.nmb_sampler_devptr(h, 0), (self.W, 6)), device="cuda")
        self.lnp = torch.as_tensor(_DevArray(lib.nmb_sampler_devptr(h, 1), (self.W,)), device="cuda")
        self.iterations = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                load().nmb_sampler_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run_mcmc(self, pos0, nsteps, storechain=True):
        """Advance the replicated ensemble `nsteps` steps.  Everything of a step -- the rank's slice of
        each half-ensemble (two launches), the all-gather of the updated slices, the chain append --
        is ordered on the sampler's CUDA stream; the host only waits once, at the end."""
        import torch
        lib = load()
        if pos0 is not None:
            check(lib.nmb_sampler_set_state(self._h, ptr(as_f64(pos0)), None), "run_mcmc")
        sp = self.split
        stream = torch.cuda.ExternalStream(int(lib.nmb_sampler_stream(self._h)))
        with torch.cuda.stream(stream):
            for _ in range(int(nsteps)):
                for half in (0, 1):
                    check(lib.nmb_sampler_halfstep(self._h, half, sp.k0, sp.nk, 0), "run_mcmc")
                    allgather_half(self.pos, self.lnp, half, sp, self.group)   # NCCL orders itself after / before this stream
                check(lib.nmb_sampler_advance(self._h, 0, 1 if storechain else 0), "run_mcmc")
        check(lib.nmb_sampler_sync(self._h), "run_mcmc")      # one host wait + NaN check (emcee: ValueError)
        self.iterations += int(nsteps)
        return self.pos.cpu().numpy(), self.lnp.cpu().numpy()

    def _get(self, what):
        lib = load()
        T = int(lib.nmb_sampler_len(self._h))
        shape = {2: (self.W, T, 6), 3: (self.W, T), 4: (self.W,)}[what]
        out = np.zeros(shape, dtype=np.int64 if what == 4 else np.float64)
        if out.size:
            check(lib.nmb_sampler_get(self._h, what, ptr(out)), "SplitEnsembleSampler")
        return out

    @property
    def chain(self):
        return self._get(2)

    @property
    def lnprobability(self):
        return self._get(3)

    @property
    def naccepted(self):
        """Acceptance counts of all walkers (each rank counts its own; summed over ranks)."""
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(self._get(4)).cuda()
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()
