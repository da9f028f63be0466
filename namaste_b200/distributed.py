"""Multi-GPU modes of the likelihood path (one process per GPU, torch.distributed).

(a) Independent candidates (BASELINE config 4): every candidate has its own light curve, priors,
    ensemble and RNG stream, so candidates are dealt to ranks round-robin and there is NO
    data-path collective -- :func:`shard_candidates`.
(b) One large ensemble split by walker (BASELINE config 5): every rank holds the full light curve
    and a replica of the ensemble, updates its slice of the active half-ensemble, and the slices
    reach the other replicas after each half-step -- by device-side stores into the peers' memory over
    NVLink (default) or by one packed NCCL all-gather -- :class:`WalkerSplit`, :func:`allgather_half`,
    :class:`SplitEnsembleSampler`.  Because the random stream is keyed by
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
    """All-gather the slices of half-ensemble `half` that each rank just updated, in place, as ONE packed
    collective: rows [nk, ndim + 1] (position, lnprob) per rank.  `pos` [W, ndim] and `lnp` [W] are torch
    tensors (CPU with gloo in the tests; the CUDA path of SplitEnsembleSampler packs with its own kernels)."""
    import torch
    import torch.distributed as dist
    if split.world == 1:
        return
    r0, nk = split.rows(half)
    mine = torch.cat([pos[r0:r0 + nk], lnp[r0:r0 + nk, None]], dim=1).contiguous()
    full = torch.empty((split.world * nk, mine.shape[1]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(full, mine, group=group)
    h0 = half * split.halfk
    pos[h0:h0 + split.halfk] = full[:, :-1]
    lnp[h0:h0 + split.halfk] = full[:, -1]


class _DevArray(object):
    """Minimal __cuda_array_interface__ view of library-owned device memory (zero copy)."""

    def __init__(self, addr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(addr), False),
                                         "version": 2, "strides": None}


class SplitEnsembleSampler(object):
    """Walker-split ensemble sampler: same chain as :class:`namaste_b200.EnsembleSampler` with the
    same seed, computed by `world` GPUs of one node (one process per GPU, initialised torch.distributed
    process group).

    ``transport`` selects how the updated half-ensemble slices reach the other ranks:

    * ``"p2p"``  -- device side: every rank maps its peers' ensemble replicas through CUDA IPC, a kernel stores
      the slice into every replica over NVLink and exchanges arrival flags (csrc/split_kernels.cuh).  The whole
      run is one C call: no NCCL launch, no Python and no host synchronisation per step.
    * ``"nccl"`` -- one packed ``all_gather_into_tensor`` of [nk, ndim + 1] rows per half-step.
    * ``"auto"`` (default) -- ``"p2p"`` when every rank can map its peers, else ``"nccl"`` (with a warning)."""

    def __init__(self, lightcurve, priors, nwalkers, seed, mode="windowed", a=2.0, group=None, transport="auto"):
        import torch
        import torch.distributed as dist
        from . import core
        if transport not in ("auto", "p2p", "nccl"):
            raise ValueError("transport must be 'auto', 'p2p' or 'nccl'")
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
        self.transport = "local" if self.world == 1 else self._connect(transport)
        if self.transport == "nccl":
            self._send = torch.empty((self.split.nk, 7), dtype=torch.float64, device="cuda")
            self._recv = torch.empty((self.split.halfk, 7), dtype=torch.float64, device="cuda")

    def _connect(self, transport):
        """Exchange the CUDA IPC handles of the ensemble replicas and map the peers; every rank takes the same
        transport (the outcome is agreed with an all-reduce)."""
        import warnings
        import torch
        import torch.distributed as dist
        if transport == "nccl":
            return "nccl"
        lib = load()
        mine = np.zeros(64, dtype=np.uint8)
        check(lib.nmb_sampler_split_export(self._h, ptr(mine)), "SplitEnsembleSampler")
        send = torch.from_numpy(mine).cuda()
        recv = torch.empty(64 * self.world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(recv, send, group=self.group)
        handles = np.ascontiguousarray(recv.cpu().numpy())
        rc = lib.nmb_sampler_split_attach(self._h, self.rank, self.world, ptr(handles))
        why = load().nmb_last_error().decode("utf-8", "replace") if rc else ""
        ok = torch.tensor([0 if rc else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            return "p2p"
        if transport == "p2p":
            raise RuntimeError("SplitEnsembleSampler: peer mapping failed on at least one rank (%s)" % (why or "another rank"))
        warnings.warn("SplitEnsembleSampler: peer mapping unavailable (%s); using the NCCL transport" % (why or "another rank"))
        return "nccl"

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                load().nmb_sampler_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run_mcmc(self, pos0, nsteps, storechain=True):
        """Advance the replicated ensemble `nsteps` steps.  Everything of a step -- the rank's slice of
        each half-ensemble (two launches), the exchange of the updated slices, the chain append --
        is ordered on the sampler's CUDA stream; the host only waits once, at the end."""
        import torch
        import torch.distributed as dist
        lib = load()
        if pos0 is not None:
            p0 = as_f64(pos0)       # keep the (possibly converted) array alive across the call
            check(lib.nmb_sampler_set_state(self._h, ptr(p0), None), "run_mcmc")
        sp = self.split
        if self.transport in ("local", "p2p"):
            check(lib.nmb_sampler_run_split(self._h, int(nsteps), 1 if storechain else 0), "run_mcmc")
        else:
            stream = torch.cuda.ExternalStream(int(lib.nmb_sampler_stream(self._h)))
            with torch.cuda.stream(stream):
                for _ in range(int(nsteps)):
                    for half in (0, 1):
                        check(lib.nmb_sampler_halfstep(self._h, half, sp.k0, sp.nk, 0), "run_mcmc")
                        check(lib.nmb_sampler_pack_half(self._h, half, sp.k0, sp.nk, ptr(self._send)), "run_mcmc")
                        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)  # ordered on this stream
                        check(lib.nmb_sampler_unpack_half(self._h, half, ptr(self._recv)), "run_mcmc")
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


class CandidateBatchSampler(object):
    """MCMC of many independent single-transit candidates on ONE GPU (one rank's share of BASELINE config 4; the
    reference runs one ``Star.RunMCMC`` per candidate, namaste/namaste.py:526-612).

    Independent candidates need no lock step.  One staged batch stepped by one sampler still moves in lock step --
    every half-step is "window kernel, then slot queue" for all candidates at once, and while the short window
    kernel and the queue's last warps run, most of the GPU idles.  This class deals the candidates into ``groups``
    contiguous groups, each with its own staged batch, sampler and CUDA stream, and steps them from ``groups`` host
    threads (the library calls release the GIL): one group's latency-bound phases overlap the other's arithmetic.
    Measured on a B200, 256 candidates x 100 walkers x 4000 points: 7.4e5 candidate-steps/s as one group, 8.8e5 as two
    (profiles/r02q2_batch_groups.txt).

    Every group is told the index of its first candidate, so the chains are those of the one-batch sampler with the
    same seed, bit for bit (``nmb_sampler_set_candidate_offset``)."""

    def __init__(self, curves, priors, nwalkers, oversamp=10, seed=None, groups=2, mode="windowed", a=2.0):
        from . import core, namaste as _nm
        from .sampler import EnsembleSampler
        curves = list(curves)
        C = len(curves)
        if C < 1:
            raise ValueError("at least one light curve")
        if isinstance(priors, np.ndarray) and priors.ndim == 3:
            tabs = as_f64(priors)
        else:
            plist = list(priors) if isinstance(priors, (list, tuple)) else [priors] * C
            tabs = np.stack([core.prior_table(p) for p in plist])
        if tabs.shape != (C, 6, 6):
            raise ValueError("priors: one [6, 6] table (or dict) per candidate")
        groups = max(1, min(int(groups), C))
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        self.ncand, self.k, self.dim = C, int(nwalkers), 6
        dev = ctypes.c_int(0)
        check(load().nmb_get_device(ctypes.byref(dev)), "CandidateBatchSampler")
        self._device = int(dev.value)   # the group threads select it: the current CUDA device is per host thread
        bounds = [(C * g) // groups for g in range(groups + 1)]
        self._slices = [slice(bounds[g], bounds[g + 1]) for g in range(groups)]
        self._lcs, self._samplers = [], []
        for sl in self._slices:
            lc = core.LightCurve(curves[sl], oversamp=oversamp)
            smp = EnsembleSampler(self.k, 6, _nm.MonoOnlyLogProb, a=a, args=(lc, tabs[sl]), mode=mode, seed=seed)
            smp._cand0 = sl.start
            self._lcs.append(lc)
            self._samplers.append(smp)
        self.iterations = 0

    def _each(self, fn):
        """fn(group index, sampler) on every group, one host thread per group; re-raises the first error."""
        import threading
        out, errs = [None] * len(self._samplers), []

        def work(g):
            try:
                check(load().nmb_set_device(self._device), "CandidateBatchSampler")
                out[g] = fn(g, self._samplers[g])
            except BaseException as e:  # noqa: BLE001 -- re-raised below
                errs.append(e)
        if len(self._samplers) == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(g,)) for g in range(len(self._samplers))]
            for t in th:
                t.start()
            for t in th:
                t.join()
        if errs:
            raise errs[0]
        return out

    def run_mcmc(self, pos0, N, thin=1, storechain=True):
        """Advance every candidate's ensemble N steps; ``pos0`` is [ncand, nwalkers, 6] (None: continue).
        Returns (pos [ncand, W, 6], lnprob [ncand, W])."""
        p = None
        if pos0 is not None:
            p = as_f64(pos0)
            if p.shape != (self.ncand, self.k, 6):
                raise ValueError("pos0 must have shape %r" % ((self.ncand, self.k, 6),))

        def step(g, smp):
            sl = self._slices[g]
            part = None if p is None else np.ascontiguousarray(p[sl])
            if part is not None and part.shape[0] == 1:
                part = part[0]
            pos, lnp, _ = smp.run_mcmc(part, N, thin=thin, storechain=storechain)
            return (pos[None], lnp[None]) if pos.ndim == 2 else (pos, lnp)
        res = self._each(step)
        self.iterations += int(N)
        return np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res])

    def _cat(self, name):
        parts = []
        for smp in self._samplers:
            v = getattr(smp, name)
            parts.append(v[None] if smp._lc.ncand == 1 else v)
        return np.concatenate(parts)

    @property
    def chain(self):
        """[ncand, nwalkers, iterations, 6]"""
        return self._cat("chain")

    @property
    def lnprobability(self):
        return self._cat("lnprobability")

    @property
    def naccepted(self):
        return self._cat("naccepted")

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(max(self.iterations, 1))

    def reset(self):
        for smp in self._samplers:
            smp.reset()
        self.iterations = 0

    def close(self):
        self._samplers = []
        for lc in self._lcs:
            lc.close()
        self._lcs = []
