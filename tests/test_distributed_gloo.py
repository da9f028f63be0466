"""Host-side multi-rank logic on CPU (gloo, world_size 2): candidate sharding and the walker-split
half-step protocol.  The per-rank 'kernel' is the NumPy stretch oracle restricted to the rank's
slice; two ranks must reproduce the single-process chain bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from namaste_b200.distributed import WalkerSplit, allgather_half, shard_candidates
from oracle import stretch_oracle as so

SIG = np.array([1.0, 0.3, 2.0, 0.5, 1.5, 0.7])


def lnp_fn(p):
    return -0.5 * np.sum((p / SIG) ** 2, axis=1)


def half_update(pos, lnp, half, k0, nk, seed, step, a=2.0):
    """Update walkers [k0, k0+nk) of half `half` in place (what nmb_sampler_halfstep does on the GPU)."""
    W, dim = pos.shape
    halfk = W // 2
    rows = np.arange(half * halfk + k0, half * halfk + k0 + nk)
    comp = pos[halfk:] if half == 0 else pos[:halfk]
    uz, uj, ua = so.draws(seed, step, half, rows)
    zz = ((a - 1.) * uz + 1) ** 2. / a
    j = np.minimum((uj * len(comp)).astype(int), len(comp) - 1)
    q = comp[j] - zz[:, None] * (comp[j] - pos[rows])
    new = lnp_fn(q)
    acc = ((dim - 1.) * np.log(zz) + new - lnp[rows]) > np.log(ua)
    pos[rows[acc]] = q[acc]
    lnp[rows[acc]] = new[acc]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, pos0, nsteps, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = pos0.shape[0]
    sp = WalkerSplit(W, world, rank)
    pos = torch.from_numpy(pos0.copy())
    lnp = torch.from_numpy(lnp_fn(pos0))
    chain = np.zeros((W, nsteps, pos0.shape[1]))
    for step in range(nsteps):
        for half in (0, 1):
            half_update(pos.numpy(), lnp.numpy(), half, sp.k0, sp.nk, seed, step)
            allgather_half(pos, lnp, half, sp)
        chain[:, step, :] = pos.numpy()
    if rank == 0:
        np.save(out, chain)
    dist.barrier()
    dist.destroy_process_group()


def test_walker_split_two_ranks_reproduce_single_process_chain(tmp_path):
    rng = np.random.default_rng(5)
    pos0 = rng.standard_normal((16, 6)) * SIG
    nsteps, seed = 12, 4242
    ref, _, _ = so.run(lnp_fn, pos0, nsteps, seed)
    out = str(tmp_path / "chain.npy")
    mp.spawn(_worker, args=(2, _free_port(), pos0, nsteps, seed, out), nprocs=2, join=True)
    assert np.array_equal(np.load(out), ref)


def test_slices_and_candidate_sharding():
    sp = [WalkerSplit(32, 4, r) for r in range(4)]
    covered = sorted(i for s in sp for half in (0, 1) for i in range(s.rows(half)[0], sum(s.rows(half))))
    assert covered == list(range(32))
    with pytest.raises(ValueError):
        WalkerSplit(30, 4, 0)
    with pytest.raises(ValueError):
        WalkerSplit(31, 1, 0)
    owners = [shard_candidates(2048, r, 8) for r in range(8)]
    assert all(len(o) == 256 for o in owners)
    assert sorted(k for o in owners for k in o) == list(range(2048))
