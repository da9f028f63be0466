"""torchrun script: the walker-split sampler on N GPUs must reproduce the single-GPU chain bit for
bit (same seed).  Launched by tests/test_gpu_multi.py or by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import namaste_b200 as nb
    from namaste_b200.distributed import SplitEnsembleSampler
    g = np.load(os.path.join(ROOT, "tests", "golden", "synth_k2_s11.npz"))
    lc = nb.LightCurve(g["lc"], oversamp=int(g["oversamp"]))
    rng = np.random.default_rng(11)
    W, nsteps, seed = 64, 8, 987654321
    pos0 = g["theta"][None, :] * (1 + 1e-4 * rng.standard_normal((W, 6)))
    ok = True
    single = None
    if rank == 0:
        single = nb.EnsembleSampler(W, 6, nb.MonoOnlyLogProb, args=(lc, g["priors"]), seed=seed)
        single.run_mcmc(pos0, nsteps)
    for transport in ("auto", "nccl"):
        split = SplitEnsembleSampler(lc, g["priors"], W, seed, transport=transport)
        split.run_mcmc(pos0, nsteps // 2)
        split.run_mcmc(None, nsteps - nsteps // 2)      # a second call continues the chain (emcee contract)
        chain, lnpc, nacc = split.chain, split.lnprobability, split.naccepted
        if rank == 0:
            same = (np.array_equal(chain, single.chain) and np.array_equal(lnpc, single.lnprobability)
                    and np.array_equal(nacc, single.naccepted))
            ok = ok and same
            print("MULTIGPU_CHECK world=%d transport=%s(%s) bitwise_equal=%s acceptance=%.3f" % (
                world, transport, split.transport, same, nacc.mean() / nsteps), flush=True)
        dist.barrier()
        del split
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
