"""The walker-split leg of bench.py alone (BASELINE configs[4] over the ranks of a torchrun launch): steps/s and the
bitwise check against the single-GPU chain.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gpu_split_leg.py [steps]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
args = argparse.Namespace(gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=steps, warmup=3, profile=False, no_cpu=True,
                          workload="auto", impl="ours")
b = bench.Bench(args)
res = [bench.measure_split(b, n_steps=steps) for _ in range(2)]
if b.rank == 0:
    for out in res:
        print("SPLIT_LEG world=%d steps/s %.1f ms/step %.4f bitwise_equal=%s replicas_identical=%s" % (
            b.world, out["steps_per_s"], out["ms_per_step"], out.get("bitwise_equal"),
            out["replicas_identical"]), flush=True)
if b.world > 1:
    b.dist.destroy_process_group()
