#!/usr/bin/env python
"""Dev tool: completion times of every kernel of one training minibatch inside the product's forked graph.

    python tools/timeline.py [--reps 5]          (1 GPU, local update)
    torchrun --nproc-per-node N tools/timeline.py --sync       (N GPUs, synchronous step; rank 0 prints)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sync", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args = argparse.Namespace(envs=256, horizon=128, spec=1, minibatch=512, epochs=4, pool_frames=4096, frames="gray",
                              algo="ppo", parallelism="sync")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    runner = bench.build_runner(args, "device", rank, world)
    itr = 0
    for _ in range(2):
        s, _ = runner.sampler.obtain_samples(itr)
        runner.algo.optimize_policy(itr, s)
        itr += 1
    eng = runner.policy.engine
    idx = torch.randperm(256 * 128, device="cuda")[:8 * 512].to(torch.int32).contiguous()
    for r in range(a.reps):
        tl = eng.profile_timeline(1 if a.sync else 0, idx, 512)
        if r == a.reps - 1 and rank == 0:
            for name, t in sorted(tl, key=lambda x: x[1]):
                print("%8.1f us  %s" % (t, name))
    eng.read_logs()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
