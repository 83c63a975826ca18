#!/usr/bin/env python
"""Dev tool: a SHORT run of the benchmark workload for Nsight Compute (never a bench number).

    ncu --profile-from-start off ... python tools/ncu_workload.py [--horizon 4] [--iters 1]

Same shapes as bench.py's config (256 envs, preset 1, mb 512, 4 epochs) with a short rollout horizon so one
PPO iteration is a few hundred launches; one warm-up iteration runs before cudaProfilerStart.
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
    ap.add_argument("--horizon", type=int, default=4)
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--feed", default="device")
    a = ap.parse_args()
    args = argparse.Namespace(envs=a.envs, horizon=a.horizon, spec=1, minibatch=512, epochs=4, pool_frames=4096)
    torch.cuda.set_device(0)
    runner = bench.build_runner(args, a.feed, 0, 1)
    itr = 0
    for _ in range(2):
        s, _ = runner.sampler.obtain_samples(itr)
        runner.algo.optimize_policy(itr, s)
        itr += 1
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(a.iters):
        s, _ = runner.sampler.obtain_samples(itr)
        runner.algo.optimize_policy(itr, s)
        itr += 1
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("launches", runner.policy.engine.launches, file=sys.stderr)


if __name__ == "__main__":
    main()
