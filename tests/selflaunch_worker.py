"""One-script multi-GPU training (tests/test_gpu_multi.py): `python tests/selflaunch_worker.py sync|async WORLD` builds
AccelRLSync / AccelRLAsync with a list of per-GPU affinities and calls train() — the runner forks ranks 1.. itself, as the
reference's launch_workers does (runners/multigpu_rl_base.py:20-45).  Under torch.distributed.run the very same script
trains on the group torchrun made.  Rank 0 prints a digest of the final parameters; the test compares the two launches."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, world = sys.argv[1], int(sys.argv[2])
    import torch
    import torch.distributed as dist
    under_torchrun = "RANK" in os.environ
    if under_torchrun:
        local = int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from accel_rl_b200.algos import mPPO, mAPPO
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRLSync, AccelRLAsync, AccelRLEvalSync
    from accel_rl_b200.sampler import ActsrvAltOvrlpSampler, AAOEvalSampler
    from accel_rl_b200.util import logger
    log_dir = sys.argv[3] if len(sys.argv) > 3 else None
    logger.configure(log_dir, quiet=True)
    rules = dict(pool_frames=128, life_base=24, life_mod=11, reward_mod=7)
    sampler_args = dict(EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules),
                        horizon=16, n_parallel=4, envs_per=4, max_decorrelation_steps=0)
    if mode == "evalsync":                                    # offline evaluation: only the master evaluates and logs
        sampler = AAOEvalSampler(eval_steps=4096, eval_envs_per=2, **sampler_args)
    else:
        sampler = ActsrvAltOvrlpSampler(**sampler_args)
    Algo, Runner = dict(sync=(mPPO, AccelRLSync), evalsync=(mPPO, AccelRLEvalSync), **{"async": (mAPPO, AccelRLAsync)})[mode]
    algo = Algo(optimizer_args=dict(minibatch_size=128, epochs=2))
    policy = AtariCnnPolicy(**cnn_specs[1])
    assert under_torchrun or not torch.cuda.is_initialized()
    interval = dict(eval_interval_steps=512 * world) if mode == "evalsync" else dict(log_interval_steps=512 * 2 * world)
    runner = Runner(algo=algo, policy=policy, sampler=sampler, n_steps=512 * 2 * world, seed=3,
                    affinities=[dict(gpu=i) for i in range(world)], **interval)
    runner.train()                                            # 3 iterations (accel_rl_base.py:84-97 rounding)
    # only rank 0 returns here in the one-script launch (the forked runners exit inside train())
    if runner.rank == 0:
        p = policy.get_param_values()
        assert np.isfinite(p).all()
        print("SELFLAUNCH_DIGEST " + json.dumps(dict(mode=mode, world=world, torchrun=under_torchrun, n_itr=runner._n_itr,
                                                      params=hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest())))
    if under_torchrun:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
