#!/usr/bin/env python
"""Dev tool: a SMALL pass over every kernel family of the path for compute-sanitizer (memcheck / racecheck / initcheck /
synccheck).  Sizes are tiny (8 envs, 2 rollout steps, one 16-row minibatch) because the tools slow kernels by 10-100x.

    compute-sanitizer --tool memcheck python tools/sanitize_workload.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    torch.cuda.set_device(0)
    from accel_rl_b200.algos import PPO, A2C
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.sampler import ActsrvAltOvrlpSampler
    from accel_rl_b200.util.seeding import set_seed
    for algo_name in ("ppo", "a2c"):
        set_seed(0)
        rules = dict(pool_frames=32, life_base=3, life_mul=3, life_mod=5, reward_mod=5, pool_seed=0)
        sampler = ActsrvAltOvrlpSampler(EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules),
                                        horizon=2, n_parallel=2, envs_per=2, max_decorrelation_steps=3,
                                        mid_batch_reset=(algo_name == "ppo"))
        env_spec, sample_size, horizon, mbr = sampler.initialize(seed=1, affinities=dict(), discount=0.99, need_extra_obs=True)
        policy = AtariCnnPolicy(**cnn_specs[1])
        policy.initialize(env_spec)
        algo = PPO(optimizer_args=dict(minibatch_size=16, epochs=1)) if algo_name == "ppo" else A2C()
        algo.initialize(policy, env_spec, sample_size, horizon, mbr)
        sampler.policy_init(policy)
        for itr in range(2):
            buf, _ = sampler.obtain_samples(itr)
            _, info = algo.optimize_policy(itr, buf)
        torch.cuda.synchronize()
        eng = policy.engine
        # standalone entry points: frame kernels, sampling, plain forward with a gather
        n = 5
        raw_a = torch.randint(0, 256, (n, 210, 160), dtype=torch.uint8, device="cuda")
        raw_b = torch.randint(0, 256, (n, 210, 160), dtype=torch.uint8, device="cuda")
        stack = torch.zeros(n, 4, 104, 80, dtype=torch.uint8, device="cuda")
        eng.frame_update(raw_a, raw_b, None, stack)
        prob = torch.softmax(torch.randn(n, 4, device="cuda"), dim=1).contiguous()
        acts = torch.zeros(n, dtype=torch.uint8, device="cuda")
        eng.sample_actions(prob, torch.rand(n, dtype=torch.float64, device="cuda"), acts)
        pr = torch.zeros(n, 4, device="cuda"); va = torch.zeros(n, device="cuda")
        eng.forward(stack, prob=pr, value=va)
        torch.cuda.synchronize()
        assert eng.device_error() == 0 and np.isfinite(info["GradNorm"]).all()
        eng.close()
    print("SANITIZE_WORKLOAD_OK")


if __name__ == "__main__":
    main()
