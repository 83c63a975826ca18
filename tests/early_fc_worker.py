"""Subprocess body of test_early_fc_update_is_bit_identical: two PPO iterations (no global-norm clipping, the PPO
default) through the graph-replayed training path; prints a digest of the final parameters / optimizer state and the
logged gradient norms.  ARL_EARLY_FC (read by the library at first use) selects whether the FC weights take their
Adam step right after their gradient is final or together with everything else."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from accel_rl_b200.algos import PPO                                   # noqa: E402
from accel_rl_b200.envs import AtariEnv                               # noqa: E402
from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs          # noqa: E402
from accel_rl_b200.sampler import ActsrvAltOvrlpSampler               # noqa: E402
from accel_rl_b200.util.seeding import set_seed                       # noqa: E402


def main():
    set_seed(7)
    rules = dict(pool_frames=128, life_base=24, life_mod=11, reward_mod=7, pool_seed=0)
    sampler = ActsrvAltOvrlpSampler(EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules),
                                    horizon=16, n_parallel=4, envs_per=2, max_decorrelation_steps=0)
    env_spec, sample_size, horizon, mbr = sampler.initialize(seed=8, affinities=dict(), discount=0.99, need_extra_obs=True)
    policy = AtariCnnPolicy(**cnn_specs[1])
    policy.initialize(env_spec)
    algo = PPO(optimizer_args=dict(minibatch_size=64, epochs=2), lr_schedule="linear")
    algo.initialize(policy, env_spec, sample_size, horizon, mbr)
    sampler.policy_init(policy)
    algo.set_n_itr(10)
    norms = []
    for itr in range(2):
        buf, _ = sampler.obtain_samples(itr)
        _, info = algo.optimize_policy(itr, buf)
        norms += [float(x) for x in info["GradNorm"]]
    torch.cuda.synchronize()
    eng = policy.engine
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    st = eng.get_opt_state()
    out = dict(params=dig(eng.get_params()), m=dig(st["m"]), v=dig(st["v"]), step=st["step"], norms=norms,
               device_error=eng.device_error())
    if os.environ.get("ARL_AB_DUMP"):        # dev aid: keep the arrays themselves
        np.savez(os.environ["ARL_AB_DUMP"], params=eng.get_params(), m=st["m"], v=st["v"], layout=np.array(eng.layout))
    eng.close()
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
