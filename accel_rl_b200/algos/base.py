"""What a Runner needs from an RL algorithm (the reference's statement of it: accel_rl/algos/base.py:3-13).

initialize(policy, env_spec, sample_size, horizon, mid_batch_reset)   bind to the policy and size the optimizer's buffers
optimize_policy(itr, samples_data) -> (opt_data, opt_infos)           one learning step on one batch of rollouts
opt_info_keys                                                         names of the per-update diagnostics in opt_infos
"""


class RLAlgorithm(object):
    opt_info_keys = ()

    def initialize(self, policy, env_spec, sample_size, horizon, mid_batch_reset):
        raise NotImplementedError(type(self).__name__ + ".initialize")

    def optimize_policy(self, itr, samples_data):
        raise NotImplementedError(type(self).__name__ + ".optimize_policy")
