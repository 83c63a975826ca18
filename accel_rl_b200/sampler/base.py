"""Sampler contract between a Runner and whatever produces the rollouts.

The runner only ever calls the four methods below plus two read-only attributes; the reference states the same contract
in accel_rl/sampler/base.py:11-47.  `BaseMbSampler` is the constructor every "minibatch" sampler shares — its keyword
names are part of the public surface (experiment scripts pass them by name), so they are kept verbatim.
"""
import numpy as np

from accel_rl_b200.util.quick_args import save_args

_CONTRACT = {
    "initialize": "initialize(seed, affinities, discount, need_extra_obs, ...) -> (env_spec, sample_size, horizon, mid_batch_reset)",
    "policy_init": "policy_init(policy): called once the policy exists; builds the agent half of the rollout buffers",
    "obtain_samples": "obtain_samples(itr) -> (samples_buf, traj_infos): one horizon of every env",
    "shutdown_worker": "shutdown_worker(): release whatever initialize() started",
}


def _required(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError("{} must implement {}".format(type(self).__name__, _CONTRACT[name]))
    method.__name__ = name
    method.__doc__ = _CONTRACT[name]
    return method


class Sampler(object):
    """Subclasses provide the four calls of `_CONTRACT`; `alternating` says whether envs are served in two groups."""

    alternating = False


for _name in _CONTRACT:
    setattr(Sampler, _name, _required(_name))


class BaseMbSampler(Sampler):
    """Common constructor of the vectorised samplers.

    EnvCls / env_args            environment class and its keyword arguments (one instance describes the spaces)
    horizon                      T, time steps per env per obtain_samples
    n_parallel, envs_per         the batch is 2 * n_parallel * envs_per envs (two alternating groups of n_parallel
                                 workers in the reference; the same numbers size the device-resident batch here)
    max_path_length              trajectories longer than this are cut and the env reset
    mid_batch_reset              reset finished envs immediately (True) or leave them idle until the batch ends
    max_decorrelation_steps      random warm-up steps per env at start-up (0 in every benchmark here)
    profile_pathname             if given, emulator worker processes run under cProfile and dump <path>_sim_<rank>.prof
    """

    def __init__(self, EnvCls, env_args, horizon, n_parallel=1, envs_per=1, max_path_length=np.inf,
                 mid_batch_reset=True, max_decorrelation_steps=2000, profile_pathname=None):
        save_args(vars(), underscore=False)

    @property
    def total_n_envs(self):
        return self._total_n_envs
