"""ActsrvAltOvrlpPollSampler (reference: accel_rl/sampler/act_server/alternating/overlap/poll_sampler.py:6-56).

The asynchronous learners' sampler: every `poll_horizon` rollout steps — before serving step s whenever
(s + 1) % poll_horizon == 0 (poll_sampler.py:29) — the policy is refreshed from the CENTRAL parameters, so a long rollout
does not act on weights that the other learners have already moved on from.  The reference copies the shared-memory
vector into the policy under the parameter lock; here the central store lives in rank 0's HBM and the refresh is one
kernel that copies it region by region under the device-side chunk locks (csrc/comm.cuh async_pull_kernel), after which
the bf16 operand copies are re-packed.  Between refreshes the rollout steps are the same kernels as the graph-replayed
rollout of ActsrvAltOvrlpSampler, launched step by step."""
import numpy as np
import torch

from accel_rl_b200.sampler.device_sampler import ActsrvAltOvrlpSampler


class CentralParams(object):
    """what `poll_init` receives as `central_shared_params`: a handle on the central store of the engine's asynchronous
    optimizer (`BaseAsyncOptimizer.central_params_handle`)"""

    def __init__(self, engine):
        self.engine = engine

    def pull(self):
        self.engine.async_pull()

    def numpy(self):
        return self.engine.async_read_central(0)


class ActsrvAltOvrlpPollSampler(ActsrvAltOvrlpSampler):
    def __init__(self, poll_horizon, **kwargs):
        super().__init__(**kwargs)
        if int(poll_horizon) < 1:
            raise ValueError("poll_horizon must be >= 1")
        self._poll_horizon = int(poll_horizon)
        self._central_shared_params = None
        self.n_polls = 0

    def poll_init(self, central_shared_params, params_lock=None, params_rwlock=None):
        """reference signature (poll_sampler.py:11-14).  `central_shared_params`: a CentralParams handle (the locks are on
        the device: `params_lock` / `params_rwlock` are accepted and unused) or a host vector, which is then loaded with
        policy.set_param_values exactly as the reference does."""
        self._central_shared_params = central_shared_params
        self._params_lock = params_lock
        self._params_rwlock = params_rwlock

    def _poll(self):
        src = self._central_shared_params
        if hasattr(src, "pull"):
            src.pull()
        else:
            self.policy.set_param_values(np.asarray(src), trainable=True)
        self.n_polls += 1

    def obtain_samples(self, itr):
        if self._central_shared_params is None:
            raise RuntimeError("poll_init(central_shared_params, params_lock) must be called before sampling")
        if self.frame_feed != "device":
            raise NotImplementedError("the poll sampler serves the device-resident emulator feed")
        eng = self.policy.engine
        B, T = self._total_n_envs, self.horizon
        self._uniforms_host.copy_(torch.from_numpy(np.random.rand(T * B).reshape(T, B)))
        self._uniforms.copy_(self._uniforms_host, non_blocking=True)
        self.h2d_bytes += T * B * 8
        eng.rollout_begin()
        for s in range(T):                                   # serve_actions (poll_sampler.py:28-50)
            if (s + 1) % self._poll_horizon == 0:
                self._poll()
            eng.rollout_step(s)
        eng.rollout_end()
        return self._finish_rollout(eng)
