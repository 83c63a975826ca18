"""Device sampler with offline evaluation episodes.

Drop-in for AAOEvalSampler (reference: accel_rl/sampler/act_server/alternating/overlap/sampler_with_eval.py:6-54,
worker_with_eval.py:66-99,194-229): same constructor (`eval_steps`, `eval_envs_per` ahead of the base sampler's
arguments), same `evaluate_policy(itr) -> traj_infos`.

Semantics kept: a separate set of `eval_envs_per * n_parallel * 2` evaluation envs; every evaluation starts by resetting
all of them with fresh TrajInfos, runs `eval_horizon = eval_steps // n_eval_envs` steps with the current policy
(actions sampled from the master process's global numpy stream, rand(n/2) per group per step like any other serve),
always resets finished envs mid-run, stores no observations, returns only the trajectories COMPLETED inside the run,
and leaves the training envs and their step buffer exactly as they were (the reference restores step_buf.obs; here the
evaluation runs in the engine's second sampler slot and never touches the first).
"""
import numpy as np
import torch

from accel_rl_b200 import _lib as L
from accel_rl_b200.sampler.device_sampler import ActsrvAltOvrlpSampler, TrajInfo
from accel_rl_b200.util.misc import struct


class AAOEvalSampler(ActsrvAltOvrlpSampler):
    def __init__(self, eval_steps, eval_envs_per, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.eval_envs_per = eval_envs_per
        self._total_n_eval_envs = eval_envs_per * self.n_parallel * 2
        self.eval_horizon = int(eval_steps) // self._total_n_eval_envs
        if self.eval_horizon < 1:
            raise ValueError("eval_steps smaller than the number of evaluation envs")

    def policy_init(self, policy):
        need = max(self._total_n_envs, self._total_n_eval_envs)
        eng = getattr(policy, "_engine", None)
        if eng is not None and need > eng.max_rows:
            # growing the engine now would drop the optimizer state the algorithm has already bound
            raise ValueError("policy engine holds %d rows but the evaluation uses %d envs: build the policy with "
                             "max_rows >= %d" % (eng.max_rows, need, need))
        policy.reserve(need)
        super().policy_init(policy)

    def _configure_engine(self):
        super()._configure_engine()
        eng, env = self.policy.engine, self._env
        B, T = self._total_n_eval_envs, self.eval_horizon
        dev = self.device
        P, H, W = env.observation_space.shape
        A = self.env_spec.action_space.n
        N = B * T
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        self.eval_buf = struct(
            rewards=z((N,), torch.float32), dones=z((N,), torch.bool), raw_reward=z((N,), torch.float32),
            need_reset=z((N,), torch.bool), actions=z((N,), torch.uint8), prob=z((N, A), torch.float32),
            value=z((N,), torch.float32))
        self.eval_step_buf = struct(obs=z((B, P, H, W), torch.uint8))
        self._eval_uniforms_host = torch.empty((T, B), dtype=torch.float64).pin_memory()
        self._eval_uniforms = z((T, B), torch.float64)
        rules = env.synth_rules
        cfg = L.SamplerCfg()
        cfg.n_envs, cfg.horizon, cfg.planes = B, T, env.num_img_obs
        p = lambda t: t.data_ptr()
        eb = self.eval_buf
        cfg.observations = None                           # serve_actions_eval stores no state-action-agent info
        cfg.extra_observations = None
        cfg.rewards, cfg.dones, cfg.raw_reward, cfg.need_reset = p(eb.rewards), p(eb.dones), p(eb.raw_reward), p(eb.need_reset)
        cfg.actions, cfg.prob, cfg.value = p(eb.actions), p(eb.prob), p(eb.value)
        cfg.step_obs = p(self.eval_step_buf.obs)
        cfg.uniforms = p(self._eval_uniforms)
        cfg.frame_pool = p(self.frame_pool)
        cfg.pool_frames = int(rules["pool_frames"])
        mpl = self.max_path_length
        cfg.max_path_length = int(min(mpl, 2 ** 31 - 1)) if np.isfinite(mpl) else 2 ** 31 - 1
        cfg.discount = float(self.discount)
        cfg.mid_batch_reset = 1                           # collect_eval always resets (worker_with_eval.py:84-91)
        cfg.clip_reward = int(bool(env.clip_reward))
        cfg.episodic_lives = int(bool(env.episodic_lives))
        for k in ("lives0", "life_base", "life_mul", "life_mod", "reward_mod", "frame_stride"):
            setattr(cfg, k, int(rules[k]))
        cfg.n_games = int(rules.get("n_games", 1))
        cfg.frame_mode = 1 if self._frame_channels == 3 else 0
        cfg.traj_cap = max(4 * B, 1024, 2 * N // 16)
        self._eval_traj_cap = cfg.traj_cap
        eng.sampler_select(1)
        try:
            eng.sampler_configure(cfg, keep=(self.eval_buf, self.eval_step_buf, self._eval_uniforms, self.frame_pool))
        finally:
            eng.sampler_select(0)
        torch.cuda.synchronize(dev)

    def evaluate_policy(self, itr):
        eng = self.policy.engine
        B, T = self._total_n_eval_envs, self.eval_horizon
        self._eval_uniforms_host.copy_(torch.from_numpy(np.random.rand(T * B).reshape(T, B)))
        self._eval_uniforms.copy_(self._eval_uniforms_host, non_blocking=True)
        eng.sampler_select(1)
        try:
            eng.sampler_reset()                           # collect_eval: env.reset() for every eval env, fresh TrajInfos
            eng.rollout_run()
            env, ln, ret, raw, nz, disc = eng.traj_read(self._eval_traj_cap)
        finally:
            eng.sampler_select(0)
        if eng.device_error():
            raise RuntimeError("device-side watchdog fired (code %d)" % eng.device_error())
        traj_infos = []
        for i in range(len(env)):
            ti = TrajInfo(self.discount)
            ti.Length, ti.Return, ti.RawReturn = int(ln[i]), float(ret[i]), float(raw[i])
            ti.NonzeroRewards, ti.DiscountedReturn = int(nz[i]), float(disc[i])
            ti.env = int(env[i])
            traj_infos.append(ti)
        return traj_infos
