"""Device-resident vectorised rollout sampler.

Drop-in for ActsrvAltOvrlpSampler (reference: accel_rl/sampler/act_server/alternating/overlap/
sampler.py:20-225, worker.py:23-153, sampler/act_server/buffers.py:7-38): same constructor, same
initialize / policy_init / obtain_samples / shutdown contract, same `samples_buf` structure
(keys, dtypes, row = env*T + t, segs_view, extra_observations) and the same consumption of the
master process's global numpy stream — so identical seeds give identical sampled actions.

What changed underneath: the 2*n_parallel CPU simulator processes, their shared-memory buffers and
semaphores are gone.  All envs are stepped on the GPU: one CUDA graph per rollout replays, for each
of the T steps, conv/FC tcgen05 tiles -> policy head + action sampling -> env step -> fused frame
pipeline, writing straight into the rollout buffers in HBM (no H2D of observations, no D2H of
prob/value).  The two alternating groups survive only as the buffer layout (group j = j-th half of
the envs) and the order in which the uniforms are drawn.
"""
import ctypes as C

import numpy as np
import torch

from accel_rl_b200 import _lib as L
from accel_rl_b200.buffers.batch import (batch_buffer, buffer_with_segs_view, buffer_length,
                                         combine_distinct_buffers, count_buffer_size)
from accel_rl_b200.envs.atari_env import make_frame_pool
from accel_rl_b200.sampler.base import BaseMbSampler
from accel_rl_b200.util.misc import struct, nbytes_unit


class TrajInfo(struct):
    """Completed-episode record (reference: sampler/util.py:75-101)."""

    def __init__(self, discount=1, **kwargs):
        super().__init__(**kwargs)
        self.Length = 0
        self.Return = 0
        self.RawReturn = 0
        self.NonzeroRewards = 0
        self.DiscountedReturn = 0
        self._discount = discount
        self._cur_discount = 1


class ActsrvAltOvrlpSampler(BaseMbSampler):
    def __init__(self, n_parallel=1, envs_per=1, frame_feed="device", host_ring_steps=8, **kwargs):
        super().__init__(n_parallel=n_parallel, envs_per=envs_per, **kwargs)
        self._total_n_envs = n_parallel * envs_per * 2
        if frame_feed not in ("device", "host"):
            raise ValueError("frame_feed must be 'device' or 'host'")
        self.frame_feed = frame_feed
        self.host_ring_steps = host_ring_steps
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------ API to the runner ----
    def initialize(self, seed, affinities, discount=1, need_extra_obs=False, worker_process_target=None):
        env = self.EnvCls(**self.env_args)              # example env (draws its start no-ops)
        if not getattr(env, "device_resident", False):
            raise TypeError("ActsrvAltOvrlpSampler (B200) needs a device-resident EnvCls (accel_rl_b200.envs.AtariEnv)")
        self._env = env
        self.discount = 1 if discount is None else discount
        self.need_extra_obs = need_extra_obs
        B, T = self._total_n_envs, self.horizon
        self.sample_size = B * T
        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        # build_env_buffer (buffers.py:7-21): env.reset(), env.step(action_space.sample())
        env.reset()
        env.spec.action_space.sample()
        P, H, W = env.observation_space.shape
        env_info = dict()
        if env.clip_reward:
            env_info["raw_reward"] = np.float32(0)
        if env.episodic_lives:
            env_info["need_reset"] = False
        examples = dict(observations=np.zeros((P, H, W), np.uint8), rewards=np.float32(0), dones=False,
                        env_infos=env_info)
        self.envs_buf = buffer_with_segs_view(examples, self.sample_size, T, dev)
        if need_extra_obs:
            self.envs_buf.extra_observations = torch.zeros((B, P, H, W), dtype=torch.uint8, device=dev)
        # two step buffers (buffers.py:24-30): obs sample + action sample each; here one buffer whose halves
        # are the two groups
        for _ in range(2):
            env.spec.observation_space.sample()
            env.spec.action_space.sample()
        self.step_buf = struct(obs=torch.zeros((B, P, H, W), dtype=torch.uint8, device=dev))
        assert self.sample_size == buffer_length(self.envs_buf)
        self.env_spec = env.spec
        self.seed = seed
        return env.spec, self.sample_size, self.horizon, self.mid_batch_reset

    def policy_init(self, policy):
        self.policy = policy
        B, T = self._total_n_envs, self.horizon
        policy.reserve(B)
        # build_policy_buffer (buffers.py:33-38): one get_action on a sampled observation
        policy.reset(n_batch=1)
        self.env_spec.observation_space.sample()
        np.random.rand()                                   # weighted_sample inside get_action
        A = self.env_spec.action_space.n
        examples = dict(actions=np.uint8(0), agent_infos=dict(prob=np.zeros(A, np.float32), value=np.float32(0)))
        policy_buf = buffer_with_segs_view(examples, self.sample_size, T, self.device)
        self.samples_buf = combine_distinct_buffers(self.envs_buf, policy_buf)
        policy.reset(n_batch=self.n_parallel * self.envs_per)
        self._configure_engine()

    def _configure_engine(self):
        eng = self.policy.engine
        env, buf = self._env, self.samples_buf
        B, T = self._total_n_envs, self.horizon
        rules = env.synth_rules
        rgb = getattr(env, "frame_mode", "gray") == "rgb"
        self._frame_channels = 3 if rgb else 1
        pool_np = make_frame_pool(rules["pool_frames"], rules.get("pool_seed", 0), channels=self._frame_channels)
        self._pool_host = pool_np
        self.frame_pool = torch.from_numpy(pool_np).to(self.device)
        self._uniforms_host = torch.empty((T, B), dtype=torch.float64).pin_memory()
        self._uniforms = torch.zeros((T, B), dtype=torch.float64, device=self.device)
        if "extra_observations" not in buf:
            self._extra_obs = torch.zeros_like(self.step_buf.obs)
        else:
            self._extra_obs = buf.extra_observations
        raw = buf.env_infos.get("raw_reward")
        nr = buf.env_infos.get("need_reset")
        self._scratch_raw = raw if raw is not None else torch.zeros(B * T, dtype=torch.float32, device=self.device)
        self._scratch_nr = nr if nr is not None else torch.zeros(B * T, dtype=torch.bool, device=self.device)
        cfg = L.SamplerCfg()
        cfg.n_envs, cfg.horizon, cfg.planes = B, T, env.num_img_obs
        p = lambda t: t.data_ptr()
        cfg.observations = p(buf.observations)
        cfg.rewards = p(buf.rewards)
        cfg.dones = p(buf.dones)
        cfg.raw_reward = p(self._scratch_raw)
        cfg.need_reset = p(self._scratch_nr)
        cfg.actions = p(buf.actions)
        cfg.prob = p(buf.agent_infos.prob)
        cfg.value = p(buf.agent_infos.value)
        cfg.extra_observations = p(self._extra_obs)
        cfg.step_obs = p(self.step_buf.obs)
        cfg.uniforms = p(self._uniforms)
        cfg.frame_pool = p(self.frame_pool)
        cfg.pool_frames = int(rules["pool_frames"])
        mpl = self.max_path_length
        cfg.max_path_length = int(min(mpl, 2 ** 31 - 1)) if np.isfinite(mpl) else 2 ** 31 - 1
        cfg.discount = float(self.discount)
        cfg.mid_batch_reset = int(bool(self.mid_batch_reset))
        cfg.clip_reward = int(bool(env.clip_reward))
        cfg.episodic_lives = int(bool(env.episodic_lives))
        for k in ("lives0", "life_base", "life_mul", "life_mod", "reward_mod", "frame_stride"):
            setattr(cfg, k, int(rules[k]))
        cfg.n_games = int(rules.get("n_games", 1))
        cfg.frame_mode = 1 if rgb else 0
        cfg.traj_cap = max(4 * B, 1024)
        self._traj_cap = cfg.traj_cap
        eng.sampler_configure(cfg, keep=(buf, self.step_buf, self.frame_pool, self._uniforms, self._extra_obs))
        eng.sampler_reset()                                # start_envs (sampler/util.py:26-57): env.reset() for every env
        self._decorrelate(eng, B)
        if self.frame_feed == "host":
            self._init_host_feed()
        torch.cuda.synchronize(self.device)

    def _decorrelate(self, eng, B):
        """start_envs with max_decorrelation_steps > 0 (sampler/util.py:33-55): every env takes
        int(fraction * max_decorrelation_steps) warm-up steps and is reset whenever its trajectory ends, so episodes do not
        end in lockstep.  The reference takes `fraction` from the wall clock inside each worker process
        (get_random_fraction, sampler/util.py:22-23); here it comes from a RandomState seeded by the sampler seed, which
        keeps runs reproducible and stays off the master's global stream (as the workers' draws do)."""
        mds = int(self.max_decorrelation_steps or 0)
        self.decorrelation_steps = np.zeros(B, np.int32)
        if mds <= 0:
            return
        if self.frame_feed == "host":
            raise NotImplementedError("frame_feed='host' replays a fixed ring of raw frames; use max_decorrelation_steps=0")
        rng = np.random.RandomState((int(self.seed) * 7919 + 17) % (2 ** 31 - 1))
        self.decorrelation_steps = (rng.rand(B) * mds).astype(np.int32)
        eng.sampler_warmup(torch.from_numpy(self.decorrelation_steps).to(self.device))

    def _init_host_feed(self):
        """Pinned ring of raw-frame step batches (what CPU emulator workers would have written) + a
        double-buffered device staging area; the H2D copy of every step runs on its own stream."""
        B = self._total_n_envs
        S = self.host_ring_steps
        rng = np.random.RandomState(1234)
        fshape = (210, 160) if self._frame_channels == 1 else (210, 160, 3)
        self._ring_host = torch.from_numpy(rng.randint(0, 256, (S, B, 2) + fshape, dtype=np.uint8)).pin_memory()
        self._staging = torch.zeros((2, B, 2) + fshape, dtype=torch.uint8, device=self.device)
        self._copy_stream = torch.cuda.Stream(self.device)
        self._copied = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._act_host = torch.empty((self.horizon, B), dtype=torch.uint8).pin_memory()
        self._act_step = torch.zeros((B,), dtype=torch.uint8, device=self.device)

    def obtain_samples(self, itr):
        eng = self.policy.engine
        B, T = self._total_n_envs, self.horizon
        # serve_actions draws rand(B/2) per group per step (sampler.py:139 -> special.py:24): the same
        # T*B doubles, in the same order, drawn up front
        self._uniforms_host.copy_(torch.from_numpy(np.random.rand(T * B).reshape(T, B)))
        self._uniforms.copy_(self._uniforms_host, non_blocking=True)
        self.h2d_bytes += T * B * 8
        if self.frame_feed == "device":
            eng.rollout_run()
        else:
            self._rollout_host_fed(eng)
        return self._finish_rollout(eng)

    def _finish_rollout(self, eng):
        env, ln, ret, raw, nz, disc = eng.traj_read(self._traj_cap)   # synchronises the stream
        self.d2h_bytes += 4 + 24 * len(env)
        if eng.device_error():
            raise RuntimeError("device-side watchdog fired (code %d)" % eng.device_error())
        traj_infos = []
        for i in range(len(env)):
            ti = TrajInfo(self.discount)
            ti.Length, ti.Return, ti.RawReturn = int(ln[i]), float(ret[i]), float(raw[i])
            ti.NonzeroRewards, ti.DiscountedReturn = int(nz[i]), float(disc[i])
            ti.env = int(env[i])
            traj_infos.append(ti)
        return self.samples_buf, traj_infos

    def _rollout_host_fed(self, eng):
        B, T = self._total_n_envs, self.horizon
        S = self.host_ring_steps
        main = torch.cuda.current_stream(self.device)
        eng.rollout_begin()
        for s in range(T):
            slot = s & 1
            with torch.cuda.stream(self._copy_stream):
                if s >= 2:
                    self._copy_stream.wait_event(self._consumed[slot])
                self._staging[slot].copy_(self._ring_host[s % S], non_blocking=True)
                self._copied[slot].record(self._copy_stream)
            main.wait_event(self._copied[slot])
            eng.rollout_step(s, self._staging[slot])
            self._consumed[slot].record(main)
            # actions of this step go back to the host (what emulator workers would consume)
            self._act_host[s].copy_(self.samples_buf.actions[s::T], non_blocking=True)
        eng.rollout_end()
        self.h2d_bytes += T * B * 2 * 210 * 160 * self._frame_channels
        self.d2h_bytes += T * B

    def shutdown(self):
        pass

    shutdown_worker = shutdown

    @property
    def alternating(self):
        return True


DeviceSampler = ActsrvAltOvrlpSampler
