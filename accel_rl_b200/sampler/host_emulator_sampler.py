"""Rollout sampler fed by emulators running in host worker processes.

The reference's ActsrvAltOvrlpSampler (sampler/act_server/alternating/overlap/sampler.py:20-225) with the labour
re-divided for a B200: 2*n_parallel simulator processes in two alternating groups still own the emulators and the
env/collector logic (accel_rl_b200/hostsim/worker.py), but they export RAW SCREENS; max / resize / stack, the policy
and the rollout buffers stay on the GPU.  Per step and group j (sampler.py:130-145 serve_actions):

    wait step_done of group j          (its emulators finished the previous step; the other group is still emulating)
    H2D   raw frame pairs + step records of group j   (pinned shared memory -> HBM, no bounce buffer)
    GPU   arl_rollout_ingest: rewards/dones/infos -> rows, fused frame kernel -> step buffer + rollout row
    GPU   arl_rollout_serve:  conv/FC tiles -> head -> sampled actions
    D2H   the group's actions, then release act_ready of group j

so each group's emulation overlaps the other group's time on the GPU, as in the reference.  Same constructor as the
device sampler plus `emu_factory(env_index) -> emulator` (default: real ALE through atari_py / ale_py) — EnvCls/env_args
still describe the env (game, clip_reward, episodic_lives, max_start_noops, num_img_obs, frame_mode).
"""
import ctypes as C
import multiprocessing as mp
from functools import partial

import numpy as np
import torch

from accel_rl_b200 import _lib as L
from accel_rl_b200.hostsim import worker as W
from accel_rl_b200.hostsim.atari_env import make_ale
from accel_rl_b200.sampler.device_sampler import ActsrvAltOvrlpSampler, TrajInfo


class HostEmulatorSampler(ActsrvAltOvrlpSampler):
    def __init__(self, emu_factory=None, **kwargs):
        kwargs.pop("frame_feed", None)
        super().__init__(**kwargs)
        self.emu_factory = emu_factory
        self._procs = []

    # the frame pool / synthetic-emulator configuration of the base class is replaced by the worker processes
    def _configure_engine(self):
        eng = self.policy.engine
        env, buf = self._env, self.samples_buf
        B, T = self._total_n_envs, self.horizon
        rgb = getattr(env, "frame_mode", "gray") == "rgb"
        self._frame_shape = (210, 160, 3) if rgb else (210, 160)
        self._uniforms_host = torch.empty((T, B), dtype=torch.float64).pin_memory()
        self._uniforms = torch.zeros((T, B), dtype=torch.float64, device=self.device)
        self._extra_obs = buf.extra_observations if "extra_observations" in buf else torch.zeros_like(self.step_buf.obs)
        raw = buf.env_infos.get("raw_reward")
        nr = buf.env_infos.get("need_reset")
        self._scratch_raw = raw if raw is not None else torch.zeros(B * T, dtype=torch.float32, device=self.device)
        self._scratch_nr = nr if nr is not None else torch.zeros(B * T, dtype=torch.bool, device=self.device)
        self._dummy_pool = torch.zeros(16, dtype=torch.uint8, device=self.device)
        cfg = L.SamplerCfg()
        cfg.n_envs, cfg.horizon, cfg.planes = B, T, env.num_img_obs
        p = lambda t: t.data_ptr()
        cfg.observations, cfg.rewards, cfg.dones = p(buf.observations), p(buf.rewards), p(buf.dones)
        cfg.raw_reward, cfg.need_reset = p(self._scratch_raw), p(self._scratch_nr)
        cfg.actions, cfg.prob, cfg.value = p(buf.actions), p(buf.agent_infos.prob), p(buf.agent_infos.value)
        cfg.extra_observations = p(self._extra_obs)
        cfg.step_obs = p(self.step_buf.obs)
        cfg.uniforms = p(self._uniforms)
        cfg.frame_pool, cfg.pool_frames = p(self._dummy_pool), 1
        mpl = self.max_path_length
        cfg.max_path_length = int(min(mpl, 2 ** 31 - 1)) if np.isfinite(mpl) else 2 ** 31 - 1
        cfg.discount = float(self.discount)
        cfg.mid_batch_reset = int(bool(self.mid_batch_reset))
        cfg.clip_reward = int(bool(env.clip_reward))
        cfg.episodic_lives = int(bool(env.episodic_lives))
        cfg.lives0 = cfg.life_base = cfg.life_mul = cfg.life_mod = cfg.reward_mod = cfg.frame_stride = 1
        cfg.frame_mode = 1 if rgb else 0
        cfg.ext_emulator = 1
        cfg.traj_cap = 16
        eng.sampler_configure(cfg, keep=(buf, self.step_buf, self._uniforms, self._extra_obs, self._dummy_pool))
        self._start_workers()
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ worker processes -----
    def _start_workers(self):
        eng, env = self.policy.engine, self._env
        B = self._total_n_envs
        n_workers = 2 * self.n_parallel
        ctx = mp.get_context("spawn")            # never fork a process that holds a CUDA context
        fbytes = int(np.prod(self._frame_shape))
        self._shared = dict(frames=ctx.RawArray(C.c_uint8, B * 2 * fbytes), ext=ctx.RawArray(C.c_uint8, B * W.EXT_DTYPE.itemsize),
                            act=ctx.RawArray(C.c_uint8, B), report=ctx.RawArray(C.c_int32, 2 * n_workers),
                            arg=ctx.RawArray(C.c_int32, 1))
        self._frames_np, self._ext_np, self._act_np = W.views(self._shared, B, self._frame_shape)
        # page-lock the shared blocks so the copies below are real DMA transfers
        self._pinned = []
        for k in ("frames", "ext", "act"):
            addr, n = C.addressof(self._shared[k]), C.sizeof(self._shared[k])
            if eng.lib.arl_host_register(C.c_void_p(addr), n) != 0:
                raise RuntimeError("cudaHostRegister failed for the shared %s block" % k)
            self._pinned.append(addr)
        self._staging = torch.zeros((B, 2) + self._frame_shape, dtype=torch.uint8, device=self.device)
        self._ext_dev = torch.zeros(B * W.EXT_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        self._act_dev = torch.zeros(B, dtype=torch.uint8, device=self.device)
        self._cmd = ctx.Value("i", W.CMD_STEP, lock=False)
        self._act_ready = [ctx.Semaphore(0) for _ in range(n_workers)]
        self._step_done = [ctx.Semaphore(0) for _ in range(n_workers)]
        self._infos_queue = ctx.Queue()
        factory = self.emu_factory
        if factory is None:
            factory = partial(_ale_for_env, game=env.game, repeat_action_probability=env.repeat_action_probability)
        env_kwargs = dict(frame_skip=4, clip_reward=bool(env.clip_reward), episodic_lives=bool(env.episodic_lives),
                          max_start_noops=int(env.max_start_noops), rgb=len(self._frame_shape) == 3)
        for w in range(n_workers):
            lo, hi = w * self.envs_per, (w + 1) * self.envs_per
            args = (w, lo, hi, B, factory, env_kwargs, self._frame_shape, self._shared, self._cmd,
                    self._act_ready[w], self._step_done[w], self._infos_queue, self.seed + w,
                    bool(self.mid_batch_reset), self.max_path_length, float(self.discount),
                    int(self.max_decorrelation_steps or 0))
            if self.profile_pathname is not None:        # cProfile per simulator worker (sampler/util.py:10-19)
                pr = ctx.Process(target=W.profiling_worker, daemon=True, args=(self.profile_pathname,) + args)
            else:
                pr = ctx.Process(target=W.worker_main, daemon=True, args=args)
            pr.start()
            self._procs.append(pr)
        # start_envs: every worker resets its envs and hands over the first screens
        for g in range(2):
            self._wait_group(g)
            self._ingest_group(-1, g)
        report = np.frombuffer(self._shared["report"], dtype=np.int32).reshape(-1, 2)
        # the policy's action count must be the emulators' (the workers index getMinimalActionSet() with it)
        A = self.env_spec.action_space.n
        bad = [(w, int(a)) for w, a in enumerate(report[:, 0]) if int(a) != A]
        if bad:
            self.shutdown()
            raise ValueError("the env spec has %d actions but emulator worker(s) %s report a minimal action set of %s: "
                             "pass n_actions / game to EnvCls so both agree" % (A, [w for w, _ in bad], [a for _, a in bad]))
        # start_envs decorrelation (sampler/util.py:33-55): warm-up rounds, every env for its own number of steps
        self.decorrelation_rounds = int(report[:, 1].max())
        for k in range(self.decorrelation_rounds):
            self._shared["arg"][0] = k
            torch.cuda.current_stream(self.device).synchronize()    # the previous round's frames have been consumed
            for g in range(2):
                self._release_group(g, W.CMD_WARM)
            for g in range(2):
                self._wait_group(g)
                self._ingest_group(-1, g)
        self._pending = [False, False]           # no emulation outstanding

    def _group_range(self, g):
        half = self._total_n_envs // 2
        return g * half, half

    def _group_workers(self, g):
        return range(g * self.n_parallel, (g + 1) * self.n_parallel)

    def _wait_group(self, g, timeout=300):
        for w in self._group_workers(g):
            waited = 0
            while not self._step_done[w].acquire(timeout=2):
                waited += 2
                if not self._procs[w].is_alive():
                    raise RuntimeError("emulator worker %d died (exit code %s)" % (w, self._procs[w].exitcode))
                if waited >= timeout:
                    raise RuntimeError("emulator worker %d did not answer within %d s" % (w, timeout))

    def _release_group(self, g, cmd=W.CMD_STEP):
        self._cmd.value = cmd
        for w in self._group_workers(g):
            self._act_ready[w].release()

    def _ingest_group(self, s, g):
        """H2D of the group's raw frames + records, then rows + frame pipeline on the GPU"""
        eng = self.policy.engine
        e0, n = self._group_range(g)
        fb = self._frames_np[0].nbytes
        eng.copy_async(self._staging.data_ptr() + e0 * fb, self._frames_np.ctypes.data + e0 * fb, n * fb, True)
        rb = W.EXT_DTYPE.itemsize
        eng.copy_async(self._ext_dev.data_ptr() + e0 * rb, self._ext_np.ctypes.data + e0 * rb, n * rb, True)
        eng.rollout_ingest(s, self._staging, self._ext_dev, e0, n)
        self.h2d_bytes += n * (fb + rb)

    def _serve_group(self, s, g):
        """policy forward + sampling for the group, its actions back to the host, then its emulators go"""
        eng = self.policy.engine
        e0, n = self._group_range(g)
        T = self.horizon
        eng.rollout_serve(s, e0, n)
        self._act_dev[e0:e0 + n].copy_(self.samples_buf.actions[e0 * T + s:(e0 + n) * T:T])
        eng.copy_async(self._act_np.ctypes.data + e0, self._act_dev.data_ptr() + e0, n, False)
        torch.cuda.current_stream(self.device).synchronize()      # actions are on the host; frames of this group consumed
        self.d2h_bytes += n
        self._release_group(g)

    # ------------------------------------------------------------------ rollout ----------------
    def obtain_samples(self, itr):
        eng = self.policy.engine
        B, T = self._total_n_envs, self.horizon
        self._uniforms_host.copy_(torch.from_numpy(np.random.rand(T * B).reshape(T, B)))
        self._uniforms.copy_(self._uniforms_host, non_blocking=True)
        self.h2d_bytes += T * B * 8
        eng.rollout_begin()
        for s in range(T):
            for g in range(2):
                if s > 0:
                    self._wait_group(g)                 # emulation of step s-1 done
                    self._ingest_group(s - 1, g)
                self._serve_group(s, g)
        for g in range(2):
            self._wait_group(g)
            self._ingest_group(T - 1, g)
        eng.rollout_end()
        if not self.mid_batch_reset:                    # reset_needed_envs after barrier_out (worker.py:150-151)
            torch.cuda.current_stream(self.device).synchronize()
            for g in range(2):
                self._release_group(g, W.CMD_RESET_NEEDED)
                self._wait_group(g)
                self._ingest_group(T, g)
        torch.cuda.current_stream(self.device).synchronize()
        if eng.device_error():
            raise RuntimeError("device-side watchdog fired (code %d)" % eng.device_error())
        traj_infos = []
        while not self._infos_queue.empty():
            e, ln, ret, raw, nz, disc = self._infos_queue.get()
            ti = TrajInfo(self.discount)
            ti.Length, ti.Return, ti.RawReturn, ti.NonzeroRewards, ti.DiscountedReturn = ln, ret, raw, nz, disc
            ti.env = e
            traj_infos.append(ti)
        return self.samples_buf, traj_infos

    def shutdown(self):
        if self._procs:
            self._cmd.value = W.CMD_QUIT
            for s in self._act_ready:
                s.release()
            for pr in self._procs:
                pr.join(timeout=10)
                if pr.is_alive():
                    pr.terminate()
            self._procs = []
            lib = self.policy.engine.lib
            for addr in self._pinned:
                lib.arl_host_unregister(C.c_void_p(addr))
            self._pinned = []

    shutdown_worker = shutdown


def _ale_for_env(env_index, game, repeat_action_probability):
    return make_ale(game, repeat_action_probability)
