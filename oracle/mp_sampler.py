"""ORACLE (test infrastructure; also the sampler leg of bench.py's reference arm / cpu_baseline).

MULTI-PROCESS restatement of the reference's ActsrvAltOvrlpSampler — the process structure the reference actually
runs, not only its buffer semantics (oracle/sampler.py is the single-process statement of those):

  master      accel_rl/sampler/act_server/alternating/overlap/sampler.py:40-95 (fork 2*n_parallel simulator
              processes in two alternating groups, shared buffers, per-worker semaphores, two barriers),
              :97-104 (obtain_samples), :120-151 (serve_actions: per step, per group: wait for the group's workers,
              policy.get_actions on the group's step buffer, publish actions, release the workers)
  worker      .../overlap/worker.py:116-153 (sampling_process), :25-60 (ResetCollector.collect),
              :63-113 (NonResetCollector), sampler/util.py:26-57 (start_envs without decorrelation), :75-101 (TrajInfo)
  buffers     act_server/buffers.py:7-38, buffers/batch.py:36-76 (row = env*T + t), buffers/array.py:7 (fork-shared
              mp.RawArray behind numpy)

While group j's actions are computed by the master, group 1-j's workers simulate (the "alternating / overlap" of
sampler.py:22-28).  The env is oracle.sampler.SynthAtariEnv (AtariEnv over the synthetic emulator).  Produces buffers
identical to OracleSampler, which is pinned to the real reference sampler's goldens (tests/test_oracle_mp_sampler.py).
"""
import ctypes
import multiprocessing as mp

import numpy as np

from oracle import sampler as osampler


def _shared(shape, dtype):
    """fork-shared numpy array (buffers/array.py:7 np_mp_array)"""
    n = int(np.prod(shape))
    ct = {np.dtype(np.uint8): ctypes.c_uint8, np.dtype(np.float32): ctypes.c_float, np.dtype(bool): ctypes.c_bool}[np.dtype(dtype)]
    raw = mp.RawArray(ct, max(n, 1))
    return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)


def _worker(group, rank, env_ids, pool, rules, T, bufs, step_obs, step_act, sync, cfg):
    """sampling_process + collector (worker.py:116-153, :25-113) for the envs `env_ids` of group `group`"""
    discount, mid_batch_reset, max_path_length = cfg["discount"], cfg["mid_batch_reset"], cfg["max_path_length"]
    envs = [osampler.SynthAtariEnv(e, pool, rules, cfg["num_img_obs"], 4, cfg["clip_reward"], cfg["episodic_lives"])
            for e in env_ids]
    half = cfg["n_envs"] // 2
    sb = [e - group * half for e in env_ids]              # rows of this worker in its group's step buffer
    traj = [osampler.TrajInfo(discount) for _ in envs]
    for i, env in enumerate(envs):                        # start_envs (sampler/util.py:26-57, no decorrelation)
        step_obs[sb[i]] = env.reset()
    sync["barrier_out"].wait()                            # worker.py:141
    act_waiter, step_blocker = sync["act_waiters"][group][rank], sync["step_blockers"][group][rank]
    while True:
        sync["barrier_in"].wait()
        if sync["quit"].value:
            return
        completed = []
        need = [False] * len(envs)
        step_blocker.release()                            # worker.py:29: the step buffer already holds the first obs
        for i, e in enumerate(env_ids):
            bufs["observations"][e * T] = step_obs[sb[i]]
        for s in range(T):
            act_waiter.acquire()
            for i, (e, env) in enumerate(zip(env_ids, envs)):
                if (not mid_batch_reset) and need[i]:
                    continue
                row = e * T + s
                o, r, d, info = env.step(step_act[sb[i]])
                traj[i].step(float(r), float(info.get("raw_reward", r)))
                over = traj[i]["Length"] > max_path_length
                wrote = True
                if over or (d and info.get("need_reset", True)):
                    d = True
                    if over and "need_reset" in info:
                        info["need_reset"] = True
                    completed.append(dict(traj[i], env=e))
                    traj[i] = osampler.TrajInfo(discount)
                    if mid_batch_reset:
                        o = env.reset()
                    else:
                        need[i] = True
                        wrote = False
                if wrote:
                    step_obs[sb[i]] = o
                    if s < T - 1:
                        bufs["observations"][row + 1] = o
                bufs["rewards"][row] = r
                bufs["dones"][row] = d
                if "raw_reward" in info:
                    bufs["raw_reward"][row] = info["raw_reward"]
                if "need_reset" in info:
                    bufs["need_reset"][row] = info["need_reset"]
            step_blocker.release()
        for t in completed:
            sync["queue"].put(t)
        sync["n_done"][group * cfg["n_parallel"] + rank] = len(completed)   # (mp.Queue hands items over asynchronously)
        sync["barrier_out"].wait()                        # worker.py:147-149
        if not mid_batch_reset:                           # worker.py:150-151: reset only after the batch is handed over
            for i, env in enumerate(envs):
                if need[i]:
                    step_obs[sb[i]] = env.reset()
            sync["barrier_reset"].wait()


class MpOracleSampler(object):
    """2*n_parallel simulator processes (two alternating groups) + this master process"""

    def __init__(self, n_parallel, envs_per, horizon, pool, rules, n_actions=4, discount=0.99, mid_batch_reset=True,
                 max_path_length=27000, num_img_obs=4, clip_reward=True, episodic_lives=True):
        ctx = mp.get_context("fork")
        self.B, self.T, self.A = 2 * n_parallel * envs_per, horizon, n_actions
        self.n_parallel, self.envs_per, self.mid_batch_reset = n_parallel, envs_per, mid_batch_reset
        B, T, P = self.B, horizon, num_img_obs
        N, half = B * T, B // 2
        oh, ow = (84, 84) if pool.ndim == 4 else (104, 80)
        self.buf = dict(observations=_shared((N, P, oh, ow), np.uint8), rewards=_shared((N,), np.float32),
                        dones=_shared((N,), bool), raw_reward=_shared((N,), np.float32), need_reset=_shared((N,), bool),
                        actions=np.zeros(N, np.uint8), prob=np.zeros((N, n_actions), np.float32),
                        value=np.zeros(N, np.float32), extra_observations=np.zeros((B, P, oh, ow), np.uint8))
        self.step_obs = [_shared((half, P, oh, ow), np.uint8) for _ in range(2)]
        self.step_act = [_shared((half,), np.uint8) for _ in range(2)]
        n_workers = 2 * n_parallel
        self.sync = dict(barrier_in=ctx.Barrier(n_workers + 1), barrier_out=ctx.Barrier(n_workers + 1),
                         barrier_reset=ctx.Barrier(n_workers + 1),
                         act_waiters=[[ctx.Semaphore(0) for _ in range(n_parallel)] for _ in range(2)],
                         step_blockers=[[ctx.Semaphore(0) for _ in range(n_parallel)] for _ in range(2)],
                         queue=ctx.Queue(), quit=ctx.RawValue(ctypes.c_bool, False),
                         n_done=ctx.RawArray(ctypes.c_int, n_workers))
        cfg = dict(discount=discount, mid_batch_reset=mid_batch_reset, max_path_length=max_path_length, n_envs=B,
                   n_parallel=n_parallel, num_img_obs=num_img_obs, clip_reward=clip_reward, episodic_lives=episodic_lives)
        shared_bufs = {k: self.buf[k] for k in ("observations", "rewards", "dones", "raw_reward", "need_reset")}
        self.procs = []
        for g in range(2):
            for r in range(n_parallel):
                ids = list(range(g * half + r * envs_per, g * half + (r + 1) * envs_per))   # env-major layout, group halves
                p = ctx.Process(target=_worker, args=(g, r, ids, pool, rules, T, shared_bufs, self.step_obs[g],
                                                      self.step_act[g], self.sync, cfg), daemon=True)
                p.start()
                self.procs.append(p)
        self.sync["barrier_out"].wait()                   # sampler.py:95: all envs started, first obs in the step buffers

    def obtain_samples(self, policy_fn, uniforms):
        """same contract as OracleSampler.obtain_samples"""
        B, T, buf, half = self.B, self.T, self.buf, self.B // 2
        self.sync["barrier_in"].wait()
        for s in range(T):                                # serve_actions (sampler.py:120-151)
            for j in range(2):
                for b in self.sync["step_blockers"][j]:
                    b.acquire()
                prob, value = policy_fn(self.step_obs[j])
                acts = osampler.weighted_sample_n(prob, uniforms[s, j * half:(j + 1) * half], self.A)
                self.step_act[j][:] = acts
                for w in self.sync["act_waiters"][j]:
                    w.release()
                rows = np.arange(j * half, (j + 1) * half) * T + s
                buf["actions"][rows] = acts
                buf["prob"][rows] = prob
                buf["value"][rows] = value
        for j in range(2):
            for b in self.sync["step_blockers"][j]:
                b.acquire()
            buf["extra_observations"][j * half:(j + 1) * half] = self.step_obs[j]
        self.sync["barrier_out"].wait()
        completed = [self.sync["queue"].get() for _ in range(sum(self.sync["n_done"]))]
        if not self.mid_batch_reset:
            self.sync["barrier_reset"].wait()
        return buf, completed

    def shutdown(self):
        self.sync["quit"].value = True
        try:
            self.sync["barrier_in"].wait(timeout=5)
        except Exception:
            pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
