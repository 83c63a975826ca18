"""Emulator worker process of the host-fed sampler.

The reference's simulator worker (sampler/act_server/alternating/overlap/worker.py:23-153) with the observation
handling removed: per step it waits for its envs' actions, steps the emulators (HostAtariEnv), and writes — into memory
shared with the master and page-locked for DMA — the two raw screens and one 12-byte record per env
(reward, raw_reward, done, need_reset, flags); the GPU does everything that touches pixels.  ResetCollector /
NonResetCollector semantics (worker.py:25-113): an env is reset when the game is over or the trajectory is longer than
max_path_length; with mid_batch_reset == False a finished env is not stepped again in the batch (FLAG_SKIP) and is
reset on CMD_RESET_NEEDED after it.  Completed TrajInfos (sampler/util.py:75-101) go to a queue.
numpy + stdlib only.
"""
import numpy as np

from accel_rl_b200.hostsim.atari_env import HostAtariEnv, FLAG_RESET, FLAG_SKIP, FLAG_NO_RECORD

CMD_STEP, CMD_RESET_NEEDED, CMD_QUIT, CMD_WARM = 0, 1, 2, 3

EXT_DTYPE = np.dtype([("reward", np.float32), ("raw_reward", np.float32), ("done", np.uint8), ("need_reset", np.uint8),
                      ("flags", np.uint8), ("pad", np.uint8)])      # arl_ext_step


def _new_traj():
    return dict(Length=0, Return=0., RawReturn=0., NonzeroRewards=0, DiscountedReturn=0., _cur=1.)


def views(shared, n_envs, frame_shape):
    """numpy views of the shared blocks: frames [B][2][frame], ext [B] records, act [B]"""
    frames = np.frombuffer(shared["frames"], dtype=np.uint8).reshape((n_envs, 2) + tuple(frame_shape))
    ext = np.frombuffer(shared["ext"], dtype=EXT_DTYPE)
    act = np.frombuffer(shared["act"], dtype=np.uint8)
    return frames, ext, act


class Collector(object):
    """The envs of one worker and the collector rules applied to them (worker.py:25-113): `start` = start_envs,
    `step` = one time step of collect(), `reset_needed` = reset_needed_envs.  Every call leaves, for each env e it
    owns, the record ext[e] and — unless the record says the observation does not advance — the raw screens
    frames[e, 0] / frames[e, 1]; completed TrajInfos are handed to `emit`."""

    def __init__(self, env_lo, env_hi, emu_factory, env_kwargs, frames, ext, act, emit, mid_batch_reset,
                 max_path_length, discount):
        self.env_lo, self.frames, self.ext, self.act, self.emit = env_lo, frames, ext, act, emit
        self.mid_batch_reset, self.max_path_length, self.discount = mid_batch_reset, max_path_length, discount
        self.envs = [HostAtariEnv(emu_factory(e), **env_kwargs) for e in range(env_lo, env_hi)]
        self.trajs = [_new_traj() for _ in self.envs]
        self.need = [False] * len(self.envs)

    def _reset(self, i):
        e = self.env_lo + i
        self.ext[e] = (0., 0., 0, 0, self.envs[i].reset(self.frames[e, 1]), 0)

    def start(self, max_decorrelation_steps=0):
        """start_envs (sampler/util.py:26-57): reset every env; with max_decorrelation_steps > 0 also draw, per env, how
        many random-action warm-up steps it takes (the reference derives the fraction from the wall clock,
        sampler/util.py:22-23; here it comes from the worker's seeded numpy stream) -> the largest count"""
        for i in range(len(self.envs)):
            self._reset(i)
        self.warm_n = [int(np.random.rand() * max_decorrelation_steps) if max_decorrelation_steps > 0 else 0
                       for _ in self.envs]
        return max(self.warm_n) if self.warm_n else 0

    def warm_step(self, k):
        """warm-up step k of start_envs (sampler/util.py:44-53): envs with warm_n > k take a uniformly random action
        (action_space.sample), are reset at once when their trajectory ends; nothing is recorded or reported"""
        for i, env in enumerate(self.envs):
            e = self.env_lo + i
            if k >= self.warm_n[i]:
                self.ext[e] = (0., 0., 0, 0, FLAG_SKIP | FLAG_NO_RECORD, 0)
                continue
            r, raw, d, nr, fl = env.step(int(np.random.randint(env.n_actions)), self.frames[e, 0], self.frames[e, 1])
            t = self.trajs[i]
            t["Length"] += 1
            if t["Length"] > self.max_path_length or (d and (True if nr is None else nr)):
                self.trajs[i] = _new_traj()
                fl = env.reset(self.frames[e, 1])
            self.ext[e] = (0., 0., 0, 0, fl | FLAG_NO_RECORD, 0)

    def reset_needed(self):                                # worker.py:106-113
        for i in range(len(self.envs)):
            if self.need[i]:
                self._reset(i)
                self.need[i] = False
            else:
                self.ext[self.env_lo + i] = (0., 0., 0, 0, FLAG_SKIP | FLAG_NO_RECORD, 0)

    def step(self):
        for i, env in enumerate(self.envs):
            e = self.env_lo + i
            if self.need[i]:                               # finished earlier in this batch: not stepped (worker.py:78)
                self.ext[e] = (0., 0., 0, 0, FLAG_SKIP | FLAG_NO_RECORD, 0)
                continue
            r, raw, d, nr, fl = env.step(int(self.act[e]), self.frames[e, 0], self.frames[e, 1])
            t = self.trajs[i]
            t["Length"] += 1
            t["Return"] += float(r)
            t["RawReturn"] += float(raw)
            t["NonzeroRewards"] += int(r != 0)
            t["DiscountedReturn"] += t["_cur"] * float(r)
            t["_cur"] *= self.discount
            over_length = t["Length"] > self.max_path_length
            if over_length or (d and (True if nr is None else nr)):
                d = True
                if over_length and nr is not None:
                    nr = True
                self.emit((e, t["Length"], t["Return"], t["RawReturn"], t["NonzeroRewards"], t["DiscountedReturn"]))
                self.trajs[i] = _new_traj()
                if self.mid_batch_reset:
                    fl = env.reset(self.frames[e, 1])
                else:
                    self.need[i] = True
                    fl = FLAG_SKIP                          # this step's reward/done ARE recorded, the obs is not advanced
            self.ext[e] = (r, raw, int(bool(d)), int(bool(nr)), fl, 0)


def worker_main(rank, env_lo, env_hi, n_envs, emu_factory, env_kwargs, frame_shape, shared, cmd, act_ready, step_done,
                infos_queue, seed, mid_batch_reset, max_path_length, discount, max_decorrelation_steps=0):
    np.random.seed(seed)                                   # initialize_worker: seed + rank (sampler/util.py:60-72)
    frames, ext, act = views(shared, n_envs, frame_shape)
    col = Collector(env_lo, env_hi, emu_factory, env_kwargs, frames, ext, act, infos_queue.put, mid_batch_reset,
                    max_path_length, discount)
    # what the master checks before the first step: the emulators' action count and this worker's warm-up length
    report = np.frombuffer(shared["report"], dtype=np.int32).reshape(-1, 2)
    report[rank, 0] = col.envs[0].n_actions if all(e.n_actions == col.envs[0].n_actions for e in col.envs) else -1
    report[rank, 1] = col.start(max_decorrelation_steps)   # start_envs
    step_done.release()
    while True:
        act_ready.acquire()
        c = cmd.value
        if c == CMD_QUIT:
            break
        if c == CMD_RESET_NEEDED:
            col.reset_needed()
        elif c == CMD_WARM:
            col.warm_step(int(shared["arg"][0]))
        else:
            col.step()
        step_done.release()


def profiling_worker(profile_pathname, rank, *args):
    """worker_main under cProfile, dumped to <profile_pathname>_sim_<rank>.prof when the worker quits
    (reference: sampler/util.py:10-19 profiling_process, enabled by the sampler's profile_pathname argument)"""
    import cProfile
    prof = cProfile.Profile()
    try:
        prof.runcall(worker_main, rank, *args)
    finally:
        prof.dump_stats("{}_sim_{}.prof".format(profile_pathname, rank))
