"""ORACLE (test infrastructure).  CPU restatement of the reference rollout path.

  env        accel_rl/envs/atari_env.py:65-78 (step), :93-100 (reset), :151-163 (_update_obs/_reset_obs),
             :165-191 (_check_life/_life_reset/_done_*), over oracle.synth_ale rules instead of ALE
  collector  accel_rl/sampler/act_server/alternating/overlap/worker.py:25-60 (ResetCollector),
             :63-113 (NonResetCollector), sampler/util.py:26-57 (start_envs, no decorrelation),
             :75-101 (TrajInfo)
  master     .../overlap/sampler.py:120-151 (serve_actions: two groups, group j = j-th half of the envs;
             per step: for j in (0,1): policy.get_actions(step_bufs[j].obs)), buffers.py:7-38,
             buffers/batch.py:36-76 (row = env*T + t)
  sampling   rllab/misc/special.py:22-27 (weighted_sample_n)

Single process; identical buffers to the real multi-process sampler (checked against it in
tests/test_oracle_vs_reference.py and through tests/golden/sampler_*.npz).
"""
import numpy as np

from oracle import frame as oframe
from oracle import synth_ale as sa


def weighted_sample_n(prob_matrix, r, n_items):
    """special.py:22-27 with the uniforms r supplied; items = arange(n) as uint8 (discrete.py:14-19)."""
    s = prob_matrix.cumsum(axis=1)                       # float32 sequential cumsum
    k = (s < r.reshape((-1, 1))).sum(axis=1)             # f32 < f64 compare
    return np.minimum(k, n_items - 1).astype(np.uint8)


class TrajInfo(dict):
    def __init__(self, discount=1.0):
        super().__init__(Length=0, Return=0.0, RawReturn=0.0, NonzeroRewards=0, DiscountedReturn=0.0)
        self._discount = discount
        self._cur = 1.0

    def step(self, r, raw):
        self["Length"] += 1
        self["Return"] += r
        self["RawReturn"] += raw
        self["NonzeroRewards"] += int(r != 0)
        self["DiscountedReturn"] += self._cur * r
        self._cur *= self._discount


class SynthAtariEnv(object):
    """AtariEnv restated over the synthetic emulator (action set NOOP/FIRE/RIGHT/LEFT: has_fire, no up)."""

    def __init__(self, env_id, pool, rules, num_img_obs=4, frame_skip=4, clip_reward=True, episodic_lives=True):
        self.e, self.pool, self.rules = env_id, pool, rules
        self.P, self.frame_skip = num_img_obs, frame_skip
        self.clip_reward, self.episodic_lives = clip_reward, episodic_lives
        self.f = 0
        # north-star mode (RGB pool, trailing channel axis): same env logic, frames go through oracle.frame.rgb_*
        self.rgb = pool.ndim == 4
        self.raw_shape = pool.shape[1:]
        self.hw = (oframe.NS_H, oframe.NS_W) if self.rgb else (oframe.H, oframe.W)
        self.obs = np.zeros((num_img_obs,) + self.hw, np.uint8)
        self.raw1 = np.zeros(self.raw_shape, np.uint8)
        self.raw2 = np.zeros(self.raw_shape, np.uint8)
        self.lives_seen = 0

    # --- emulator ---
    def _act(self):
        self.f += 1
        return sa.synth_reward(self.rules, self.e, self.f)

    def _screen(self):
        return self.pool[sa.frame_index(self.rules, self.e, self.f)]

    def _lives(self):
        return sa.synth_lives(self.rules, self.e, self.f)

    # --- AtariEnv ---
    def _update_obs(self):
        self.raw2 = self._screen()
        if self.rgb:
            img = oframe.rgb_downsample(oframe.rgb_to_gray(np.maximum(self.raw1, self.raw2)))
            self.obs = np.concatenate([self.obs[1:], img[np.newaxis]])
        else:
            self.obs = oframe.update_obs(self.obs, self.raw1, self.raw2)

    def _reset_obs(self):
        self.obs = np.zeros_like(self.obs)
        self.raw1 = np.zeros(self.raw_shape, np.uint8)
        self.raw2 = np.zeros(self.raw_shape, np.uint8)

    def _life_reset(self):
        self._act()          # act(0)
        self._act()          # act(1): FIRE is in the action set
        self.lives_seen = self._lives()

    def reset(self):
        self.f = 0           # reset_game
        self._reset_obs()
        self._life_reset()
        # max_start_noops = 0: randint(0, 1) == 0 no-ops
        self._update_obs()
        return self.obs.copy()

    def step(self, action):
        reward = np.float32(0.0)
        for _ in range(self.frame_skip - 1):
            reward += np.float32(self._act())
        self.raw1 = self._screen()
        reward += np.float32(self._act())
        self._update_obs()
        info = dict()
        if self.clip_reward:
            info["raw_reward"] = reward
            reward = np.sign(reward)
        game_over = self._lives() == 0
        lives = self._lives()
        lost_life = (lives < self.lives_seen) and (lives > 0)
        if self.episodic_lives:
            info["need_reset"] = game_over
            if lost_life:
                self._life_reset()
                self._reset_obs()
                self._update_obs()
            done = lost_life or game_over
        else:
            if lost_life:
                self._life_reset()
            done = game_over
        return self.obs.copy(), reward, done, info


class OracleSampler(object):
    def __init__(self, n_envs, horizon, pool, rules, n_actions=4, discount=0.99, mid_batch_reset=True,
                 max_path_length=27000, num_img_obs=4, clip_reward=True, episodic_lives=True):
        assert n_envs % 2 == 0
        self.B, self.T, self.A = n_envs, horizon, n_actions
        self.discount, self.mid_batch_reset, self.max_path_length = discount, mid_batch_reset, max_path_length
        self.envs = [SynthAtariEnv(e, pool, rules, num_img_obs, 4, clip_reward, episodic_lives) for e in range(n_envs)]
        N, P = n_envs * horizon, num_img_obs
        oh, ow = self.envs[0].hw

        class _HW(object):
            H, W = oh, ow
        oframe_hw = _HW
        self.buf = dict(
            observations=np.zeros((N, P, oframe_hw.H, oframe_hw.W), np.uint8),
            rewards=np.zeros(N, np.float32), dones=np.zeros(N, bool),
            raw_reward=np.zeros(N, np.float32), need_reset=np.zeros(N, bool),
            actions=np.zeros(N, np.uint8), prob=np.zeros((N, n_actions), np.float32), value=np.zeros(N, np.float32),
            extra_observations=np.zeros((n_envs, P, oh, ow), np.uint8))
        self.step_obs = np.zeros((n_envs, P, oh, ow), np.uint8)
        self.traj = [TrajInfo(discount) for _ in range(n_envs)]
        self.need = [False] * n_envs
        for e, env in enumerate(self.envs):            # start_envs
            self.step_obs[e] = env.reset()

    def decorrelate(self, n_steps):
        """start_envs with max_decorrelation_steps > 0 (sampler/util.py:33-55) for GIVEN per-env warm-up step counts (the
        reference draws them from the wall clock): env e takes n_steps[e] steps, is reset at once whenever its trajectory
        ends (need_reset / over-length), and its running TrajInfo carries into the first rollout.  The synthetic
        emulator ignores actions, so the reference's random actions need no counterpart."""
        for e, env in enumerate(self.envs):
            traj = TrajInfo(self.discount)
            o = self.step_obs[e]
            for _ in range(int(n_steps[e])):
                o, r, d, info = env.step(0)
                traj.step(float(r), float(info.get("raw_reward", r)))
                if traj["Length"] > self.max_path_length or (d and info.get("need_reset", True)):
                    o = env.reset()
                    traj = TrajInfo(self.discount)
            self.step_obs[e] = o
            self.traj[e] = traj

    def obtain_samples(self, policy_fn, uniforms):
        """policy_fn(obs (n,P,104,80) u8) -> (prob (n,A) f32, value (n,) f32); uniforms (T, B) float64 in the
        master's consumption order (step-major, group 0 half then group 1 half)."""
        B, T, buf = self.B, self.T, self.buf
        half = B // 2
        completed = []
        need = [False] * B
        for e in range(B):
            buf["observations"][e * T] = self.step_obs[e]
        for s in range(T):
            for j in range(2):
                lo, hi = j * half, (j + 1) * half
                prob, value = policy_fn(self.step_obs[lo:hi])
                acts = weighted_sample_n(prob, uniforms[s, lo:hi], self.A)
                for e in range(lo, hi):
                    row = e * T + s
                    buf["actions"][row] = acts[e - lo]
                    buf["prob"][row] = prob[e - lo]
                    buf["value"][row] = value[e - lo]
                for e in range(lo, hi):
                    if (not self.mid_batch_reset) and need[e]:
                        continue
                    env, row = self.envs[e], e * T + s
                    o, r, d, info = env.step(acts[e - lo])
                    self.traj[e].step(float(r), float(info.get("raw_reward", r)))
                    over = self.traj[e]["Length"] > self.max_path_length
                    wrote_obs = True
                    if over or (d and info.get("need_reset", True)):
                        d = True
                        if over and "need_reset" in info:
                            info["need_reset"] = True
                        completed.append(dict(self.traj[e], env=e))
                        self.traj[e] = TrajInfo(self.discount)
                        if self.mid_batch_reset:
                            o = env.reset()
                        else:
                            need[e] = True
                            wrote_obs = False
                    if wrote_obs:
                        self.step_obs[e] = o
                        if s < T - 1:
                            buf["observations"][row + 1] = o
                    buf["rewards"][row] = r
                    buf["dones"][row] = d
                    if "raw_reward" in info:
                        buf["raw_reward"][row] = info["raw_reward"]
                    if "need_reset" in info:
                        buf["need_reset"][row] = info["need_reset"]
        buf["extra_observations"][:] = self.step_obs
        if not self.mid_batch_reset:
            for e in range(B):
                if need[e]:
                    self.step_obs[e] = self.envs[e].reset()
        return buf, completed
