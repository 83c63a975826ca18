"""Test emulator for the host-fed sampler: the oracle's synthetic ALE rules (oracle/synth_ale.py) behind the ALE calls
HostAtariEnv makes, constructible inside a spawned worker process from picklable arguments."""
import numpy as np

from oracle import synth_ale as sa

_POOLS = {}


class FakeALE(object):
    def __init__(self, env_id, rules, pool):
        self.e, self.rules, self.pool, self.f = env_id, rules, pool, 0

    def getMinimalActionSet(self):
        return np.array([0, 1, 3, 4], dtype=np.int32)      # NOOP FIRE RIGHT LEFT (Breakout)

    def reset_game(self):
        self.f = 0

    def act(self, a):
        self.f += 1
        return sa.synth_reward(self.rules, self.e, self.f)

    def lives(self):
        return sa.synth_lives(self.rules, self.e, self.f)

    def game_over(self):
        return self.lives() == 0

    def getScreenGrayscale(self, buf):
        buf[:] = self.pool[sa.frame_index(self.rules, self.e, self.f)].reshape(buf.shape)
        return buf

    getScreenRGB = getScreenGrayscale


def make(env_index, rules, pool_seed=0, channels=1):
    key = (rules["pool_frames"], pool_seed, channels)
    if key not in _POOLS:
        _POOLS[key] = sa.make_pool(rules["pool_frames"], pool_seed, channels)
    return FakeALE(env_index, rules, _POOLS[key])
