"""Host mirror of the device's synthetic emulator (csrc/kernels.cuh synth_*: SynthCfg, synth_hash, synth_reward,
synth_lives, synth_frame_index) behind the ALE calls HostAtariEnv makes.  Lets the host-emulator sampler run — and be
benchmarked — on machines without an Atari emulator: the same deterministic game rules the device-resident sampler
uses, but stepped by worker processes on host cores and shipped to the GPU as raw screens.

    screen(e, f) = pool[(e + frame_stride * f) % pool_frames]
    reward(e, f) in {0, 1, 4, -1} from a 32-bit hash of (e, f), paid by the act() that enters frame f
    lives(e, f)  = max(0, lives0 - f // (life_base + (e * life_mul) % life_mod));  game over at 0 lives
numpy only (imported by the worker processes).
"""
import numpy as np

DEFAULT_RULES = dict(pool_frames=1024, lives0=5, life_base=400, life_mul=31, life_mod=257, reward_mod=389,
                     frame_stride=263)
_M32 = 0xFFFFFFFF
_POOLS = {}


def frame_pool(pool_frames, seed=0, channels=1):
    key = (pool_frames, seed, channels)
    if key not in _POOLS:
        shape = (pool_frames, 210, 160) if channels == 1 else (pool_frames, 210, 160, channels)
        _POOLS[key] = np.random.RandomState(seed).randint(0, 256, shape, dtype=np.uint8)
    return _POOLS[key]


class SynthEmulator(object):
    def __init__(self, env_id, rules=None, pool_seed=0, channels=1):
        r = dict(DEFAULT_RULES)
        if rules:
            r.update(rules)
        self.e, self.f = int(env_id), 0
        self.pool = frame_pool(r["pool_frames"], pool_seed, channels)
        self.pool_frames, self.stride = r["pool_frames"], r["frame_stride"]
        # game mix: env e plays game e % n_games (own pool slice, reward table, life clock), as csrc/kernels.cuh synth_game
        n_games = int(r.get("n_games", 1))
        self.game = self.e % n_games if n_games > 1 else 0
        self.fpg = self.pool_frames // n_games if n_games > 1 else self.pool_frames
        self.reward_mod, self.lives0 = r["reward_mod"] + 6 * self.game, r["lives0"]
        self.period = r["life_base"] + 17 * self.game + (self.e * r["life_mul"]) % r["life_mod"]

    def getMinimalActionSet(self):
        return np.array([0, 1, 3, 4], dtype=np.int32)          # NOOP FIRE RIGHT LEFT (Breakout)

    def reset_game(self):
        self.f = 0

    def act(self, a):
        self.f += 1
        h = (self.e * 0x9E3779B1 + self.f * 0x85EBCA77 + 0x165667B1) & _M32
        h ^= h >> 15
        h = (h * 0x2C1B3C6D) & _M32
        h ^= h >> 12
        h = (h * 0x297A2D39) & _M32
        h ^= h >> 15
        if h % self.reward_mod:
            return 0.0
        k = (h // self.reward_mod) & 3
        return 4.0 if k == 2 else (-1.0 if k == 3 else 1.0)

    def lives(self):
        return max(0, self.lives0 - self.f // self.period)

    def game_over(self):
        return self.lives() == 0

    def getScreenGrayscale(self, buf):
        buf[...] = self.pool[self.game * self.fpg + (self.e + self.stride * self.f) % self.fpg].reshape(buf.shape)
        return buf

    getScreenRGB = getScreenGrayscale


def make(env_index, rules=None, pool_seed=0, channels=1):
    """picklable emulator factory for HostEmulatorSampler(emu_factory=functools.partial(make, rules=...))"""
    return SynthEmulator(env_index, rules, pool_seed, channels)
