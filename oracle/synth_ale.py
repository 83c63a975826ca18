"""ORACLE (test infrastructure).  Synthetic emulator ("SynthALE") — CPU definition.

Stands in for atari_py.ALEInterface (absent; the emulator is out of scope).  It
offers exactly the calls accel_rl/envs/atari_env.py makes (:31-46, :69-71, :94-98, :149,
:166-179): getMinimalActionSet, getScreenGrayscale, act, lives, game_over, reset_game, setFloat,
loadROM.  All behaviour is a deterministic function of (env id, emulator frame counter):

    screen(e, f)  = pool[(e + frame_stride * f) % pool_frames]
    reward(e, f)  = {0 | 1 | 4 | -1} from a 32-bit hash of (e, f)        (paid by act() entering frame f)
    lives(e, f)   = max(0, lives0 - f // (life_base + (e * life_mul) % life_mod))
    game_over     = lives == 0

accel_rl_b200/csrc/kernels.cuh (env_step_kernel, synth_*) implements the same rules on the device.
"""
import numpy as np

DEFAULT_RULES = dict(pool_frames=1024, lives0=5, life_base=400, life_mul=31, life_mod=257, reward_mod=389,
                     frame_stride=263)

M32 = 0xFFFFFFFF


def synth_hash(e, f):
    h = (e * 0x9E3779B1 + f * 0x85EBCA77 + 0x165667B1) & M32
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & M32
    h ^= h >> 12
    h = (h * 0x297A2D39) & M32
    h ^= h >> 15
    return h


def game_of(rules, e):
    """game mix (BASELINE configs[2] "4-game mix"): env e plays game e % n_games; every game has its own slice of the
    frame pool, its own reward table (reward_mod + 6 g) and life clock (life_base + 17 g)"""
    n = rules.get("n_games", 1)
    return e % n if n > 1 else 0


def synth_reward(rules, e, f):
    h = synth_hash(e, f)
    rm = rules["reward_mod"] + 6 * game_of(rules, e)
    if h % rm != 0:
        return 0.0
    k = (h // rm) & 3
    return 4.0 if k == 2 else (-1.0 if k == 3 else 1.0)


def life_period(rules, e):
    return rules["life_base"] + 17 * game_of(rules, e) + (e * rules["life_mul"]) % rules["life_mod"]


def synth_lives(rules, e, f):
    return max(0, rules["lives0"] - f // life_period(rules, e))


def frame_index(rules, e, f):
    n = rules.get("n_games", 1)
    if n > 1:
        fpg = rules["pool_frames"] // n
        return game_of(rules, e) * fpg + (e + rules["frame_stride"] * f) % fpg
    return (e + rules["frame_stride"] * f) % rules["pool_frames"]


def make_pool(pool_frames, seed=0, channels=1):
    """Synthetic grayscale frame pool (pool_frames, 210, 160) uint8 (channels=3: RGB, trailing axis)."""
    rng = np.random.RandomState(seed)
    shape = (pool_frames, 210, 160) if channels == 1 else (pool_frames, 210, 160, channels)
    return rng.randint(0, 256, shape, dtype=np.uint8)


class SynthALE(object):
    """Drop-in for atari_py.ALEInterface inside the reference's AtariEnv."""

    # set by the harness before envs are constructed
    rules = dict(DEFAULT_RULES)
    pool = None
    next_env_id = 0

    def __init__(self):
        self.env_id = SynthALE.next_env_id
        SynthALE.next_env_id += 1
        self.f = 0

    # --- configuration calls (ignored) ---
    def setFloat(self, key, value):
        pass

    def setInt(self, key, value):
        pass

    def loadROM(self, path):
        pass

    def getMinimalActionSet(self):
        return np.array([0, 1, 3, 4], dtype=np.int32)  # NOOP FIRE RIGHT LEFT (Breakout)

    # --- emulation ---
    def reset_game(self):
        self.f = 0

    def act(self, a):
        self.f += 1
        return synth_reward(self.rules, self.env_id, self.f)

    def lives(self):
        return synth_lives(self.rules, self.env_id, self.f)

    def game_over(self):
        return self.lives() == 0

    def getScreenGrayscale(self, buf=None):
        frame = self.pool[frame_index(self.rules, self.env_id, self.f)]
        if buf is None:
            return frame.reshape(210, 160, 1).copy()
        buf[:] = frame.reshape(buf.shape)
        return buf
