"""Atari environment description (reference: accel_rl/envs/atari_env.py).

The reference steps one ALE emulator per env on CPU worker processes.  Here the environments are
device-resident: the frame pipeline (_update_obs, atari_env.py:151-157), frame-skip / reward /
life / reset logic (:65-100, :165-191) run for ALL envs in env_step_kernel + frame_kernel
(csrc/kernels.cuh), driven by the synthetic emulator rules (the real emulator is out of scope,
SURVEY.md §2a).  This class therefore only carries the constructor arguments, spaces and the
synthetic-emulator rule set; `AtariEnv.device_resident` tells the sampler not to look for step().
"""
import numpy as np

from accel_rl_b200.spaces.discrete import Discrete
from accel_rl_b200.spaces.uintbox import UintBox

W, H = (80, 104)  # atari_env.py:13

# Minimal action sets (ALE getMinimalActionSet) for the games the BASELINE configs name
MINIMAL_ACTIONS = {"breakout": 4, "pong": 6, "space_invaders": 6, "seaquest": 18, "qbert": 6, "beam_rider": 9}

# Game mixes (BASELINE configs[2] "A2C 4-game Atari mix"): env e plays games[e % len(games)]; one policy serves them all,
# its action count padded to the largest minimal action set of the mix (SURVEY.md §8(d) synthetic-inputs row)
GAME_MIXES = {"mix4": ("breakout", "pong", "space_invaders", "beam_rider")}

DEFAULT_SYNTH_RULES = dict(pool_frames=1024, lives0=5, life_base=400, life_mul=31, life_mod=257, reward_mod=389,
                           frame_stride=263, pool_seed=0)


class EnvSpec(object):
    def __init__(self, observation_space, action_space):
        self._observation_space = observation_space
        self._action_space = action_space

    observation_space = property(lambda self: self._observation_space)
    action_space = property(lambda self: self._action_space)


class AtariEnv(object):
    device_resident = True

    def __init__(self, game="pong", frame_skip=4, num_img_obs=4, clip_reward=True, episodic_lives=True,
                 max_start_noops=30, repeat_action_probability=0., synth_rules=None, n_actions=None, frame_mode="gray"):
        # frame_mode "gray": the reference pipeline (grayscale 210x160 screens -> 104x80, atari_env.py:147-157);
        # "rgb": the north-star pipeline (RGB 210x160x3 screens -> gray -> 84x84), builder-defined (oracle/frame.py:rgb_*)
        if frame_mode not in ("gray", "rgb"):
            raise ValueError("frame_mode must be 'gray' or 'rgb'")
        self.frame_mode = frame_mode
        if frame_skip != 4:
            raise NotImplementedError("device env implements frame_skip=4 (atari_env.py:19 default)")
        if num_img_obs not in (1, 4):
            raise NotImplementedError("num_img_obs must be 1 or 4")
        self._game = game
        self._frame_skip = frame_skip
        self._num_img_obs = num_img_obs
        self._clip_reward = clip_reward
        self._episodic_lives = episodic_lives
        self._max_start_noops = max_start_noops
        self._repeat_action_probability = repeat_action_probability
        self.synth_rules = dict(DEFAULT_SYNTH_RULES)
        if synth_rules:
            self.synth_rules.update(synth_rules)
        if game in GAME_MIXES:
            games = GAME_MIXES[game]
            self.synth_rules["n_games"] = len(games)
            if self.synth_rules["pool_frames"] % len(games):
                raise ValueError("pool_frames must be a multiple of the number of games in the mix")
            default_n = max(MINIMAL_ACTIONS[g] for g in games)
        else:
            self.synth_rules.setdefault("n_games", 1)
            default_n = MINIMAL_ACTIONS.get(game, 4)
        n = n_actions if n_actions is not None else default_n
        self._action_space = Discrete(n)
        obs_hw = (H, W) if frame_mode == "gray" else (84, 84)
        self._observation_space = UintBox(shape=(num_img_obs,) + obs_hw, bits=8)
        # the reference ctor ends with self.reset(), whose start no-ops draw from the global stream
        # (atari_env.py:63,97); keep the draw so master-process RNG consumption stays identical
        self._draw_start_noops()

    def _draw_start_noops(self):
        return int(np.random.randint(0, self._max_start_noops + 1))

    def reset(self):
        """Host-visible reset only consumes the start-noop draw; observations live on the device."""
        self._draw_start_noops()
        return np.zeros(self._observation_space.shape, np.uint8)

    action_space = property(lambda self: self._action_space)
    observation_space = property(lambda self: self._observation_space)
    spec = property(lambda self: EnvSpec(self._observation_space, self._action_space))
    game = property(lambda self: self._game)
    frame_skip = property(lambda self: self._frame_skip)
    num_img_obs = property(lambda self: self._num_img_obs)
    clip_reward = property(lambda self: self._clip_reward)
    episodic_lives = property(lambda self: self._episodic_lives)
    max_start_noops = property(lambda self: self._max_start_noops)
    repeat_action_probability = property(lambda self: self._repeat_action_probability)


def make_frame_pool(pool_frames, seed=0, channels=1):
    """Synthetic emulator frames (pool_frames, 210, 160[, 3]) uint8 — same generator as the oracle."""
    shape = (pool_frames, 210, 160) if channels == 1 else (pool_frames, 210, 160, channels)
    return np.random.RandomState(seed).randint(0, 256, shape, dtype=np.uint8)
