"""Host side of an Atari env whose pixels are processed on the GPU.

Mirrors the emulator-control half of the reference's AtariEnv (accel_rl/envs/atari_env.py:65-100 step / reset,
:165-191 _check_life / _life_reset / _done_*): frame skip with the reward summed over the repeats, the screen grabbed
after the 3rd and the 4th repeat, reward clipping with info["raw_reward"], episodic-life handling, start no-ops drawn
from the worker's numpy stream.  What it does NOT do is touch observations: instead of max / resize / stack
(:151-157, the GPU's fused frame kernel) it writes the two raw screens into the caller's buffers and returns a flag
saying how the device must treat the stack:

    FLAG_RESET   the stack is zeroed and frame 1 counts as zeros (`_reset_obs` then `_update_obs`, :159-163): after
                 reset() and after a life loss under episodic_lives

The emulator is any object with the ALE calls the reference makes: act, lives, game_over, reset_game,
getMinimalActionSet, getScreenGrayscale(buf) (or getScreenRGB(buf) for the RGB pipeline).  numpy only: this module is
imported by the emulator worker processes.
"""
import numpy as np

FLAG_RESET = 1       # arl_ext_step.flags (include/accelrl_b200.h)
FLAG_SKIP = 2        # observation not advanced
FLAG_NO_RECORD = 4   # env not stepped: nothing recorded

# ALE action ids -> names (reference: envs/atari_env.py:194-213)
ACTION_MEANING = {
    0: "NOOP", 1: "FIRE", 2: "UP", 3: "RIGHT", 4: "LEFT", 5: "DOWN", 6: "UPRIGHT", 7: "UPLEFT", 8: "DOWNRIGHT",
    9: "DOWNLEFT", 10: "UPFIRE", 11: "RIGHTFIRE", 12: "LEFTFIRE", 13: "DOWNFIRE", 14: "UPRIGHTFIRE", 15: "UPLEFTFIRE",
    16: "DOWNRIGHTFIRE", 17: "DOWNLEFTFIRE",
}


class HostAtariEnv(object):
    def __init__(self, emulator, frame_skip=4, clip_reward=True, episodic_lives=True, max_start_noops=30, rgb=False):
        self.ale = emulator
        self._action_set = list(emulator.getMinimalActionSet())
        meanings = [ACTION_MEANING[int(i)] for i in self._action_set]
        self._has_fire = "FIRE" in meanings
        self._has_up = "UP" in meanings
        self._frame_skip = frame_skip
        self._clip_reward = clip_reward
        self._episodic_lives = episodic_lives
        self._max_start_noops = max_start_noops
        self._rgb = rgb
        self._lives = 0

    @property
    def n_actions(self):
        return len(self._action_set)

    def _screen(self, out):
        if self._rgb:
            self.ale.getScreenRGB(out)
        else:
            self.ale.getScreenGrayscale(out)

    def _life_reset(self):                     # atari_env.py:172-179
        self.ale.act(0)
        if self._has_fire:
            self.ale.act(1)
        if self._has_up:
            self.ale.act(2)
        self._lives = self.ale.lives()

    def _check_life(self):                     # atari_env.py:165-170
        lives = self.ale.lives()
        lost_life = (lives < self._lives) and (lives > 0)
        if lost_life:
            self._life_reset()
        return lost_life

    def reset(self, frame2):
        """atari_env.py:93-100; writes the screen `_update_obs` would read; -> flags"""
        self.ale.reset_game()
        self._life_reset()
        for _ in range(np.random.randint(0, self._max_start_noops + 1)):
            self.ale.act(0)
        self._screen(frame2)
        return FLAG_RESET

    def step(self, action, frame1, frame2):
        """atari_env.py:65-78 -> (reward f32, raw_reward f32, done, need_reset or None, flags)"""
        a = self._action_set[action]
        reward = np.float32(0.)
        for _ in range(self._frame_skip - 1):
            reward += np.float32(self.ale.act(a))
        self._screen(frame1)
        reward += np.float32(self.ale.act(a))
        self._screen(frame2)
        raw = reward
        if self._clip_reward:
            reward = np.sign(reward)
        flags = 0
        if self._episodic_lives:               # _done_episodic_lives, :185-191
            need_reset = bool(self.ale.game_over())
            lost_life = self._check_life()
            if lost_life:
                self._screen(frame2)           # _reset_obs + _update_obs on the screen after the life reset
                flags = FLAG_RESET
            done = lost_life or need_reset
        else:                                  # _done_no_epidosic_lives, :181-183
            self._check_life()
            need_reset = None
            done = bool(self.ale.game_over())
        return np.float32(reward), np.float32(raw), done, need_reset, flags


def make_ale(game="pong", repeat_action_probability=0.):
    """Emulator factory for real Atari: atari_py (the reference's dependency) or ale_py, whichever is installed."""
    try:
        import atari_py
        ale = atari_py.ALEInterface()
        ale.setFloat(b"repeat_action_probability", repeat_action_probability)
        ale.loadROM(atari_py.get_game_path(game))
        return ale
    except ImportError:
        pass
    try:
        import ale_py
    except ImportError:
        raise ImportError("no Atari emulator installed (atari_py / ale_py): pass an emu_factory of your own")
    ale = ale_py.ALEInterface()
    ale.setFloat("repeat_action_probability", repeat_action_probability)
    import ale_py.roms as roms
    ale.loadROM(getattr(roms, rom_attr(game)))
    return ale


def rom_attr(game):
    """ale_py names its ROMs in CamelCase: 'space_invaders' -> 'SpaceInvaders' (atari_py uses the snake_case name)"""
    return "".join(part.capitalize() for part in str(game).split("_"))
