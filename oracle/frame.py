"""ORACLE (test infrastructure).  Atari frame pipeline restated in numpy.

Follows accel_rl/envs/atari_env.py:151-157 (_update_obs), :159-163 (_reset_obs):
    max of the two raw grayscale frames -> drop rows 208,209 -> cv2.resize(.., (80,104), dst)
    [the reference passes cv2.INTER_NEAREST positionally into `dst`, so the interpolation is the
    default INTER_LINEAR, which at an exact 2x shrink is the 2x2 box mean (a+b+c+d+2)>>2 — checked
    bit-equal against cv2 in tests/test_oracle_vs_reference.py] -> obs = concat(obs[1:], img).
"""
import numpy as np

H, W = 104, 80


def downsample(max_frame):
    """(210,160) u8 -> (104,80) u8, 2x2 box mean with round-half-up."""
    m = max_frame[:208].astype(np.uint16)
    s = m[0::2, 0::2] + m[0::2, 1::2] + m[1::2, 0::2] + m[1::2, 1::2] + 2
    return (s >> 2).astype(np.uint8)


def update_obs(obs, raw1, raw2):
    """one env: obs (P,104,80) -> new obs (oldest..newest)."""
    img = downsample(np.maximum(raw1, raw2))
    return np.concatenate([obs[1:], img[np.newaxis]])


def update_obs_batch(stack, raw_a, raw_b, reset_mask=None):
    """stack (n,P,104,80); raw_a may be None (zeros); reset_mask[i]: stack and raw_a zeroed first."""
    n = stack.shape[0]
    out = np.empty_like(stack)
    for i in range(n):
        rs = reset_mask is not None and reset_mask[i]
        a = np.zeros_like(raw_b[i]) if (raw_a is None or rs) else raw_a[i]
        base = np.zeros_like(stack[i]) if rs else stack[i]
        out[i] = update_obs(base, a, raw_b[i])
    return out
