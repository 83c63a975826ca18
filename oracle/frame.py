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


# ---------------------------------------------------------------------------------------------------------
# North-star frame mode (BASELINE.json north_star): RGB -> gray -> 84x84.  NOT in the reference (its emulator
# returns grayscale, envs/atari_env.py:147-149), so the arithmetic below is BUILDER-DEFINED and frozen here —
# parity unpinned (SURVEY.md §8c "RGB->gray").  Integer-exact by construction.
# ---------------------------------------------------------------------------------------------------------
NS_H, NS_W = 84, 84


def rgb_to_gray(rgb):
    """(...,3) u8 -> (...) u8: Y = (77 R + 150 G + 29 B + 128) >> 8 (NTSC luma in 8-bit fixed point)."""
    r = rgb.astype(np.uint32)
    return ((77 * r[..., 0] + 150 * r[..., 1] + 29 * r[..., 2] + 128) >> 8).astype(np.uint8)


def _area_weights(n_in, n_out, unit_in, unit_out):
    """integer overlap lengths of output cell o = [unit_out*o, unit_out*(o+1)) with input cell i = [unit_in*i, ...)"""
    w = np.zeros((n_out, n_in), np.int64)
    for o in range(n_out):
        lo, hi = unit_out * o, unit_out * (o + 1)
        for i in range(n_in):
            a, b = max(lo, unit_in * i), min(hi, unit_in * (i + 1))
            if b > a:
                w[o, i] = b - a
    return w


_WY = _area_weights(210, NS_H, 2, 5)      # rows: 210/84 = 5/2, sum 5
_WX = _area_weights(160, NS_W, 21, 40)    # cols: 160/84 = 40/21, sum 40


def rgb_downsample(gray):
    """(210,160) u8 -> (84,84) u8: exact area average, (sum wy*wx*Y + 100) // 200."""
    s = _WY @ gray.astype(np.int64) @ _WX.T
    return ((s + 100) // 200).astype(np.uint8)


def rgb_update_obs_batch(stack, raw_a, raw_b, reset_mask=None):
    """stack (n,P,84,84) u8; raw_* (n,210,160,3) u8 (raw_a may be None); -> new stack (oldest..newest)."""
    n = stack.shape[0]
    out = np.empty_like(stack)
    for i in range(n):
        rs = reset_mask is not None and reset_mask[i]
        a = np.zeros_like(raw_b[i]) if (raw_a is None or rs) else raw_a[i]
        base = np.zeros_like(stack[i]) if rs else stack[i]
        img = rgb_downsample(rgb_to_gray(np.maximum(a, raw_b[i])))
        out[i] = np.concatenate([base[1:], img[np.newaxis]])
    return out
