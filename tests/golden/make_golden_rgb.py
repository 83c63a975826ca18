"""Known-answer fixture for the north-star RGB frame mode (builder-defined arithmetic, oracle/frame.py:rgb_*).
The reference has no RGB path, so this fixture pins the ORACLE against regressions (not against the reference):
    python tests/golden/make_golden_rgb.py   ->  tests/golden/frames_rgb.npz
Frames are low-entropy sprite scenes (rectangles on a flat background) so the archive stays small."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import frame as oframe  # noqa: E402


def scene(rng):
    f = np.empty((210, 160, 3), np.uint8)
    f[:] = rng.randint(0, 256, 3)
    for _ in range(12):
        y, x = rng.randint(0, 200), rng.randint(0, 150)
        h, w = rng.randint(1, 40), rng.randint(1, 40)
        f[y:y + h, x:x + w] = rng.randint(0, 256, 3)
    return f


def main():
    rng = np.random.RandomState(20261017)
    n, P, steps = 3, 4, 3
    raw_a = np.stack([[scene(rng) for _ in range(n)] for _ in range(steps)])
    raw_b = np.stack([[scene(rng) for _ in range(n)] for _ in range(steps)])
    reset = np.zeros((steps, n), np.uint8)
    reset[1, 2] = 1
    stack = np.zeros((n, P, 84, 84), np.uint8)
    outs = []
    for s in range(steps):
        stack = oframe.rgb_update_obs_batch(stack, raw_a[s], raw_b[s], reset[s])
        outs.append(stack.copy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frames_rgb.npz"), raw_a=raw_a, raw_b=raw_b, reset=reset,
                        stacks=np.stack(outs))


if __name__ == "__main__":
    main()
