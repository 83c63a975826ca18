"""Discrete action space (reference: accel_rl/spaces/discrete.py)."""
import numpy as np


def weighted_sample_n(prob_matrix, items):
    """Host version of rllab/misc/special.py:22-27 — consumes np.random.rand(n) from the global
    legacy stream exactly like the reference (used outside the device rollout, e.g. get_actions
    on host arrays)."""
    s = prob_matrix.cumsum(axis=1)
    r = np.random.rand(prob_matrix.shape[0])
    k = (s < r.reshape((-1, 1))).sum(axis=1)
    return items[np.minimum(k, len(items) - 1)]


class Discrete(object):
    def __init__(self, n):
        self._n = int(n)
        self._dtype = "uint8" if n <= 2 ** 8 else ("uint16" if n <= 2 ** 16 else "uint32")  # discrete.py:12-18
        self._items_arr = np.arange(n).astype(self._dtype)

    n = property(lambda self: self._n)
    dtype = property(lambda self: self._dtype)
    flat_dim = property(lambda self: self._n)
    default_value = 0

    def sample(self):
        return np.random.randint(self._n, dtype=self._dtype)

    def sample_n(self, n):
        return np.random.randint(low=0, high=self._n, size=n, dtype=self._dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == () and x.dtype.kind in "iu" and 0 <= x < self._n

    def weighted_sample(self, weights):
        cs = np.cumsum(weights)
        idx = int((cs < np.random.rand()).sum())
        return self._items_arr[min(idx, self._n - 1)]

    def weighted_sample_n(self, weights_matrix):
        return weighted_sample_n(np.asarray(weights_matrix), self._items_arr)

    def __eq__(self, other):
        return isinstance(other, Discrete) and other.n == self._n

    def __hash__(self):
        return hash(self._n)

    def __repr__(self):
        return "Discrete(%d)" % self._n
