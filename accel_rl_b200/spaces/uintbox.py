"""Unsigned-integer box observation space (reference: accel_rl/spaces/uintbox.py)."""
import numpy as np


class UintBox(object):
    def __init__(self, shape, low=0, high=None, bits=8):
        assert bits in (8, 16, 32, 64)
        self.dtype = "uint%d" % bits
        top = 2 ** bits - 1
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(top if high is None else high, dtype=self.dtype)
        assert 0 <= self.low < self.high <= top
        self._shape = tuple(shape)

    shape = property(lambda self: self._shape)
    flat_dim = property(lambda self: int(np.prod(self._shape)))
    bounds = property(lambda self: (self.low, self.high))

    def sample(self):
        return np.random.randint(low=self.low, high=self.high, size=self._shape, dtype=self.dtype)

    def sample_n(self, n):
        return np.random.randint(low=self.low, high=self.high, size=(n,) + self._shape, dtype=self.dtype)

    def contains(self, x):
        return x.shape == self._shape and (x >= self.low).all() and (x <= self.high).all()

    def __eq__(self, other):
        return isinstance(other, UintBox) and other.shape == self._shape and other.dtype == self.dtype

    def __repr__(self):
        return "UintBox%s" % (self._shape,)
