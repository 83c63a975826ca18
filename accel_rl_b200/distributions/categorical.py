"""Categorical distribution — host-side numerics (reference: accel_rl/distributions/categorical.py).
The symbolic (*_sym) members of the reference are Theano expressions; their arithmetic lives in
head_kernel (csrc/kernels.cuh) here.  TINY matches categorical.py:6."""
import numpy as np

TINY = 1e-8


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


class Categorical(object):
    def __init__(self, dim):
        self._dim = dim

    dim = property(lambda self: self._dim)
    dist_info_keys = property(lambda self: ["prob"])

    def kl(self, old_dist_info, new_dist_info):
        p, q = _np(old_dist_info["prob"]), _np(new_dist_info["prob"])
        return np.sum(p * (np.log(p + TINY) - np.log(q + TINY)), axis=-1)

    def entropy(self, info):
        p = _np(info["prob"])
        return -np.sum(p * np.log(p + TINY), axis=1)

    def log_likelihood(self, xs, dist_info):
        p = _np(dist_info["prob"])
        xs = _np(xs).astype(np.int64)
        return np.log(p[np.arange(len(xs)), xs] + TINY)
