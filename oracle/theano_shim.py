"""ORACLE (test infrastructure, never imported by the product path).

A tiny EAGER numpy stand-in for the handful of Theano / Lasagne names that the reference's in-tree copy of the
update rules uses (accel_rl/optimizers/update_methods_stats.py:11-32 rmsprop, :55-87 adam — "exact copy from
Lasagne updates"), so that file can be EXECUTED unmodified in the build container and its outputs committed as golden
vectors (tests/golden/make_golden_updates.py -> tests/golden/update_rules.npz).

What it models and how:
  * theano.shared(value)       -> a Shared holding a numpy array; arithmetic on it evaluates immediately on the current
                                  value.  The reference functions create their state (accumulators, m, v, t) with
                                  theano.shared *inside* the call; re-tracing the function every step therefore has to
                                  hand back the SAME state objects: shared() calls are matched by call order within a
                                  `session.step()` (first step creates them, later steps return them).
  * the returned `updates`     -> an OrderedDict {Shared: new value}.  Every expression was evaluated from the
                                  pre-update values (nothing is written before apply_updates), which is Theano's
                                  simultaneous-update semantics.
  * dtypes                     -> floatX = float32.  Theano casts Python float literals to floatX and int8
                                  constants promote to float32 against float32 operands; numpy >= 2 (NEP 50) does the
                                  same for float32 arrays combined with Python scalars, and T.constant(1) is handed
                                  back as float32(1).  Every operation is therefore a float32 numpy operation.
Nothing here is derived from Theano source (absent from this machine); it is the builder's statement of those
semantics, used ONLY to run reference code that is otherwise pure arithmetic.
"""
import sys
import types
from collections import OrderedDict

import numpy as np


class Shared(object):
    __array_priority__ = 1000.0

    def __init__(self, value):
        self.value = np.array(value)
        self.broadcastable = (False,) * self.value.ndim

    def get_value(self, borrow=False):
        return self.value

    def set_value(self, v):
        self.value = np.asarray(v, dtype=self.value.dtype).reshape(self.value.shape)

    # eager arithmetic on the current value
    def __add__(self, o): return self.value + _val(o)
    def __radd__(self, o): return _val(o) + self.value
    def __sub__(self, o): return self.value - _val(o)
    def __rsub__(self, o): return _val(o) - self.value
    def __mul__(self, o): return self.value * _val(o)
    def __rmul__(self, o): return _val(o) * self.value
    def __truediv__(self, o): return self.value / _val(o)
    def __rtruediv__(self, o): return _val(o) / self.value
    def __pow__(self, o): return self.value ** _val(o)
    def __rpow__(self, o): return _val(o) ** self.value
    def __hash__(self): return id(self)
    def __eq__(self, o): return self is o


def _val(x):
    return x.value if isinstance(x, Shared) else x


class Session(object):
    """Re-traces a reference update function once per optimisation step with persistent shared state."""

    def __init__(self):
        self.vars = []
        self.cursor = 0

    def shared(self, value, **kw):
        if self.cursor < len(self.vars):
            v = self.vars[self.cursor]
        else:
            v = Shared(value)
            self.vars.append(v)
        self.cursor += 1
        return v

    def step(self, fn, *args, **kw):
        """call fn (rmsprop / adam), apply its updates simultaneously, return the per-parameter steps"""
        self.cursor = 0
        updates, steps = fn(*args, **kw)
        new = [(k, np.array(v, dtype=k.value.dtype)) for k, v in updates.items()]
        for k, v in new:
            k.set_value(v)
        return steps


def install(session):
    """register the shim modules under the names update_methods_stats.py imports; returns the previous entries"""
    names = ["theano", "theano.tensor", "lasagne", "lasagne.updates", "lasagne.utils"]
    saved = {n: sys.modules.get(n) for n in names}
    th = types.ModuleType("theano")
    tt = types.ModuleType("theano.tensor")
    la = types.ModuleType("lasagne")
    lu = types.ModuleType("lasagne.updates")
    lut = types.ModuleType("lasagne.utils")
    th.shared = session.shared
    th.tensor = tt
    th.config = types.SimpleNamespace(floatX="float32")
    tt.constant = lambda x: np.float32(x)
    tt.sqrt = np.sqrt
    tt.reshape = lambda x, shp: np.reshape(x, shp)
    lu.get_or_compute_grads = lambda loss_or_grads, params: list(loss_or_grads)   # grads are handed in as a list
    lut.floatX = lambda x: np.asarray(x, dtype=np.float32)
    la.updates, la.utils, la.util = lu, lut, lut
    for n, m in zip(names, (th, tt, la, lu, lut)):
        sys.modules[n] = m
    return saved


def install_eager_tensor(extra_stubs=()):
    """For reference code that only BUILDS elementwise expressions (the loss terms of algos/pg/{ppo,a2c,aac_base}.py,
    distributions/categorical.py:*_sym, algos/pg/util.py:valids_mean): `theano.tensor` as eager float32 numpy.  Symbolic
    inputs are numpy arrays, so every `*_sym` function returns the VALUE of its expression.  `extra_stubs`: module names
    that the loaded files import but never use on these paths (registered as empty modules with a permissive base class).
    Returns the previous sys.modules entries for restore()."""
    names = ["theano", "theano.tensor"] + list(extra_stubs)
    saved = {n: sys.modules.get(n) for n in names}
    th = types.ModuleType("theano")
    tt = types.ModuleType("theano.tensor")
    f32 = lambda x: np.asarray(x, dtype=np.float32) if np.asarray(x).dtype.kind == "f" else np.asarray(x)
    tt.clip = lambda x, lo, hi: np.clip(f32(x), np.float32(lo), np.float32(hi))
    tt.minimum = lambda a, b: np.minimum(f32(a), f32(b))
    tt.maximum = lambda a, b: np.maximum(f32(a), f32(b))
    tt.mean = lambda x, axis=None: np.mean(f32(x), axis=axis, dtype=np.float32)
    tt.sum = lambda x, axis=None: np.sum(f32(x), axis=axis, dtype=np.float32 if np.asarray(x).dtype.kind == "f" else None)
    tt.log = lambda x: np.log(f32(x))
    tt.exp = lambda x: np.exp(f32(x))
    tt.sqr = lambda x: f32(x) * f32(x)
    tt.arange = np.arange
    tt.shape = lambda x: np.asarray(x).shape
    tt.cast = lambda x, dt: np.asarray(x).astype(dt)
    th.tensor = tt
    th.config = types.SimpleNamespace(floatX="float32")
    sys.modules["theano"], sys.modules["theano.tensor"] = th, tt
    for n in extra_stubs:
        m = types.ModuleType(n)
        m.__getattr__ = lambda item, _n=n: type(item, (object,), {"__init__": lambda self, *a, **k: None})
        sys.modules[n] = m
    for n in extra_stubs:                       # wire parents -> children where both are stubs
        if "." in n:
            parent, child = n.rsplit(".", 1)
            if parent in sys.modules:
                setattr(sys.modules[parent], child, sys.modules[n])
    return saved


def restore(saved):
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m


__all__ = ["Shared", "Session", "install", "install_eager_tensor", "restore", "OrderedDict"]
