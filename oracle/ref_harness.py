"""ORACLE (test infrastructure).  Import harness for the REAL reference source.

Only usable where /root/reference exists (the build container).  It registers permissive stub
modules for the engines that are absent (theano, lasagne, pyprind, path, posix_ipc), a fake
`atari_py` backed by oracle.synth_ale.SynthALE, and aliases the removed inspect.getargspec, then
imports the reference's modules by their real dotted names so its own code runs unmodified:
AtariEnv, ActsrvAltOvrlpSampler (forked workers, shared buffers, semaphores), buffers,
gen_adv_est / discount_returns, iterate_mb_idxs, weighted_sample_n, set_seed.
Nothing is copied from the reference; tests/golden/make_golden.py uses this to generate fixtures.
"""
import importlib
import inspect
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "accel_rl"))


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub; callable; usable as a base class."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__all__ = []

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        full = self.__name__ + "." + item
        if full in sys.modules:
            return sys.modules[full]
        child = _Stub(full)
        sys.modules[full] = child
        setattr(self, item, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __mro_entries__(self, bases):
        return (object,)


_STUB_NAMES = [
    "theano", "theano.tensor", "theano.tensor.nnet", "theano.tensor.extra_ops", "theano.tensor.signal",
    "theano.sandbox", "theano.sandbox.rng_mrg", "theano.gpuarray", "theano.ifelse", "theano.gradient",
    "lasagne", "lasagne.layers", "lasagne.init", "lasagne.nonlinearities", "lasagne.updates", "lasagne.utils",
    "lasagne.random", "lasagne.regularization", "pyprind", "path", "posix_ipc",
]

_installed = False


def install(pool=None, rules=None):
    """Put the stubs and the fake atari_py in sys.modules; idempotent."""
    global _installed
    from oracle import synth_ale
    if rules is not None:
        synth_ale.SynthALE.rules = dict(rules)
    if pool is not None:
        synth_ale.SynthALE.pool = pool
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec
    for n in _STUB_NAMES:
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)
    for n in _STUB_NAMES:  # wire parents -> children
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[n])
    # lasagne.random.set_rng / get_rng are called by rllab.misc.ext.set_seed
    import numpy as np
    lr = sys.modules["lasagne.random"]
    lr._rng = np.random
    lr.set_rng = lambda r: setattr(lr, "_rng", r)
    lr.get_rng = lambda: lr._rng
    fake = types.ModuleType("atari_py")
    fake.ALEInterface = synth_ale.SynthALE
    fake.get_game_path = lambda game: os.path.join(REF_ROOT, "README.md")  # any existing path
    sys.modules["atari_py"] = fake
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def ref(module):
    """import a reference module by dotted name (after install())."""
    return importlib.import_module(module)


def _worker_target(**kwargs):
    """Wraps the reference's sampling_process: gives every forked worker a distinct env-id base so
    the synthetic emulator of env i of worker (group, rank) is env id unique_ID*envs_per + i,
    which is that env's position in the reference's env-major buffer layout."""
    from oracle import synth_ale
    synth_ale.SynthALE.next_env_id = kwargs["unique_ID"] * kwargs["envs_per"]
    worker = ref("accel_rl.sampler.act_server.alternating.overlap.worker")
    return worker.sampling_process(**kwargs)


def make_sampler(n_parallel, envs_per, horizon, mid_batch_reset=True, max_path_length=27000, env_args=None):
    """A real ActsrvAltOvrlpSampler over the reference AtariEnv + SynthALE."""
    env_mod = ref("accel_rl.envs.atari_env")
    smp_mod = ref("accel_rl.sampler.act_server.alternating.overlap.sampler")
    args = dict(game="breakout", max_start_noops=0)
    if env_args:
        args.update(env_args)
    sampler = smp_mod.ActsrvAltOvrlpSampler(
        EnvCls=env_mod.AtariEnv, env_args=args, horizon=horizon, n_parallel=n_parallel, envs_per=envs_per,
        max_path_length=max_path_length, mid_batch_reset=mid_batch_reset, max_decorrelation_steps=0)
    return sampler


def initialize_sampler(sampler, seed, discount, need_extra_obs=True):
    from oracle import synth_ale
    synth_ale.SynthALE.next_env_id = 10 ** 6  # the master's example env (its id is irrelevant)
    return sampler.initialize(seed=seed, affinities=dict(), discount=discount, need_extra_obs=need_extra_obs,
                              worker_process_target=_worker_target)
