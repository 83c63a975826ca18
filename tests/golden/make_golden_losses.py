"""Golden vectors for the PPO / A2C loss TERMS, produced by EXECUTING the reference's own statements of them under
oracle/theano_shim.py's eager float32 numpy stand-in for `theano.tensor` (inputs are arrays, so each symbolic expression
evaluates to its value):

  policy term   accel_rl/algos/pg/ppo.py  `BasePPO.pi_loss`  (ratio, clip by clip_param * lr_mult, min of the surrogates)
                accel_rl/algos/pg/a2c.py  `BaseA2C.pi_loss`  (log-likelihood * advantage)
                — both files import `accel_rl.optimizers.async…` (a keyword since Python 3.7) and do not parse here, so the
                  `pi_loss` method is cut out of the file text at run time and exec'd; nothing is copied into this repo
  value / entropy terms, pi_kl / v_kl   accel_rl/algos/pg/aac_base.py, the statements between `v_err =` and `constraints =`
                of `initialize` (cut out and exec'd the same way)
  distribution  accel_rl/distributions/categorical.py (loaded whole, unmodified): likelihood_ratio_sym, log_likelihood_sym,
                entropy_sym, kl_sym
  means         accel_rl/algos/pg/util.py: valids_mean (loaded whole, unmodified)

Only runnable in the build container; tests/golden/loss_terms.npz is what travels.

    python tests/golden/make_golden_losses.py
"""
import importlib.util
import os
import re
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import theano_shim as S  # noqa: E402

REF = "/root/reference/accel_rl"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cut_method(path, cls, name):
    """source of method `name` of class `cls` in `path`, dedented (the file itself need not parse)"""
    text = open(path).read()
    start = text.index("class %s(" % cls)
    m = re.search(r"\n    def %s\(.*?(?=\n    def |\nclass |\Z)" % name, text[start:], re.S)
    return textwrap.dedent(m.group(0).strip("\n"))


def cut_statements(path, first, last):
    """the statements from the line starting with `first` up to and including the one starting with `last`"""
    out, on = [], False
    for line in open(path).read().splitlines():
        if line.strip().startswith(first):
            on = True
        if on:
            out.append(line)
        if on and line.strip().startswith(last):
            break
    return textwrap.dedent("\n".join(out))


def main():
    saved = S.install_eager_tensor(extra_stubs=("rllab", "rllab.distributions", "rllab.distributions.base",
                                                "theano.sandbox", "theano.sandbox.rng_mrg"))
    try:
        import theano.tensor as T
        cat = load(os.path.join(REF, "distributions", "categorical.py"), "ref_categorical")
        util = load(os.path.join(REF, "algos", "pg", "util.py"), "ref_pg_util")
        ns = dict(T=T, valids_mean=util.valids_mean, np=np)
        exec(cut_method(os.path.join(REF, "algos", "pg", "ppo.py"), "BasePPO", "pi_loss"), ns)
        ppo_pi_loss = ns.pop("pi_loss")
        exec(cut_method(os.path.join(REF, "algos", "pg", "a2c.py"), "BaseA2C", "pi_loss"), ns)
        a2c_pi_loss = ns.pop("pi_loss")
        rest = cut_statements(os.path.join(REF, "algos", "pg", "aac_base.py"), "v_err =", "constraints =")
        dist = cat.Categorical(6)
        policy = types.SimpleNamespace(distribution=dist)

        rng = np.random.RandomState(12)
        n, A = 257, 6
        def softmax(z):
            e = np.exp(z - z.max(axis=1, keepdims=True))
            return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
        old_prob = softmax(rng.randn(n, A))
        new_prob = softmax(np.log(old_prob) + 0.35 * rng.randn(n, A))      # ratios on both sides of the clip range
        new_prob[:40] = old_prob[:40]                                       # exact ties of the two surrogates
        act = rng.randint(0, A, n).astype(np.uint8)
        adv = rng.randn(n).astype(np.float32)
        ret = rng.randn(n).astype(np.float32)
        new_value = (ret + 0.3 * rng.randn(n)).astype(np.float32)
        old_value = (new_value + 0.1 * rng.randn(n)).astype(np.float32)
        valids_arr = (rng.rand(n) < 0.7).astype(np.int8)
        out = dict(old_prob=old_prob, new_prob=new_prob, act=act, adv=adv, ret=ret, new_value=new_value,
                   old_value=old_value, valids=valids_arr)
        cases = []
        for algo, v_coeff in (("ppo", 1.0), ("a2c", 0.25)):
            for use_valids in (False, True):
                for lr_mult in ((1.0, 0.6) if algo == "ppo" else (1.0,)):
                    self_ = types.SimpleNamespace(clip_param=0.2, _lr_mult=lr_mult, v_loss_coeff=v_coeff, ent_loss_coeff=0.01)
                    self_.pi_loss = types.MethodType(ppo_pi_loss if algo == "ppo" else a2c_pi_loss, self_)
                    env = dict(ns, self=self_, policy=policy, dist=dist, act=act.astype(np.int64), adv=adv, ret=ret,
                               new_value=new_value, old_value=old_value, valids=valids_arr if use_valids else None,
                               old_dist_info=dict(prob=old_prob), new_dist_info=dict(prob=new_prob))
                    exec(rest, env)
                    tag = "%s_v%d_lr%g" % (algo, int(use_valids), lr_mult)
                    cases.append(tag)
                    out[tag] = np.array([env["pi_loss"], env["v_loss"], env["ent_loss"], env["pi_kl"], env["v_kl"]],
                                        dtype=np.float64)
        out["cases"] = np.array(cases)
        np.savez_compressed(os.path.join(HERE, "loss_terms.npz"), **out)
        for c in cases:
            print(c, out[c])
    finally:
        S.restore(saved)


if __name__ == "__main__":
    main()
