"""Golden vectors for the optimiser update rules, produced by EXECUTING the reference's own in-tree statement of them
(/root/reference/accel_rl/optimizers/update_methods_stats.py:11-32 rmsprop, :55-87 adam, loaded unmodified from where it
lies) under oracle/theano_shim.py's eager float32 numpy stand-in for the few Theano/Lasagne names the file uses.
Only runnable in the build container; tests/golden/update_rules.npz is what travels.

    python tests/golden/make_golden_updates.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import theano_shim as S  # noqa: E402

REF_FILE = "/root/reference/accel_rl/optimizers/update_methods_stats.py"


def load_reference(session):
    saved = S.install(session)
    spec = importlib.util.spec_from_file_location("ref_update_methods_stats", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, saved


def run(kind, shapes, grads, lr_mults, **hyper):
    """-> list of flat parameter vectors after every step; parameters live in Shared variables like the reference's"""
    sess = S.Session()
    mod, saved = load_reference(sess)
    try:
        rng = np.random.RandomState(21)
        params = [S.Shared((rng.randn(*s) * 0.05).astype(np.float32)) for s in shapes]
        p0 = np.concatenate([p.value.ravel() for p in params])
        outs = []
        for g, lm in zip(grads, lr_mults):
            gl, i = [], 0
            for s in shapes:
                n = int(np.prod(s))
                gl.append(g[i:i + n].reshape(s))
                i += n
            fn = mod.adam if kind == "adam" else mod.rmsprop
            # the learner multiplies the learning rate by the lr_mult input (algos/pg/ppo.py:27, a2c.py:26)
            kw = dict(hyper)
            kw["learning_rate"] = np.float32(hyper["learning_rate"]) * np.float32(lm)
            sess.step(fn, gl, params, **kw)
            outs.append(np.concatenate([p.value.ravel() for p in params]))
        return p0, np.stack(outs)
    finally:
        S.restore(saved)


def main():
    shapes = [(6, 3, 2, 2), (6,), (40, 8), (8,), (8, 4), (4,), (8, 1), (1,)]
    n = int(sum(int(np.prod(s)) for s in shapes))
    rng = np.random.RandomState(4)
    steps = 6
    grads = (rng.randn(steps, n) * 0.02).astype(np.float32)
    grads[2, :7] = 0.0                      # exact zeros (dead units): v stays 0 where g was always 0
    grads[:, 5] = 0.0
    lr_mults = np.array([1.0, 0.9, 0.8, 0.7, 0.6, 0.5], np.float32)
    out = dict(shapes=np.array([",".join(map(str, s)) for s in shapes]), grads=grads, lr_mults=lr_mults)
    p0, pa = run("adam", shapes, grads, lr_mults, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-5)
    out["p0"] = p0
    out["adam"] = pa
    p0b, pr = run("rmsprop", shapes, grads, lr_mults, learning_rate=7e-4, rho=0.9, epsilon=1e-6)
    assert np.array_equal(p0, p0b)
    out["rmsprop"] = pr
    np.savez_compressed(os.path.join(HERE, "update_rules.npz"), **out)
    print("update_rules.npz written:", n, "parameters,", steps, "steps")


if __name__ == "__main__":
    main()
