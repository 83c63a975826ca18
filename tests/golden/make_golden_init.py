"""Golden vectors for the dense-layer initialiser, produced by EXECUTING the reference's NormCInit.sample
(accel_rl/policies/layers.py:9-19, loaded unmodified through oracle/ref_harness.py's stub modules) on the global numpy
stream.  Only runnable in the build container; tests/golden/norm_c_init.npz is what travels.

    python tests/golden/make_golden_init.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness  # noqa: E402


def main():
    ref_harness.install()
    import theano
    theano.config.floatX = "float32"
    layers = ref_harness.ref("accel_rl.policies.layers")
    out = dict()
    np.random.seed(31)
    # the order the network builder draws them (policies/pg/networks/pg_cnn.py: hidden layer, pi head, v head)
    for name, shape, std in (("hidden", (96, 32), 1.0), ("pi", (32, 6), 0.01), ("v", (32, 1), 1.0)):
        out[name] = layers.NormCInit(std).sample(shape)
        assert out[name].dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "norm_c_init.npz"), **out)
    print({k: (v.shape, float(np.abs(v).sum())) for k, v in out.items()})


if __name__ == "__main__":
    main()
