"""helpers shared by the GPU tests"""
import numpy as np
import torch

from oracle import net as onet


def make_policy(spec_id=1, max_rows=None, seed=0, n_actions=4, planes=4, nonzero_bias=True, hw=(104, 80)):
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.envs.atari_env import EnvSpec
    from accel_rl_b200.spaces import Discrete, UintBox
    spec = cnn_specs[spec_id]
    flat = onet.init_params(onet.CNN_SPECS[spec_id], (planes,) + tuple(hw), n_actions, np.random.RandomState(seed),
                            np.random.RandomState(seed + 1))
    if nonzero_bias:
        flat = flat + np.float32(0.01) * np.random.RandomState(seed + 2).randn(flat.size).astype(np.float32) * (flat == 0)
    pol = AtariCnnPolicy(initial_param_values=flat, max_rows=max_rows, **spec)
    pol.initialize(EnvSpec(UintBox((planes,) + tuple(hw)), Discrete(n_actions)))
    return pol, flat, onet.CNN_SPECS[spec_id]


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def t2n(t):
    return t.detach().cpu().numpy()
