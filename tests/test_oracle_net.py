"""Self-checks of the network oracle (oracle/net.py) against INDEPENDENT restatements — no GPU.

Theano / Lasagne are not available, so the network half of the oracle is "parity unpinned" against the reference's
execution.  What can be pinned is that the oracle computes what its docstring says Lasagne computes:
  * Conv2DLayer(flip_filters=True) = the textbook TRUE convolution: checked against scipy.signal.convolve2d;
  * DenseLayer on the C-order flatten, softmax / linear heads: checked against plain numpy;
  * the losses' gradients (autograd) against central finite differences of the same scalar in fp64;
  * Adam / RMSProp / total_norm_constraint against the closed forms in optimizers/update_methods_stats.py and
    optimizers/util.py:70-76 written out with python floats."""
import math

import numpy as np
import pytest
import torch
from scipy.signal import convolve2d

from oracle import net as onet

SPEC = dict(conv_filter_sizes=[4, 3], conv_filters=[3, 5], conv_strides=[2, 1], conv_pads=[0, 1], hidden_sizes=[7])
IN_SHAPE = (2, 14, 12)
A = 3


def _np_forward(flat, obs):
    ps = onet.unflatten(flat, SPEC, IN_SHAPE, A)
    x = obs.astype(np.float64)
    for l in range(2):
        W, b = ps[2 * l], ps[2 * l + 1]
        s, p = SPEC["conv_strides"][l], SPEC["conv_pads"][l]
        xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
        n, c, h, w = xp.shape
        k = W.shape[2]
        out = np.zeros((n, W.shape[0], (h - k) // s + 1, (w - k) // s + 1))
        for i in range(n):
            for o in range(W.shape[0]):
                acc = sum(convolve2d(xp[i, ci], W[o, ci], mode="valid") for ci in range(c))   # true convolution
                out[i, o] = acc[::s, ::s]
        if l == 0:
            out = out * (1.0 / 255.0)
        x = np.maximum(out + b[None, :, None, None], 0.0)
    x = x.reshape(x.shape[0], -1)
    x = np.maximum(x @ ps[4] + ps[5], 0.0)
    logits = x @ ps[6] + ps[7]
    e = np.exp(logits - logits.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True), (x @ ps[8] + ps[9]).reshape(-1)


def _problem(seed=0, n=6):
    rng = np.random.RandomState(seed)
    flat = onet.init_params(SPEC, IN_SHAPE, A, np.random.RandomState(seed), np.random.RandomState(seed + 1)).astype(np.float64)
    flat += 0.05 * rng.randn(flat.size)
    obs = rng.randint(0, 256, (n,) + IN_SHAPE, dtype=np.uint8)
    act = rng.randint(0, A, n).astype(np.uint8)
    adv, ret = rng.randn(n), rng.randn(n)
    oldp = rng.dirichlet(np.ones(A), n)
    return flat, obs, act, adv, ret, oldp


def test_forward_is_true_convolution_dense_softmax():
    flat, obs, *_ = _problem()
    p, v = onet.forward(torch.tensor(flat), torch.tensor(obs), SPEC, A, dtype=torch.float64)
    p_np, v_np = _np_forward(flat, obs)
    np.testing.assert_allclose(p.numpy(), p_np, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(v.numpy(), v_np, rtol=1e-10, atol=1e-12)
    assert onet.n_params(SPEC, IN_SHAPE, A) == flat.size


@pytest.mark.parametrize("algo", ["ppo", "a2c"])
@pytest.mark.parametrize("use_valids", [False, True])
def test_loss_gradient_matches_finite_differences(algo, use_valids):
    flat, obs, act, adv, ret, oldp = _problem(seed=3)
    valids = np.array([1, 1, 0, 1, 1, 0], np.int8) if use_valids else None

    def loss_at(f):
        with torch.no_grad():
            p, v = onet.forward(torch.tensor(f), torch.tensor(obs), SPEC, A, dtype=torch.float64)
            pl, vl, el = onet.losses(p, v, torch.tensor(act), torch.tensor(adv), torch.tensor(ret), torch.tensor(oldp), algo,
                                     valids=None if valids is None else torch.tensor(valids), v_coeff=0.5)
        return float(pl + vl + el)
    loss, g, parts = onet.loss_and_grad(flat, obs, act, adv, ret, oldp, SPEC, A, algo, dtype=torch.float64, valids=valids,
                                        v_coeff=0.5)
    assert abs(loss - loss_at(flat)) < 1e-12 and abs(sum(parts) - loss) < 1e-12
    rng = np.random.RandomState(1)
    for i in rng.choice(flat.size, 40, replace=False):
        h = 1e-6
        fp, fm = flat.copy(), flat.copy()
        fp[i] += h; fm[i] -= h
        fd = (loss_at(fp) - loss_at(fm)) / (2 * h)
        assert abs(fd - g[i]) <= 1e-6 + 1e-5 * abs(g[i]), (i, fd, g[i])


def test_update_rules_closed_form():
    rng = np.random.RandomState(0)
    n = 50
    p0 = rng.randn(n).astype(np.float32)
    gs = [rng.randn(n).astype(np.float32) * 0.1 for _ in range(3)]
    adam = onet.Adam(n, 1e-3, epsilon=1e-5)
    p = p0.copy()
    m = np.zeros(n); v = np.zeros(n)
    want = p0.astype(np.float64)
    for t, g in enumerate(gs, 1):
        p = adam.step(p, g, 0.5)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g.astype(np.float64) ** 2
        a_t = 1e-3 * 0.5 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)            # update_methods_stats.py:70-73
        want = want - a_t * m / (np.sqrt(v) + 1e-5)
    np.testing.assert_allclose(p, want, rtol=2e-6, atol=1e-7)
    rms = onet.RMSProp(n, 7e-4)
    p = p0.copy(); acc = np.zeros(n); want = p0.astype(np.float64)
    for g in gs:
        p = rms.step(p, g)
        acc = 0.9 * acc + 0.1 * g.astype(np.float64) ** 2                         # update_methods_stats.py:25-29
        want = want - 7e-4 * g / np.sqrt(acc + 1e-6)
    np.testing.assert_allclose(p, want, rtol=2e-6, atol=1e-7)
    g = gs[0]
    norm = math.sqrt(float(np.sum(g.astype(np.float64) ** 2)))
    gc, nn = onet.total_norm_clip(g, 0.05)                                       # optimizers/util.py:70-76
    assert abs(nn - norm) < 1e-9 and norm > 0.05
    np.testing.assert_allclose(gc, g * (0.05 / (1e-7 + norm)), rtol=1e-6)
    gc2, _ = onet.total_norm_clip(g, None)
    assert gc2 is g
