"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement (torch, fp32 or fp64) of the Theano/Lasagne graph the reference builds for the
A2C/PPO Atari path.  Theano and Lasagne are absent from /root/reference and from this image
(environment.yml:27 unpinned theano, :51 Lasagne@20efd95): the semantics below are the
builder-frozen definition of those libraries' behaviour — **parity unpinned** by reference tests
(the reference has none).  What is restated, and from where:

  network   accel_rl/policies/pg/networks/pg_cnn.py:45-86, policies/layers.py:22-40
            (uint8 -> * 1/255 -> Conv2DLayer(ReLU)* -> DenseLayer(ReLU) -> {softmax pi, linear V})
            Lasagne Conv2DLayer: W (out,in,kh,kw), flip_filters=True (true convolution);
            DenseLayer flattens trailing dims in C order, W (in,out).
  init      policies/layers.py:11-19 (NormCInit), pg_cnn.py:25-29 (GlorotUniform gain 1)
  params    rllab/core/parameterized.py:74-88 flat vector in Lasagne get_all_params order
  losses    algos/pg/aac_base.py:60-70, ppo.py:42-51, a2c.py:43-46,
            distributions/categorical.py:35-88 (TINY=1e-8), algos/pg/util.py:49-53 (valids_mean)
  clip      optimizers/util.py:70-76 -> Lasagne total_norm_constraint(epsilon=1e-7)
  updates   optimizers/update_methods_stats.py:11-32 (rmsprop), :55-87 (adam)
  loops     optimizers/single/ppo_optimizer.py:57-76, single/a2c_optimizer.py:45,
            optimizers/util.py:8-18 (iterate_mb_idxs), sync: optimizers/util.py:63-67 (x 1/n_gpu)

PPO surrogate gradient at exact ties of min(surr1, surr2) is taken once (torch.minimum splits it,
the clip passes it inside the closed interval) — the standard PPO gradient.

`emulate_bf16=True` rounds the tensor-core operands (conv/FC weights and the stored activations)
to bf16 at the points the CUDA path does, so parity tests can use tight tolerances.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

TINY = 1e-8

CNN_SPECS = {  # accel_rl/policies/atari_cnn_specs.py:10-32
    0: dict(conv_filter_sizes=[8, 4], conv_filters=[16, 32], conv_strides=[4, 2], conv_pads=[0, 1],
            hidden_sizes=[256]),
    1: dict(conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64], conv_strides=[4, 2, 1],
            conv_pads=[0, 1, 1], hidden_sizes=[512]),
}


def param_shapes(spec, in_shape, n_actions):
    """Lasagne get_all_params order: conv W,b ..., hidden W,b, pi W,b, v W,b."""
    c, h, w = in_shape
    shapes = []
    for f, k, s, p in zip(spec["conv_filters"], spec["conv_filter_sizes"], spec["conv_strides"],
                          spec["conv_pads"]):
        shapes.append((f, c, k, k))
        shapes.append((f,))
        h = (h + 2 * p - k) // s + 1
        w = (w + 2 * p - k) // s + 1
        c = f
    n_in = c * h * w
    for hs in spec["hidden_sizes"]:
        shapes.append((n_in, hs))
        shapes.append((hs,))
        n_in = hs
    shapes += [(n_in, n_actions), (n_actions,), (n_in, 1), (1,)]
    return shapes


def n_params(spec, in_shape, n_actions):
    return int(sum(int(np.prod(s)) for s in param_shapes(spec, in_shape, n_actions)))


def init_params(spec, in_shape, n_actions, conv_rng, global_rng):
    """Flat fp32 parameter vector.  conv W: GlorotUniform from the lasagne rng (set_seed gives it
    its own RandomState, rllab/misc/ext.py:198-207); dense W: NormCInit from the GLOBAL rng
    (np.random.randn, layers.py:16-19) with std 1.0 / 0.01 / 1.0; biases 0."""
    shapes = param_shapes(spec, in_shape, n_actions)
    n_conv = len(spec["conv_filters"])
    n_hidden = len(spec["hidden_sizes"])
    out = []
    for i, shp in enumerate(shapes):
        if len(shp) == 4:  # conv W
            f, c, kh, kw = shp
            fan_in, fan_out = c * kh * kw, f * kh * kw
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            out.append(conv_rng.uniform(-lim, lim, size=shp).astype(np.float32))
        elif len(shp) == 2:
            layer = (i - 2 * n_conv) // 2
            std = 0.01 if layer == n_hidden else 1.0  # pi head 0.01, hidden and v head 1.0
            w = global_rng.randn(*shp).astype(np.float32)
            w *= std / np.sqrt(np.square(w).sum(axis=0, keepdims=True))
            out.append(w)
        else:
            out.append(np.zeros(shp, np.float32))
    return np.concatenate([a.ravel() for a in out])


def unflatten(flat, spec, in_shape, n_actions):
    shapes = param_shapes(spec, in_shape, n_actions)
    out, i = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(flat[i:i + n].reshape(s))
        i += n
    assert i == flat.numel() if isinstance(flat, torch.Tensor) else i == flat.size
    return out


def _bf16(x):
    """round-to-nearest-even to bf16, straight-through for autograd"""
    r = x.detach().to(torch.float32).to(torch.bfloat16).to(x.dtype)
    return x + (r - x).detach()


def forward(flat, obs_u8, spec, n_actions, emulate_bf16=False, dtype=torch.float32, pixel_scale=255.0,
            return_acts=False):
    """-> prob (n, A), value (n,).  obs_u8: uint8 tensor (n, C, H, W)."""
    in_shape = tuple(obs_u8.shape[1:])
    ps = unflatten(flat, spec, in_shape, n_actions)
    n_conv = len(spec["conv_filters"])
    x = obs_u8.to(dtype)
    acts = []
    for l in range(n_conv):
        W, b = ps[2 * l], ps[2 * l + 1]
        Wc = torch.flip(W, dims=(2, 3))  # flip_filters=True: true convolution == correlation with flipped W
        if emulate_bf16:
            Wc = _bf16(Wc)
        s, p = spec["conv_strides"][l], spec["conv_pads"][l]
        if l == 0:
            # scale layer: input * (1/scale) (layers.py:40); the CUDA path scales the accumulator
            y = F.conv2d(x, Wc, None, stride=s, padding=p) * (1.0 / pixel_scale) + b.view(1, -1, 1, 1)
        else:
            y = F.conv2d(x, Wc, b, stride=s, padding=p)
        x = torch.relu(y)
        if emulate_bf16:
            x = _bf16(x)
        acts.append(x)
    x = x.reshape(x.shape[0], -1)  # C-order flatten of (C,H,W)
    k = 2 * n_conv
    for _ in spec["hidden_sizes"]:
        W, b = ps[k], ps[k + 1]
        if emulate_bf16:
            W = _bf16(W)
        x = torch.relu(x @ W + b)
        if emulate_bf16:
            x = _bf16(x)
        acts.append(x)
        k += 2
    logits = x @ ps[k] + ps[k + 1]
    prob = torch.softmax(logits, dim=1)
    value = (x @ ps[k + 2] + ps[k + 3]).reshape(-1)
    if return_acts:
        return prob, value, acts
    return prob, value


def losses(prob, value, act, adv, ret, old_prob, algo, clip_param=0.2, lr_mult=1.0, v_coeff=1.0,
           ent_coeff=0.01, valids=None, tie_grad=1):
    """(pi_loss, v_loss, ent_loss) — aac_base.py:60-70; algo in {"ppo", "a2c"}.
    tie_grad: inside the clip range surr_1 == surr_2 exactly; 1 = min() passes the gradient once (Theano >= 0.9
    Minimum.L_op, standard PPO), 2 = to both branches (older Theano: eq(out, x) * gz and eq(out, y) * gz)."""
    def vmean(x):
        if valids is None:
            return x.mean()
        v = valids.to(x.dtype)
        return (v * x).sum() * (1.0 / v.sum())

    idx = torch.arange(prob.shape[0])
    a = act.long()
    if algo == "ppo":
        ratio = (prob[idx, a] + TINY) / (old_prob[idx, a] + TINY)
        cp = clip_param * lr_mult
        s1, s2 = ratio * adv, torch.clamp(ratio, 1.0 - cp, 1.0 + cp) * adv
        surr = torch.minimum(s1, s2)
        if tie_grad == 2:
            surr = surr + (s1 - s1.detach()) * (s1 == s2).to(s1.dtype)     # same value, one more gradient share at ties
        pi_loss = -vmean(surr)
    else:
        pi_loss = -vmean(torch.log(prob[idx, a] + TINY) * adv)
    v_loss = v_coeff * vmean((value - ret) ** 2)
    ent = -(prob * torch.log(prob + TINY)).sum(dim=1)
    ent_loss = -ent_coeff * vmean(ent)
    return pi_loss, v_loss, ent_loss


def loss_and_grad(flat_np, obs_u8, act, adv, ret, old_prob, spec, n_actions, algo, emulate_bf16=False,
                  dtype=torch.float32, valids=None, **kw):
    """-> (loss float, flat grad np.float32/64).  Gradients by autograd (oracle only)."""
    flat = torch.tensor(flat_np, dtype=dtype, requires_grad=True)
    prob, value = forward(flat, torch.as_tensor(obs_u8), spec, n_actions, emulate_bf16, dtype)
    t = lambda x: None if x is None else torch.as_tensor(np.asarray(x))
    pl, vl, el = losses(prob, value, t(act), t(adv).to(dtype), t(ret).to(dtype), t(old_prob).to(dtype), algo,
                        valids=t(valids), **kw)
    loss = pl + vl + el
    loss.backward()
    return float(loss), flat.grad.detach().numpy(), (float(pl), float(vl), float(el))


def total_norm_clip(grad, clip):
    """optimizers/util.py:70-76 -> (grad', pre-clip norm).  clip None: norm only."""
    norm = float(np.sqrt(np.sum(np.square(grad.astype(np.float64)))))
    if clip is None or clip <= 0:
        return grad, norm
    return grad * np.float32(min(norm, clip) / (1e-7 + norm)), norm


class Adam:
    """update_methods_stats.py:55-87 (one shared t; PPO uses epsilon=1e-5, ppo.py:26)."""

    def __init__(self, n, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.m = np.zeros(n, np.float32)
        self.v = np.zeros(n, np.float32)
        self.t = 0
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, epsilon

    def step(self, p, g, lr_mult=1.0):
        self.t += 1
        a_t = np.float32(self.lr * lr_mult * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        self.m = np.float32(self.b1) * self.m + np.float32(1 - self.b1) * g
        self.v = np.float32(self.b2) * self.v + np.float32(1 - self.b2) * g * g
        return p - a_t * self.m / (np.sqrt(self.v) + np.float32(self.eps))


class RMSProp:
    """update_methods_stats.py:11-32 (A2C: lr 7e-4, rho 0.9, eps 1e-6)."""

    def __init__(self, n, lr=7e-4, rho=0.9, epsilon=1e-6):
        self.v = np.zeros(n, np.float32)
        self.lr, self.rho, self.eps = lr, rho, epsilon

    def step(self, p, g, lr_mult=1.0):
        self.v = np.float32(self.rho) * self.v + np.float32(1 - self.rho) * g * g
        return p - np.float32(self.lr * lr_mult) * g / np.sqrt(self.v + np.float32(self.eps))


def iterate_mb_idxs(batch_size, data_length, rng, shuffle=True):
    """optimizers/util.py:8-18 (tail dropped; shuffle draws from the global legacy RandomState)."""
    indices = np.arange(data_length)
    if shuffle:
        rng.shuffle(indices)
    for start in range(0, data_length - batch_size + 1, batch_size):
        yield indices[start:start + batch_size]


def ppo_optimize(flat, opt, data, spec, n_actions, rng, epochs=4, minibatch_size=512, grad_norm_clip=None,
                 lr_mult=1.0, emulate_bf16=False, n_ranks_grads=None, **loss_kw):
    """optimizers/single/ppo_optimizer.py:57-76 -> (new flat, losses, grad_norms).
    data = (obs, act, adv, ret, old_value, old_prob)."""
    obs, act, adv, ret, _, old_prob = data
    losses_, norms = [], []
    for _ in range(epochs):
        for idx in iterate_mb_idxs(minibatch_size, len(obs), rng):
            loss, g, _ = loss_and_grad(flat, obs[idx], act[idx], adv[idx], ret[idx], old_prob[idx], spec,
                                       n_actions, "ppo", emulate_bf16, lr_mult=lr_mult, **loss_kw)
            g, norm = total_norm_clip(g.astype(np.float32), grad_norm_clip)
            flat = opt.step(flat, g, lr_mult)
            losses_.append(loss)
            norms.append(norm)
    return flat, losses_, norms
