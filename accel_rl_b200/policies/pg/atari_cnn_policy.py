"""AtariCnnPolicy (reference: accel_rl/policies/pg/atari_cnn_policy.py:15-119, network
policies/pg/networks/pg_cnn.py:45-86, inits policies/layers.py:11-19).

Same constructor, same public methods; the Theano functions `_f_prob`, `_f_value`, `_f_prob_value`
(atari_cnn_policy.py:65-67) are one tcgen05 forward pass in libaccelrl_b200 (arl_policy_forward).
Parameters are a flat fp32 device vector in the reference's order (rllab/core/parameterized.py:74-88):
conv{i}.W (out,in,kh,kw), conv{i}.b, hidden_0.W (in,out), hidden_0.b, output_pi.W, output_pi.b,
output_v.W, output_v.b — so snapshots interchange with the reference.
"""
import math

import numpy as np
import torch

from accel_rl_b200.distributions.categorical import Categorical
from accel_rl_b200.engine import Engine
from accel_rl_b200.spaces.discrete import Discrete
from accel_rl_b200.util import seeding
from accel_rl_b200.util.quick_args import save_args, retrieve_args


def rectify(x):  # placeholders for the reference's lasagne.nonlinearities arguments
    return x


def softmax(x):
    return x


class AtariCnnPolicy(object):
    def __init__(self, conv_filters, conv_filter_sizes, conv_strides, conv_pads, hidden_sizes=[],
                 hidden_nonlinearity=rectify, output_pi_nonlinearity=softmax, pixel_scale=255.,
                 initial_param_values=None, max_rows=None):
        save_args(vars(), underscore=True)
        self.initial_param_values = initial_param_values
        self._engine = None
        self._reserve = int(max_rows) if max_rows else 0
        self._host_params = None

    # ------------------------------------------------------------------ initialisation -------
    def initialize(self, env_spec, **kwargs):
        assert isinstance(env_spec.action_space, Discrete)
        s = retrieve_args(self)
        if s.hidden_nonlinearity is not rectify or s.output_pi_nonlinearity is not softmax:
            raise NotImplementedError("CUDA path implements ReLU hidden units and a softmax pi head")
        self._env_spec = env_spec
        self._dist = Categorical(env_spec.action_space.n)
        self._shapes = self._param_shapes(env_spec)
        self._names = self._param_names()
        self.param_short_names = [n.replace("atari_cnn_", "").replace("conv_hidden", "conv").replace("hidden", "fc")
                                  for n in self._names]
        self._host_params = self._init_param_values()
        if self.initial_param_values is not None:
            self._host_params = np.asarray(self.initial_param_values, dtype=np.float32).copy()

    def _param_shapes(self, env_spec):
        c, h, w = env_spec.observation_space.shape
        shapes = []
        for f, k, st, p in zip(self._conv_filters, self._conv_filter_sizes, self._conv_strides, self._conv_pads):
            p = p[0] if isinstance(p, (tuple, list)) else p
            shapes += [(f, c, k, k), (f,)]
            h, w, c = (h + 2 * p - k) // st + 1, (w + 2 * p - k) // st + 1, f
        n_in = c * h * w
        for hs in self._hidden_sizes:
            shapes += [(n_in, hs), (hs,)]
            n_in = hs
        a = env_spec.action_space.n
        return shapes + [(n_in, a), (a,), (n_in, 1), (1,)]

    def _param_names(self):
        names = []
        for i in range(len(self._conv_filters)):
            names += ["atari_cnn_conv_hidden_%d.W" % i, "atari_cnn_conv_hidden_%d.b" % i]
        for i in range(len(self._hidden_sizes)):
            names += ["atari_cnn_hidden_%d.W" % i, "atari_cnn_hidden_%d.b" % i]
        return names + ["atari_cnn_output_pi.W", "atari_cnn_output_pi.b", "atari_cnn_output_v.W", "atari_cnn_output_v.b"]

    def _init_param_values(self):
        """pg_cnn.py:25-29: conv W GlorotUniform (lasagne rng), dense W NormCInit(1.0 / 0.01 / 1.0) from the
        global np.random stream (layers.py:16-19), biases 0 — drawn in layer-construction order."""
        conv_rng = seeding.get_conv_init_rng()
        out = []
        n_conv, n_hid = len(self._conv_filters), len(self._hidden_sizes)
        for i, shp in enumerate(self._shapes):
            if len(shp) == 4:
                f, c, kh, kw = shp
                lim = math.sqrt(6.0 / ((c + f) * kh * kw))
                out.append(np.asarray(conv_rng.uniform(low=-lim, high=lim, size=shp), dtype=np.float32))
            elif len(shp) == 2:
                layer = (i - 2 * n_conv) // 2
                std = 0.01 if layer == n_hid else 1.0
                w = np.random.randn(*shp).astype(np.float32)
                w *= std / np.sqrt(np.square(w).sum(axis=0, keepdims=True))
                out.append(w)
            else:
                out.append(np.zeros(shp, np.float32))
        return np.concatenate([a.ravel() for a in out])

    # ------------------------------------------------------------------ engine ---------------
    def reserve(self, rows):
        """Tell the policy the largest batch it will see (sampler envs, training minibatch).  Only BEFORE the engine
        exists does this size the workspaces; a live engine is never re-created (the optimizer, the sampler and the
        multi-GPU learners hold state inside it) — larger inference batches are simply served in slices of
        engine.max_rows (see _forward / get_actions)."""
        rows = int(rows)
        if self._engine is None and rows > self._reserve:
            self._reserve = rows

    @property
    def engine(self):
        if self._engine is None:
            if self._host_params is None:
                raise RuntimeError("policy.initialize(env_spec) must be called first")
            pads = self._conv_pads
            self._engine = Engine(self._conv_filters, self._conv_filter_sizes, self._conv_strides, pads,
                                  self._hidden_sizes, self._env_spec.action_space.n,
                                  self._env_spec.observation_space.shape, self._pixel_scale,
                                  max_rows=max(self._reserve, 1))
            self._engine.set_params(self._host_params)
        return self._engine

    # ------------------------------------------------------------------ inference ------------
    def _forward(self, observations, want_prob=True, want_value=True, uniforms=None, actions=None):
        eng = self.engine
        is_np = not torch.is_tensor(observations)
        obs = torch.as_tensor(np.ascontiguousarray(observations)) if is_np else observations
        obs = obs.to(eng.device, non_blocking=True).contiguous()
        n = obs.shape[0]
        prob = torch.empty((n, eng.n_actions), dtype=torch.float32, device=eng.device) if want_prob else None
        value = torch.empty((n,), dtype=torch.float32, device=eng.device) if want_value else None
        # batches larger than the engine's workspaces (e.g. a diagnostic dist_info over the whole rollout) go through in
        # slices: rows are independent, so the result is the same
        for lo in range(0, n, eng.max_rows):
            hi = min(n, lo + eng.max_rows)
            eng.forward(obs[lo:hi], n=hi - lo, prob=None if prob is None else prob[lo:hi],
                        value=None if value is None else value[lo:hi],
                        uniforms=None if uniforms is None else uniforms[lo:hi],
                        actions=None if actions is None else actions[lo:hi])
        if is_np:
            prob = prob.cpu().numpy() if prob is not None else None
            value = value.cpu().numpy() if value is not None else None
        return prob, value

    def dist_info(self, observations, state_infos=None):
        return dict(prob=self._forward(observations, True, False)[0])

    def value(self, observations, state_infos=None):
        return self._forward(observations, False, True)[1]

    def dist_info_value(self, observations, state_infos=None):
        prob, value = self._forward(observations)
        return dict(prob=prob, value=value)

    def get_action(self, observation, deterministic=False):
        probs, values = self._forward(np.asarray(observation)[None])
        prob, value = probs[0], values[0]
        if deterministic:
            action = np.argmax(prob)
        else:
            action = self.action_space.weighted_sample(prob)   # one np.random.rand() (special.py:18)
        return action, dict(prob=prob, value=value)

    def get_actions(self, observations):
        """atari_cnn_policy.py:108-111.  Host arrays in -> host arrays out; the uniforms are drawn from the
        global legacy stream exactly like weighted_sample_n, the comparison runs on the device."""
        eng = self.engine
        is_np = not torch.is_tensor(observations)
        n = len(observations)
        u = torch.from_numpy(np.random.rand(n)).to(eng.device)
        act = torch.empty((n,), dtype=torch.uint8, device=eng.device)
        prob, value = self._forward(observations, uniforms=u, actions=act)
        if is_np:
            return act.cpu().numpy(), dict(prob=prob, value=value)
        return act, dict(prob=prob, value=value)

    def reset(self, n_batch=None):
        pass

    def reset_one(self, idx):
        pass

    # ------------------------------------------------------------------ properties -----------
    vectorized = property(lambda self: True)
    recurrent = property(lambda self: False)
    state_info_keys = property(lambda self: [])
    distribution = property(lambda self: self._dist)
    action_space = property(lambda self: self._env_spec.action_space)
    observation_space = property(lambda self: self._env_spec.observation_space)

    # ------------------------------------------------------------------ parameters -----------
    def get_params(self, **tags):
        """list of (name, offset, shape) views into the flat vector, Lasagne order"""
        out, i = [], 0
        for name, shp in zip(self._names, self._shapes):
            n = int(np.prod(shp))
            out.append((name, i, shp))
            i += n
        return out

    def get_param_shapes(self, **tags):
        return list(self._shapes)

    def get_param_values(self, **tags):
        if self._engine is not None:
            return self._engine.get_params()
        return self._host_params.copy()

    def set_param_values(self, flattened_params, **tags):
        flat = np.asarray(flattened_params, dtype=np.float32).reshape(-1)
        if self._engine is not None:
            self._engine.set_params(flat)
        else:
            self._host_params = flat.copy()

    def flat_to_params(self, flattened_params, **tags):
        out, i = [], 0
        for shp in self._shapes:
            n = int(np.prod(shp))
            out.append(np.asarray(flattened_params[i:i + n]).reshape(shp))
            i += n
        return out
