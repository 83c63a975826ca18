"""Advantage actor-critic base (reference: accel_rl/algos/pg/aac_base.py:15-176).

initialize() — where the reference builds the Theano loss graph (aac_base.py:32-97) — hands the
optimizer a declarative loss spec; the loss, its gradient and the update are CUDA kernels.
process_samples() — bootstrap value + per-env GAE / discounted returns (aac_base.py:108-145,
algos/pg/util.py:6-63) — is one value-only forward pass and one warp-scan kernel over the rollout
buffers already resident in HBM.
"""
import torch

from accel_rl_b200.algos.base import RLAlgorithm
from accel_rl_b200.buffers.batch import buffer_with_segs_view
from accel_rl_b200.util.quick_args import save_args
import numpy as np

LR_SCHEDULES = ["linear"]


def with_defaults(optimizer_args, defaults):
    """the user's optimizer_args win; anything they leave out comes from the algorithm's defaults
    (the reference fills the caller's dict in place, a2c.py:30-36 / ppo.py:32-37; a copy is returned here)"""
    merged = dict(defaults)
    merged.update(optimizer_args or {})
    return merged


class AdvActorCriticBase(RLAlgorithm):
    def __init__(self, discount, gae_lambda, v_loss_coeff=1, ent_loss_coeff=0.01, standardize_adv=False,
                 lr_schedule=None):
        if lr_schedule is not None and lr_schedule not in LR_SCHEDULES:
            raise ValueError("Unrecognized lr_schedule: {}, should be None (for constant) or in: {}".format(
                lr_schedule, LR_SCHEDULES))
        save_args(vars(), underscore=False)
        self.need_extra_obs = True  # (signal sent to the sampler)

    def initialize(self, policy, env_spec, sample_size, horizon, mid_batch_reset):
        if mid_batch_reset and policy.recurrent:
            raise NotImplementedError
        self.policy = policy
        self._use_valids = not (mid_batch_reset and not policy.recurrent)   # aac_base.py:53-58
        self._dist_info_keys = policy.distribution.dist_info_keys
        self._state_info_keys = policy.state_info_keys
        self._batch_size = sample_size
        self._mid_batch_reset = mid_batch_reset
        self._horizon = horizon
        self._lr_mult = 1.0
        policy.reserve(self.optimizer.max_rows(sample_size))
        policy.reserve(sample_size // horizon)
        eng = policy.engine
        dev = eng.device
        opt_examples = dict(advantages=np.float32(1), returns=np.float32(1))
        if self._use_valids:
            opt_examples["valids"] = np.int8(1)
        self._opt_buf = buffer_with_segs_view(opt_examples, sample_size, horizon, dev)
        self._last_values = torch.zeros(sample_size // horizon, dtype=torch.float32, device=dev)
        self.optimizer.initialize(
            inputs=None,
            losses=dict(kind=self.loss_kind, v_loss_coeff=self.v_loss_coeff, ent_loss_coeff=self.ent_loss_coeff,
                        clip_param=getattr(self, "clip_param", 0.), tie_grad=getattr(self, "tie_grad", 1)),
            constraints=None,
            target=policy,
            lr_mult=self._lr_mult,
        )

    def set_n_itr(self, n_itr):
        self.n_itr = n_itr

    # AccelRLEval calls these around sampler.evaluate_policy (runners/accel_rl.py:137-139); in the reference only DQN
    # defines them (algos/dqn/dqn.py:209-216: swap in the eval epsilon).  A policy-gradient policy samples from its own
    # distribution in evaluation too, so there is nothing to switch.
    def prep_eval(self, itr):
        pass

    def post_eval(self, itr):
        pass

    def optimize_policy(self, itr, samples_data):
        opt_data = self.process_samples(itr, samples_data)
        opt_input_values = self.prep_opt_inputs(itr, samples_data, opt_data)
        _, grad_norm = self.optimizer.optimize(opt_input_values)
        return opt_data, dict(GradNorm=grad_norm)

    def process_samples(self, itr, samples_data):
        eng = self.policy.engine
        B = self._batch_size // self._horizon
        eng.forward(samples_data["extra_observations"], n=B, value=self._last_values)   # aac_base.py:112
        opt = self._opt_buf
        env_infos = samples_data["env_infos"]
        need_reset = env_infos.get("need_reset", samples_data["dones"]) if self._use_valids else None
        eng.gae(samples_data["rewards"], samples_data["agent_infos"]["value"], samples_data["dones"], need_reset,
                self._last_values, self.discount, self.gae_lambda, opt["advantages"], opt["returns"],
                opt.get("valids"), B, self._horizon, self.standardize_adv)
        return opt

    def prep_opt_inputs(self, itr, samples_data, opt_data):
        agent_infos = samples_data["agent_infos"]
        opt_input_values = (samples_data["observations"], samples_data["actions"], opt_data["advantages"],
                            opt_data["returns"], agent_infos["value"])
        opt_input_values += tuple(agent_infos[k] for k in self._dist_info_keys)
        if self._use_valids:
            opt_input_values += (opt_data["valids"],)
        if self.lr_schedule == "linear":
            self._lr_mult = max((self.n_itr - itr) / self.n_itr, 0.)     # aac_base.py:165-168
            self.optimizer.set_lr_mult(self._lr_mult)
        return opt_input_values

    def constraint_values(self, samples_data, opt_data=None):
        """(pi_kl, v_kl) of aac_base.py:68-70 — mean KL(old policy || current policy) and mean squared change of the value
        over the rollout rows (valids-weighted when the validity mask is in use).  The reference hands these two
        expressions to its optimizers, none of which compiles them (optimizers/single/ppo_optimizer.py:30-58 ignores
        `constraints`); here they are a diagnostic evaluated on demand: one forward pass of the current parameters over
        the rollout's observations, the two means taken where the rows are (HBM)."""
        agent_infos = samples_data["agent_infos"]
        new = self.policy.dist_info_value(samples_data["observations"])
        p, q = agent_infos["prob"], new["prob"]
        tiny = 1e-8                                                   # distributions/categorical.py:6
        kl = (p * (torch.log(p + tiny) - torch.log(q + tiny))).sum(dim=-1)
        dv = (new["value"] - agent_infos["value"]) ** 2
        if self._use_valids:
            valids = (self._opt_buf if opt_data is None else opt_data)["valids"].to(kl.dtype)
            n = valids.sum()
            return float((kl * valids).sum() / n), float((dv * valids).sum() / n)
        return float(kl.mean()), float(dv.mean())

    @property
    def opt_info_keys(self):
        return ["GradNorm"]
