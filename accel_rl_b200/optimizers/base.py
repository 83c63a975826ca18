"""Optimizer interface (reference: accel_rl/optimizers/base.py:3-13).

`initialize(inputs, losses, constraints, target, givens, lr_mult)` keeps its name and arguments; the
reference passes Theano expressions, here `losses` is the declarative loss spec built by
AdvActorCriticBase.initialize and `target` is the policy (whose engine owns params/grads/state)."""
from accel_rl_b200.optimizers import update_methods


class BaseOptimizer(object):
    def initialize(self, inputs, losses, constraints, target, givens=None, lr_mult=1):
        raise NotImplementedError

    def optimize(self, inputs):
        raise NotImplementedError

    @property
    def parallelism_tag(self):
        raise NotImplementedError

    # ---- shared plumbing ------------------------------------------------------------------
    def max_rows(self, sample_size):
        """largest batch one gradient evaluation sees"""
        return sample_size

    def _configure_engine(self, losses, target, lr_mult):
        self._target = target
        self._engine = target.engine
        um = self._update_method
        if isinstance(um, str):
            um = update_methods.BY_NAME[um]
        args = um.resolve(**(self._update_method_args or {}))
        clip = self._grad_norm_clip
        self._engine.opt_configure(
            algo=0 if losses["kind"] == "ppo" else 1,
            clip_param=float(losses.get("clip_param", 0.)),
            v_loss_coeff=float(losses["v_loss_coeff"]),
            ent_loss_coeff=float(losses["ent_loss_coeff"]),
            update=um.kind,
            learning_rate=float(self._learning_rate),
            beta1=float(args.get("beta1", 0.9)), beta2=float(args.get("beta2", 0.999)),
            epsilon=float(args["epsilon"]), rho=float(args.get("rho", 0.9)),
            grad_norm_clip=float(clip) if clip is not None else -1.0,
            ppo_tie_grad=int(losses.get("tie_grad", 1)))
        self._engine.reset_opt_state()
        self.set_lr_mult(lr_mult)

    def set_lr_mult(self, lr_mult):
        self._engine.set_lr_mult(lr_mult)

    # ---- snapshot / resume (the reference's snapshot holds parameters only) ---------
    def get_state(self):
        return self._engine.get_opt_state()

    def set_state(self, state):
        self._engine.set_opt_state(state)

    def _bind_inputs(self, inputs):
        obs, act, adv, ret, old_value, old_prob = inputs[:6]
        valids = inputs[6] if len(inputs) > 6 else None
        self._engine.bind_train_inputs(obs, act, adv, ret, old_value, old_prob, valids)
        return int(obs.shape[0])
