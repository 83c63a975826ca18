"""PPO and its multi-GPU variants (reference: accel_rl/algos/pg/ppo.py:11-76).

The clipped surrogate  -mean(min(r*A, clip(r, 1-eps, 1+eps)*A)),  r = (pi_new(a)+1e-8)/(pi_old(a)+1e-8),
eps = clip_param * lr_mult (ppo.py:42-51) is evaluated by head_kernel<1> in csrc/kernels.cuh.  Defaults
(ppo.py:14-31): discount 0.99, gae_lambda 0.95, clip_param 0.2, 4 epochs of 512-row shuffled minibatches, Adam lr 1e-3
with epsilon 1e-5, no gradient-norm clipping.  `num_slices` is accepted and ignored: the reference passes it to an
optimizer that has no such argument.
"""
from accel_rl_b200.algos.pg.aac_base import AdvActorCriticBase, with_defaults
from accel_rl_b200.optimizers import update_methods
from accel_rl_b200.optimizers.async_.async_ppo_optimizer import AsyncPpoOptimizer
from accel_rl_b200.optimizers.single.ppo_optimizer import PpoOptimizer
from accel_rl_b200.optimizers.sync.sync_ppo_optimizer import SyncPpoOptimizer


class BasePPO(AdvActorCriticBase):
    loss_kind = "ppo"
    default_optimizer = None
    optimizer_defaults = dict(num_slices=1, learning_rate=1e-3, epochs=4, minibatch_size=512,
                              update_method=update_methods.adam, update_method_args=dict(epsilon=1e-5),
                              grad_norm_clip=None, shuffle=True)

    def __init__(self, OptimizerCls=None, optimizer_args=None, discount=0.99, gae_lambda=0.95, clip_param=0.2, tie_grad=1,
                 **kwargs):
        # tie_grad (not a reference argument): inside the clip range the two surrogates of ppo.py:47-49 are exactly equal and
        # T.minimum's gradient at a tie depends on the Theano version — 1: passed once (Theano >= 0.9, and standard PPO);
        # 2: to both branches, i.e. twice the policy gradient there (older Theano)
        cls = OptimizerCls if OptimizerCls is not None else self.default_optimizer
        if cls is None:
            raise TypeError("BasePPO needs an OptimizerCls (use PPO, mPPO or mAPPO)")
        if tie_grad not in (1, 2):
            raise ValueError("tie_grad must be 1 or 2")
        self.clip_param = clip_param
        self.tie_grad = tie_grad
        self.optimizer = cls(**with_defaults(optimizer_args, self.optimizer_defaults))
        super().__init__(discount=discount, gae_lambda=gae_lambda, **kwargs)


class PPO(BasePPO):
    """one GPU"""
    default_optimizer = PpoOptimizer


class mPPO(BasePPO):
    """synchronous data parallel"""
    default_optimizer = SyncPpoOptimizer


class mAPPO(BasePPO):
    """asynchronous data parallel"""
    default_optimizer = AsyncPpoOptimizer
