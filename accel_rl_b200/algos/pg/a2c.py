"""A2C and its multi-GPU variants (reference: accel_rl/algos/pg/a2c.py:14-97).

The policy-gradient term  -mean(log(pi(a|s) + 1e-8) * A)  (a2c.py:43-46) is evaluated by head_kernel<1> in
csrc/kernels.cuh together with the value and entropy terms; this module only fixes hyper-parameter defaults and picks
the optimizer class.  Defaults (a2c.py:16-29): discount 0.99, gae_lambda 1 (plain discounted returns),
v_loss_coeff 0.25, RMSProp lr 7e-4 with global-norm clip 0.5.
"""
from accel_rl_b200.algos.pg.aac_base import AdvActorCriticBase, with_defaults
from accel_rl_b200.optimizers import update_methods
from accel_rl_b200.optimizers.async_.async_a2c_optimizer import AsyncA2cOptimizer
from accel_rl_b200.optimizers.single.a2c_optimizer import A2cOptimizer
from accel_rl_b200.optimizers.sync.sync_a2c_optimizer import SyncA2cOptimizer


class BaseA2C(AdvActorCriticBase):
    loss_kind = "a2c"
    default_optimizer = None
    optimizer_defaults = dict(learning_rate=7e-4, update_method=update_methods.rmsprop, update_method_args=dict(),
                              grad_norm_clip=0.5)

    def __init__(self, OptimizerCls=None, optimizer_args=None, discount=0.99, gae_lambda=1, v_loss_coeff=0.25, **kwargs):
        cls = OptimizerCls if OptimizerCls is not None else self.default_optimizer
        if cls is None:
            raise TypeError("BaseA2C needs an OptimizerCls (use A2C, mA2C or mA3C)")
        self.optimizer = cls(**with_defaults(optimizer_args, self.optimizer_defaults))
        super().__init__(discount=discount, gae_lambda=gae_lambda, v_loss_coeff=v_loss_coeff, **kwargs)


class A2C(BaseA2C):
    """one GPU"""
    default_optimizer = A2cOptimizer


class mA2C(BaseA2C):
    """synchronous data parallel: gradients averaged over the GPUs every step"""
    default_optimizer = SyncA2cOptimizer


class mA3C(BaseA2C):
    """asynchronous data parallel: every learner pushes its gradient into the central optimizer state"""
    default_optimizer = AsyncA2cOptimizer
    optimizer_defaults = dict(BaseA2C.optimizer_defaults, update_method_name="rmsprop", n_update_chunks=3)
