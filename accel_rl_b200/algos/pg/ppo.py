"""PPO (reference: accel_rl/algos/pg/ppo.py:11-76).  pi_loss (ppo.py:42-51) is evaluated in
head_kernel<1> (csrc/kernels.cuh).  `num_slices` is accepted and ignored (the reference passes it to
an optimizer that has no such argument)."""
from accel_rl_b200.algos.pg.aac_base import AdvActorCriticBase
from accel_rl_b200.optimizers.single.ppo_optimizer import PpoOptimizer
from accel_rl_b200.optimizers.sync.sync_ppo_optimizer import SyncPpoOptimizer
from accel_rl_b200.optimizers.async_.async_ppo_optimizer import AsyncPpoOptimizer
from accel_rl_b200.optimizers import update_methods


class BasePPO(AdvActorCriticBase):
    loss_kind = "ppo"

    def __init__(self, OptimizerCls, optimizer_args=None, discount=0.99, gae_lambda=0.95, clip_param=0.2, **kwargs):
        default_optimizer_args = dict(
            num_slices=1,
            learning_rate=1e-3,
            epochs=4,
            minibatch_size=64 * 8,
            update_method=update_methods.adam,
            update_method_args=dict(epsilon=1e-5),
            grad_norm_clip=None,
            shuffle=True,
        )
        if optimizer_args is None:
            optimizer_args = default_optimizer_args
        else:
            for k, v in default_optimizer_args.items():
                optimizer_args.setdefault(k, v)
        self.optimizer = OptimizerCls(**optimizer_args)
        self.clip_param = clip_param
        super().__init__(discount=discount, gae_lambda=gae_lambda, **kwargs)


class PPO(BasePPO):
    """Single GPU"""

    def __init__(self, OptimizerCls=PpoOptimizer, **kwargs):
        super().__init__(OptimizerCls=OptimizerCls, **kwargs)


class mPPO(BasePPO):
    """Multi-GPU Synchronous"""

    def __init__(self, OptimizerCls=SyncPpoOptimizer, **kwargs):
        super().__init__(OptimizerCls=OptimizerCls, **kwargs)


class mAPPO(BasePPO):
    """Multi-GPU Asynchronous"""

    def __init__(self, OptimizerCls=AsyncPpoOptimizer, **kwargs):
        super().__init__(OptimizerCls=OptimizerCls, **kwargs)
