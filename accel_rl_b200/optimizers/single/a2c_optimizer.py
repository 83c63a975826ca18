"""A2cOptimizer (reference: accel_rl/optimizers/single/a2c_optimizer.py:11-50): one full-batch
gradient + clip + RMSProp step per optimize() call."""
import torch

from accel_rl_b200.optimizers.base import BaseOptimizer


class A2cOptimizer(BaseOptimizer):
    def __init__(self, learning_rate, update_method, update_method_args=None, grad_norm_clip=None):
        self._learning_rate = learning_rate
        self._update_method = update_method
        self._update_method_args = update_method_args or dict()
        self._grad_norm_clip = grad_norm_clip
        self._idx_dev = None

    def initialize(self, inputs, losses, constraints, target, givens=None, lr_mult=1):
        self._configure_engine(losses, target, lr_mult)

    def optimize(self, inputs):
        n = self._bind_inputs(inputs)
        if self._idx_dev is None or self._idx_dev.numel() != n:
            self._idx_dev = torch.arange(n, dtype=torch.int32, device=self._engine.device)
        self._engine.train_minibatches(self._idx_dev, n, 1)
        losses, grad_norms = self._engine.read_logs()
        return float(losses[0]), float(grad_norms[0])

    @property
    def parallelism_tag(self):
        return "single"
