"""AsyncA2cOptimizer (reference: accel_rl/optimizers/async/async_a2c_optimizer.py:15-109): per optimize()
one full-batch local gradient (clipped locally) pushed into the central RMSProp/Adam state, then the new central
parameters pulled."""
import torch

from accel_rl_b200.optimizers.async_.base import BaseAsyncOptimizer


class AsyncA2cOptimizer(BaseAsyncOptimizer):
    def __init__(self, learning_rate, update_method_name, update_method_args=None, n_update_chunks=1,
                 grad_norm_clip=None, update_method=None):
        self._learning_rate = learning_rate
        self._grad_norm_clip = grad_norm_clip
        self._check_update_name(update_method_name, n_update_chunks)
        self._update_method_name = update_method_name
        self._update_method_args = update_method_args or dict()
        self.n_update_chunks = n_update_chunks
        self._idx_dev = None

    def initialize(self, inputs, losses, constraints, target, givens=None, lr_mult=1):
        self._configure_async(losses, target, lr_mult)

    def optimize(self, inputs):
        n = self._bind_inputs(inputs)
        if self._idx_dev is None or self._idx_dev.numel() != n:
            self._idx_dev = torch.arange(n, dtype=torch.int32, device=self._engine.device)
        self._engine.grad_minibatch(self._idx_dev, n)      # _compute_grad  (async_a2c_optimizer.py:103-109)
        self._engine.async_push_pull()                     # _push_update + _f_copy (:98-100)
        losses, grad_norms = self._engine.read_logs()
        return float(losses[0]), float(grad_norms[0])
