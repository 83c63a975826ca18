"""SyncA2cOptimizer (reference: accel_rl/optimizers/sync/sync_a2c_optimizer.py:13-59)."""
import torch

from accel_rl_b200.optimizers.single.a2c_optimizer import A2cOptimizer
from accel_rl_b200.optimizers.sync.base import BaseSyncOptimizer


class SyncA2cOptimizer(BaseSyncOptimizer, A2cOptimizer):
    def optimize(self, inputs):
        n = self._bind_inputs(inputs)
        if self._idx_dev is None or self._idx_dev.numel() != n:
            self._idx_dev = torch.arange(n, dtype=torch.int32, device=self._engine.device)
        self._engine.grad_minibatch(self._idx_dev, n)
        self._engine.sync_allreduce_update()
        losses, grad_norms = self._engine.read_logs()
        return float(losses[0]), float(grad_norms[0])
