"""SyncPpoOptimizer (reference: accel_rl/optimizers/sync/sync_ppo_optimizer.py:13-78).  The reference
defines no __init__; the evident intent — same constructor as PpoOptimizer — is
implemented.  Per minibatch: local gradient -> fused P2P all-reduce + average + clip + Adam."""
from accel_rl_b200.optimizers.single.ppo_optimizer import PpoOptimizer
from accel_rl_b200.optimizers.sync.base import BaseSyncOptimizer


class SyncPpoOptimizer(BaseSyncOptimizer, PpoOptimizer):
    def _do_updates(self, data_length):
        n_mb = self._upload_indices(data_length)
        eng, mb = self._engine, self._minibatch_size
        k = 0
        for _ in range(self._epochs):
            for _ in range(n_mb):
                eng.grad_minibatch(self._idx_dev[k * mb:(k + 1) * mb], mb)   # _compute_grad
                eng.sync_allreduce_update()                                  # _share_grad + _do_one_update
                k += 1
        losses, grad_norms = eng.read_logs()
        return list(losses), list(grad_norms)
