"""SyncPpoOptimizer (reference: accel_rl/optimizers/sync/sync_ppo_optimizer.py:13-78).  The reference
defines no __init__; the evident intent — same constructor as PpoOptimizer — is
implemented.  Per minibatch: local gradient -> fused P2P all-reduce + average + clip + Adam."""
from accel_rl_b200.optimizers.single.ppo_optimizer import PpoOptimizer
from accel_rl_b200.optimizers.sync.base import BaseSyncOptimizer


class SyncPpoOptimizer(BaseSyncOptimizer, PpoOptimizer):
    def _do_updates(self, data_length):
        n_mb = self._upload_indices(data_length)
        # per minibatch: _compute_grad, then _share_grad + _do_one_update fused in one kernel — replayed as ONE CUDA
        # graph per minibatch (every rank replays the same number of graphs, so the cross-GPU barriers line up)
        self._engine.train_minibatches(self._idx_dev, self._minibatch_size, n_mb * self._epochs, sync=True)
        losses, grad_norms = self._engine.read_logs()
        return list(losses), list(grad_norms)
