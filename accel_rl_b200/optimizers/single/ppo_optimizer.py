"""PpoOptimizer (reference: accel_rl/optimizers/single/ppo_optimizer.py:11-76).

`_f_load` (the H2D copy of the whole rollout, ppo_optimizer.py:44,63) does not exist: the rollout is
already in HBM.  `_f_opt` per shuffled minibatch — gather rows by index, forward, losses, backward,
global-norm clip, Adam (ppo_optimizer.py:49-76, optimizers/util.py:70-76) — is one CUDA graph
(tcgen05 conv/FC tiles + head/loss + wgrad/dgrad + fused clip/update) replayed epochs*N/mb times
with the index block uploaded once per optimize() call."""
import numpy as np
import torch

from accel_rl_b200.optimizers.base import BaseOptimizer
from accel_rl_b200.optimizers.util import epoch_index_block


class PpoOptimizer(BaseOptimizer):
    def __init__(self, learning_rate, update_method, update_method_args, epochs, minibatch_size,
                 grad_norm_clip=None, shuffle=True, num_slices=1):
        self._learning_rate = learning_rate
        self._update_method = update_method
        self._update_method_args = update_method_args
        self._epochs = epochs
        self._minibatch_size = minibatch_size
        self._shuffle = shuffle
        self._grad_norm_clip = grad_norm_clip
        self._idx_dev = None

    def max_rows(self, sample_size):
        return min(self._minibatch_size, sample_size)

    def initialize(self, inputs, losses, constraints, target, givens=None, lr_mult=1):
        self._configure_engine(losses, target, lr_mult)

    def optimize(self, inputs):
        data_length = self._bind_inputs(inputs)
        return self._do_updates(data_length)

    def _upload_indices(self, data_length):
        idx, n_mb = epoch_index_block(self._minibatch_size, data_length, self._epochs, self._shuffle)
        if self._idx_dev is None or self._idx_dev.numel() != idx.size:
            self._idx_dev = torch.zeros(idx.size, dtype=torch.int32, device=self._engine.device)
            self._idx_host = torch.empty(idx.size, dtype=torch.int32).pin_memory()
        self._idx_host.copy_(torch.from_numpy(idx))
        self._idx_dev.copy_(self._idx_host, non_blocking=True)
        return n_mb

    def _do_updates(self, data_length):
        n_mb = self._upload_indices(data_length)
        count = n_mb * self._epochs
        self._engine.train_minibatches(self._idx_dev, self._minibatch_size, count)
        losses, grad_norms = self._engine.read_logs()
        return list(losses), list(grad_norms)

    @property
    def parallelism_tag(self):
        return "single"
