"""AsyncPpoOptimizer (reference: accel_rl/optimizers/async/async_ppo_optimizer.py — a SyntaxError upstream,
SURVEY.md F6-v; the evident intent is implemented: PpoOptimizer's epoch/minibatch loop with every minibatch
gradient pushed into the central Adam state under the chunk locks and the new parameters pulled)."""
from accel_rl_b200.optimizers.async_.base import BaseAsyncOptimizer
from accel_rl_b200.optimizers.single.ppo_optimizer import PpoOptimizer


class AsyncPpoOptimizer(BaseAsyncOptimizer, PpoOptimizer):
    def __init__(self, learning_rate, epochs, minibatch_size, update_method_name="adam", update_method_args=None,
                 n_update_chunks=1, grad_norm_clip=None, shuffle=True, update_method=None, num_slices=1):
        self._check_update_name(update_method_name, n_update_chunks)
        if update_method_args is None and update_method is not None:
            update_method_args = dict()
        PpoOptimizer.__init__(self, learning_rate=learning_rate, update_method=update_method_name,
                              update_method_args=update_method_args or dict(), epochs=epochs, minibatch_size=minibatch_size,
                              grad_norm_clip=grad_norm_clip, shuffle=shuffle)
        self._update_method_name = update_method_name
        self.n_update_chunks = n_update_chunks

    def initialize(self, inputs, losses, constraints, target, givens=None, lr_mult=1):
        self._configure_async(losses, target, lr_mult)

    def _do_updates(self, data_length):
        n_mb = self._upload_indices(data_length)
        # per minibatch: local gradient -> chunk-locked push into the central state -> pull (one CUDA graph each)
        self._engine.train_minibatches(self._idx_dev, self._minibatch_size, n_mb * self._epochs, sync="async")
        losses, grad_norms = self._engine.read_logs()
        return list(losses), list(grad_norms)

    @property
    def parallelism_tag(self):
        return "asynchronous"
