"""reference: accel_rl/optimizers/async/base.py:11-104, chunked_updates.py:53-120.

The reference keeps the central (params, m, v) in host shared memory guarded by one mp.Lock per chunk and moves
every chunk over PCIe both ways.  Here the central store lives in rank 0's HBM (CUDA IPC), the chunk locks are
system-scope CAS words next to it, and push + pull are ONE kernel per update (csrc/comm.cuh
async_push_pull_kernel): local global-norm clip -> per lock region {lock, Adam/RMSProp on central (p, m, v) with the
local gradient, write the new p to the central store and to the local parameters, unlock}.

init_comm(rank, n_runners, par_objs): `par_objs` is the runner's dict; its "exchange" entry is a callable
exchange(bytes) -> [bytes per rank] (torch.distributed all_gather of the 64-byte IPC handle) standing in for the
reference's Manager dict + barrier (base.py:21-41)."""
from accel_rl_b200.optimizers.base import BaseOptimizer
from accel_rl_b200.optimizers import update_methods

CHUNKED_UPDATE_NAMES = ["rmsprop", "adam"]      # chunked_updates.py:6
WHOLE_UPDATE_NAMES = CHUNKED_UPDATE_NAMES + []  # chunked_updates.py:7


class BaseAsyncOptimizer(BaseOptimizer):
    def _check_update_name(self, update_method_name, n_update_chunks):
        # same tests and messages as async_a2c_optimizer.py:26-32
        if n_update_chunks == 1 and update_method_name not in WHOLE_UPDATE_NAMES:
            raise ValueError("update method '{}' not available for NON-chunked "
                             "updates, choose from: {}".format(update_method_name, WHOLE_UPDATE_NAMES))
        elif update_method_name not in CHUNKED_UPDATE_NAMES:
            raise ValueError("update method '{}' not available, for CHUNKED "
                             "updates, choose from: {}".format(update_method_name, CHUNKED_UPDATE_NAMES))

    def _configure_async(self, losses, target, lr_mult):
        # `update_method_name` (a string) + `update_method_args` replace the single-GPU `update_method` descriptor
        self._update_method = update_methods.BY_NAME[self._update_method_name]
        self._configure_engine(losses, target, lr_mult)

    def init_comm(self, rank, n_runners, par_objs):
        self._rank = rank
        self._n_runners = n_runners
        exchange = par_objs["exchange"] if isinstance(par_objs, dict) else par_objs
        self.n_lock_regions = self._engine.async_init(rank, n_runners, self.n_update_chunks, exchange)

    @property
    def parallelism_tag(self):
        return "asynchronous"

    @property
    def central_shared_params(self):
        """host copy of the central parameter vector (reference: the shared-memory array itself, base.py:48-50)"""
        return self._engine.async_read_central(0)

    @property
    def central_params_handle(self):
        """what ActsrvAltOvrlpPollSampler.poll_init takes as `central_shared_params`: pulls the central parameters into
        this learner's policy on the device (poll_sampler.py:29-39)"""
        from accel_rl_b200.sampler.poll_sampler import CentralParams
        return CentralParams(self._engine)

    @property
    def central_params_lock(self):
        raise NotImplementedError("the chunk locks live on the device (csrc/comm.cuh); read central_shared_params instead")
