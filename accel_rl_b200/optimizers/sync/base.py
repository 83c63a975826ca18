"""reference: accel_rl/optimizers/sync/base.py:8-24.

init_comm(gpu_comm, rank, n_gpu): `gpu_comm` is a callable exchange(bytes) -> [bytes per rank]
(torch.distributed all_gather of the 64-byte CUDA IPC handle); NCCL is only this bootstrap.  The
reference's `_share_grad` (in-place NCCL all-reduce, sync/base.py:22-24) and `_f_update` are fused in
sync_allreduce_update_kernel (csrc/comm.cuh): P2P loads reduce my slice, average (x 1/n_gpu), clip on
the global norm, update, P2P stores publish the new parameters."""
from accel_rl_b200.optimizers.base import BaseOptimizer


class BaseSyncOptimizer(BaseOptimizer):
    def init_comm(self, gpu_comm, rank, n_gpu):
        self._gpu_comm = gpu_comm
        self._n_gpu = n_gpu
        self._rank = rank
        self._engine.comm_init(rank, n_gpu, gpu_comm)

    @property
    def parallelism_tag(self):
        return "synchronous"
