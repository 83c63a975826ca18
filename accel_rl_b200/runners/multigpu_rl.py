"""Multi-GPU runners (reference: accel_rl/runners/multigpu_rl.py:7-40, multigpu_rl_base.py:12-250).

The reference forks one full runner per GPU from a master process and ships the NCCL clique id and
the rank-0 parameters through an mp.Manager dict.  Here the launch is one process per GPU
(torchrun / torch.distributed, backend nccl): every rank constructs the same AccelRLSync with its own
`affinities` entry; rank r seeds with seed + 100*r (multigpu_rl_base.py:28), rank 0's initial
parameters are broadcast (:119,:142-143), and n_itr is computed from sample_size * n_runners
(:62-63).  Worker traj_infos are gathered to rank 0 for logging (:216-231)."""
import numpy as np
import torch
import torch.distributed as dist

from accel_rl_b200.runners.accel_rl import AccelRL
from accel_rl_b200.util import logger


class AccelRLSync(AccelRL):
    def __init__(self, affinities=None, seed=None, **kwargs):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n_runners = dist.get_world_size() if dist.is_initialized() else 1
        if isinstance(affinities, (list, tuple)):
            self.all_affinities = list(affinities)
            affinities = affinities[self.rank] if self.rank < len(affinities) else dict()
        super().__init__(affinities=affinities, seed=seed, **kwargs)
        self._base_seed = seed

    def startup(self, master=True):
        if self._base_seed is not None:
            self.seed = self._base_seed + 100 * self.rank
        n_itr = super().startup(master=True)
        self.init_comm()
        if self.rank != 0:
            logger.configure(None, quiet=True)
        return n_itr

    def get_n_itr(self, sample_size):
        n_itr = super().get_n_itr(sample_size * self.n_runners)
        self._sample_size = sample_size * self.n_runners
        return n_itr

    def init_comm(self):
        eng = self.policy.engine
        # rank 0's initial parameters to everyone (reference ships them through a Manager dict)
        if self.n_runners > 1:
            dist.broadcast(eng.params, src=0)
            eng.pack()

        def exchange(handle):
            if self.n_runners == 1:
                return [handle]
            out = [None] * self.n_runners
            dist.all_gather_object(out, handle)
            return out

        self.algo.optimizer.init_comm(exchange, self.rank, self.n_runners)
        self._initial_param_vector = self.policy.get_param_values()
        if self.n_runners > 1:
            dist.barrier()

    def store_diagnostics(self, itr, samples_data, opt_data, traj_infos, opt_infos):
        if self.n_runners > 1 and (itr + 1) % self._log_interval_itrs == 0:
            gathered = [None] * self.n_runners
            dist.all_gather_object(gathered, [dict(t) for t in self._pending_trajs + list(traj_infos)])
            self._pending_trajs = []
            traj_infos = [t for g in gathered for t in g]
        elif self.n_runners > 1:
            self._pending_trajs += list(traj_infos)
            traj_infos = []
        super().store_diagnostics(itr, samples_data, opt_data, traj_infos, opt_infos)

    def init_logging(self):
        self._pending_trajs = []
        super().init_logging()

    @property
    def parallelism_tag(self):
        return "synchronous"


class AccelRLAsync(AccelRLSync):
    """reference: multigpu_rl.py:18-26 / multigpu_rl_base.py:161-208.  Same launch as AccelRLSync (one process per
    GPU, rank r seeds with seed + 100*r, rank 0's initial parameters broadcast) but no collective on the data path:
    every learner pushes its locally clipped gradient into the central (params, m, v) store in rank 0's HBM under
    the chunk locks and pulls the new parameters (optimizers/async_/base.py)."""

    def init_comm(self):
        eng = self.policy.engine
        if self.n_runners > 1:
            dist.broadcast(eng.params, src=0)      # par_objs.dict["initial_param_values"] (multigpu_rl_base.py:171,189)
            eng.pack()
            torch.cuda.synchronize()

        def exchange(handle):
            if self.n_runners == 1:
                return [handle]
            out = [None] * self.n_runners
            dist.all_gather_object(out, handle)
            return out

        self.algo.optimizer.init_comm(self.rank, self.n_runners, dict(exchange=exchange))
        if hasattr(self.sampler, "poll_init"):
            # the reference leaves this call to the experiment script (nothing in its tree makes it)
            self.sampler.poll_init(self.algo.optimizer.central_params_handle, None)
        self._initial_param_vector = self.policy.get_param_values()
        if self.n_runners > 1:
            dist.barrier()

    @property
    def parallelism_tag(self):
        return "asynchronous"
