"""Multi-GPU runners (reference: accel_rl/runners/multigpu_rl.py:7-40, multigpu_rl_base.py:12-250).

The reference forks one full runner per GPU from a master process and ships the NCCL clique id and
the rank-0 parameters through an mp.Manager dict.  Here every GPU has one process joined in a
torch.distributed group (backend nccl); rank r seeds with seed + 100*r (multigpu_rl_base.py:28), rank 0's
initial parameters are broadcast (:119,:142-143), and n_itr is computed from sample_size * n_runners
(:62-63).  Worker traj_infos are gathered to rank 0 for logging (:216-231).

Two ways to get the processes, same training either way (tests/test_gpu_multi.py compares them bit for bit):
  * one script, like the reference: construct the runner with `affinities=[dict(gpu=0), dict(gpu=1), ...]` in a
    plain `python script.py`; `train()` forks ranks 1.. itself (launch_workers, as multigpu_rl_base.py:20-45) and
    forms the group over 127.0.0.1.  The runner must be built before the process touches CUDA (fork);
  * torchrun / torch.distributed.run: every rank runs the script, the group already exists, nothing is forked."""
import multiprocessing as mp
import os
import socket
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

from accel_rl_b200.runners.accel_rl import AccelRL, AccelRLEval
from accel_rl_b200.util import logger
from accel_rl_b200.util.misc import make_seed


def _free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _MultiGpuBase(object):
    """launch + process group + seeds + n_itr of the multi-GPU runners (multigpu_rl_base.py:12-108); mixed in front of
    AccelRL / AccelRLEval together with a communication flavour (_SyncComm / _AsyncComm) and a log flavour"""

    def __init__(self, affinities=None, seed=None, **kwargs):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n_runners = dist.get_world_size() if dist.is_initialized() else 1
        if isinstance(affinities, (list, tuple)):
            self.all_affinities = list(affinities)
            affinities = affinities[self.rank] if self.rank < len(affinities) else dict()
        else:
            self.all_affinities = [affinities]
        super().__init__(affinities=affinities, seed=seed, **kwargs)
        self._base_seed = seed
        self.worker_procs = []
        self._own_group = False

    def launch_workers(self):
        """one script, N GPUs (multigpu_rl_base.py:20-45): fork a full runner for every rank >= 1 and join them in a
        process group.  Nothing to do under torchrun (the group exists) or with one affinity."""
        n = len(self.all_affinities)
        if dist.is_initialized() or n <= 1:
            return
        if torch.cuda.is_initialized():
            raise RuntimeError("AccelRLSync/AccelRLAsync fork their per-GPU runners: build and train the runner before this "
                               "process initialises CUDA, or launch one process per GPU with torch.distributed.run")
        if self._base_seed is None:
            self._base_seed = make_seed()                     # one base seed for all ranks (multigpu_rl_base.py:22-23)
        port = _free_port()
        ctx = mp.get_context("fork")
        self.worker_procs = [ctx.Process(target=self._worker_main, args=(rank, n, port)) for rank in range(1, n)]
        for w in self.worker_procs:
            w.start()
        self._join_group(0, n, port)

    def _join_group(self, rank, n, port):
        backend = os.environ.get("ACCELRL_DIST_BACKEND", "nccl")
        dist.init_process_group(backend, init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=n)
        self._own_group = True
        self.rank, self.n_runners = rank, n
        self.affinities = self.all_affinities[rank]

    def _worker_main(self, rank, n, port):
        """a forked runner of rank >= 1 (the reference's WorkerCls.train, multigpu_rl_base.py:77-108)"""
        code = 0
        try:
            self.worker_procs = []
            self._join_group(rank, n, port)
            self.train()
        except BaseException:
            traceback.print_exc()
            code = 1
        finally:
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(code)                                    # no atexit handlers of the parent's state in a forked child

    def shutdown(self):
        super().shutdown()
        if self._own_group:
            dist.barrier()
            dist.destroy_process_group()
            self._own_group = False
        for w in self.worker_procs:
            w.join(60)
        bad = [w.exitcode for w in self.worker_procs if w.exitcode != 0]
        self.worker_procs = []
        if bad:
            raise RuntimeError("worker runners exited with %s" % bad)

    def startup(self, master=True):
        self.launch_workers()
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

    def _exchange(self, handle):
        """all-gather of one picklable object (the 64-byte IPC handles)"""
        if self.n_runners == 1:
            return [handle]
        out = [None] * self.n_runners
        dist.all_gather_object(out, handle)
        return out


class _SyncComm(object):
    """multigpu_rl_base.py:109-153"""

    def init_comm(self):
        eng = self.policy.engine
        # rank 0's initial parameters to everyone (reference ships them through a Manager dict)
        if self.n_runners > 1:
            dist.broadcast(eng.params, src=0)
            eng.pack()

        self.algo.optimizer.init_comm(self._exchange, self.rank, self.n_runners)
        self._initial_param_vector = self.policy.get_param_values()
        if self.n_runners > 1:
            dist.barrier()

    def check_replicas(self):
        """synchronous learners hold bit-identical parameter replicas by construction (the FC slices are published by their
        owners, the small tensors are updated from identically ordered sums): compared at every log point through a 64-bit
        checksum per rank — a mismatch means a lost update and must not train on silently"""
        if self.n_runners == 1:
            return
        p = self.policy.engine.params
        # (position-weighted so that swapped values do not cancel; int64 arithmetic wraps, which is fine for a checksum)
        words = p.view(torch.int32).to(torch.int64)
        weights = torch.arange(1, p.numel() + 1, device=p.device, dtype=torch.int64)
        checksum = int((words * weights).sum().item())
        sums = [None] * self.n_runners
        dist.all_gather_object(sums, checksum)
        if len(set(sums)) != 1:
            raise RuntimeError("synchronous learners hold different parameters (checksums per rank: %s)" % sums)

    @property
    def parallelism_tag(self):
        return "synchronous"


class _OnlineLog(object):
    """multigpu_rl_base.py:216-231: the master logs the trajectories of every runner"""

    def store_diagnostics(self, itr, samples_data, opt_data, traj_infos, opt_infos):
        if self.n_runners > 1 and (itr + 1) % self._log_interval_itrs == 0:
            if hasattr(self, "check_replicas"):
                self.check_replicas()
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


class _EvalLog(object):
    """multigpu_rl_base.py:233-250: offline evaluation on a multi-GPU run — only the master evaluates and logs; the other
    runners skip the evaluation, keep no diagnostics and wait for the master at the log point"""

    def eval_policy(self, itr):
        if self.rank != 0:
            return None, None
        return super().eval_policy(itr)

    def store_diagnostics(self, itr, samples_data, opt_data, traj_infos, opt_infos):
        if self.rank == 0:
            super().store_diagnostics(itr, samples_data, opt_data, traj_infos, opt_infos)

    def log_diagnostics(self, itr, eval_traj_infos, eval_time):
        if hasattr(self, "check_replicas"):
            self.check_replicas()
        if self.rank == 0:
            super().log_diagnostics(itr, eval_traj_infos, eval_time)
        if self.n_runners > 1:
            dist.barrier()


class _AsyncComm(object):
    """reference: multigpu_rl.py:18-26 / multigpu_rl_base.py:161-208.  Same launch as the synchronous runners (one process
    per GPU, rank r seeds with seed + 100*r, rank 0's initial parameters broadcast) but no collective on the data path:
    every learner pushes its locally clipped gradient into the central (params, m, v) store in rank 0's HBM under
    the chunk locks and pulls the new parameters (optimizers/async_/base.py)."""

    def init_comm(self):
        eng = self.policy.engine
        if self.n_runners > 1:
            dist.broadcast(eng.params, src=0)      # par_objs.dict["initial_param_values"] (multigpu_rl_base.py:171,189)
            eng.pack()
            torch.cuda.synchronize()

        self.algo.optimizer.init_comm(self.rank, self.n_runners, dict(exchange=self._exchange))
        if hasattr(self.sampler, "poll_init"):
            # the reference leaves this call to the experiment script (nothing in its tree makes it)
            self.sampler.poll_init(self.algo.optimizer.central_params_handle, None)
        self._initial_param_vector = self.policy.get_param_values()
        if self.n_runners > 1:
            dist.barrier()

    @property
    def parallelism_tag(self):
        return "asynchronous"


class AccelRLSync(_MultiGpuBase, _SyncComm, _OnlineLog, AccelRL):
    """multigpu_rl.py:7-15"""


class AccelRLAsync(_MultiGpuBase, _AsyncComm, _OnlineLog, AccelRL):
    """multigpu_rl.py:18-26"""


class AccelRLEvalSync(_MultiGpuBase, _SyncComm, _EvalLog, AccelRLEval):
    """multigpu_rl.py:29-37 (needs a sampler with evaluate_policy, e.g. AAOEvalSampler)"""


class AccelRLEvalAsync(_MultiGpuBase, _AsyncComm, _EvalLog, AccelRLEval):
    """multigpu_rl.py:40-48"""
