"""reference: accel_rl/optimizers/async/base.py:11-104 (host shared-memory parameter server with
chunk locks).  Round 1 ships the synchronous learner; the asynchronous one is the next §8 row
 — constructing it is allowed (so algos.mA3C/mAPPO import), using it raises."""
from accel_rl_b200.optimizers.base import BaseOptimizer


class BaseAsyncOptimizer(BaseOptimizer):
    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs
        self.n_update_chunks = kwargs.get("n_update_chunks", 1)

    def initialize(self, *args, **kwargs):
        raise NotImplementedError("asynchronous multi-learner path is not built yet (SURVEY.md §8 row a11)")

    def init_comm(self, rank, n_runners, par_objs):
        raise NotImplementedError("asynchronous multi-learner path is not built yet (SURVEY.md §8 row a11)")

    @property
    def parallelism_tag(self):
        return "asynchronous"
