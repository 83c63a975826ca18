"""CPU stand-in for tests/selflaunch_worker.py (tests/test_dist_gloo.py): the one-script launch of AccelRLSync — fork ranks
1.., join a group over 127.0.0.1 — with the training loop replaced by one collective, on the gloo backend."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ACCELRL_DIST_BACKEND"] = "gloo"

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from accel_rl_b200.runners.multigpu_rl import AccelRLSync  # noqa: E402


class StubRunner(AccelRLSync):
    def __init__(self, affinities, seed=None, fail_rank=None):
        self.all_affinities, self.affinities = list(affinities), affinities[0]
        self._base_seed, self.worker_procs, self._own_group = seed, [], False
        self.rank, self.n_runners, self.fail_rank = 0, 1, fail_rank

    def train(self):
        self.launch_workers()
        seeds = [None] * self.n_runners
        dist.all_gather_object(seeds, (self._base_seed + 100 * self.rank, self.affinities["gpu"]))
        t = torch.tensor([float(self.rank + 1)])
        dist.all_reduce(t)
        if self.rank == self.fail_rank:
            raise ValueError("injected failure")
        dist.barrier()
        dist.destroy_process_group()
        self._own_group = False
        if self.rank == 0:
            for w in self.worker_procs:
                w.join(30)
            print("LAUNCH " + json.dumps(dict(world=self.n_runners, total=t.item(), seeds=seeds,
                                              exit=[w.exitcode for w in self.worker_procs])))


if __name__ == "__main__":
    n = int(sys.argv[1])
    fail = int(sys.argv[2]) if len(sys.argv) > 2 else None
    StubRunner([dict(gpu=i) for i in range(n)], seed=None if n == 3 else 7, fail_rank=fail).train()
