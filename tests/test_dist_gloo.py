"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: rank seeding, handle exchange order, iteration
count from the global batch, trajectory gathering, and the n-rank averaged update restated in the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine(object):
    """records what the sync optimizer asks of the engine (no CUDA)"""

    def __init__(self):
        self.params = torch.zeros(8)
        self.calls = []
        self.handles = None

    def comm_init(self, rank, world, exchange):
        self.handles = exchange(bytes([rank]) * 64)
        self.calls.append(("comm_init", rank, world))

    def pack(self):
        self.calls.append(("pack",))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from accel_rl_b200.runners.multigpu_rl import AccelRLSync
    from accel_rl_b200.optimizers.sync.base import BaseSyncOptimizer

    class Opt(BaseSyncOptimizer):
        def __init__(self, eng):
            self._engine = eng

    class Algo(object):
        need_extra_obs = True
        opt_info_keys = ["GradNorm"]

    class Pol(object):
        def __init__(self, eng):
            self.engine = eng

        def get_param_values(self):
            return self.engine.params.numpy().copy()

    eng = _FakeEngine()
    eng.params += float(rank + 1)                 # ranks start different; rank 0 must win
    algo = Algo()
    algo.optimizer = Opt(eng)
    r = AccelRLSync(algo=algo, policy=Pol(eng), sampler=None, n_steps=1e6, seed=5, log_interval_steps=1e5,
                    affinities=[dict(gpu=0), dict(gpu=1)])
    r.seed = r._base_seed + 100 * r.rank          # what startup() does before seeding
    n_itr = r.get_n_itr(32768)
    r.init_logging()
    r.init_comm()
    # trajectory gathering on a logging iteration
    r._log_interval_itrs = 1
    traj = [dict(Length=10 + rank, Return=1.0, RawReturn=1.0, NonzeroRewards=1, DiscountedReturn=0.9)]

    class S(object):
        agent_infos = dict(prob=np.full((4, 4), 0.25, np.float32))
    r.policy.distribution = __import__("accel_rl_b200.distributions", fromlist=["Categorical"]).Categorical(4)
    r._log_entropy = False
    r.store_diagnostics(0, S(), None, traj, dict(GradNorm=[1.0]))          # (runs check_replicas: identical here)
    params_after_init = eng.params.numpy().copy()
    # the replica guard: one ulp of difference on one rank must stop the run on every rank
    if rank == 1:
        eng.params[3] = torch.nextafter(eng.params[3], torch.tensor(10.0))
    try:
        r.check_replicas()
        mismatch = False
    except RuntimeError:
        mismatch = True
    out[rank] = dict(seed=r.seed, n_itr=n_itr, sample_size=r._sample_size, params=params_after_init,
                     handles=[h[0] for h in eng.handles], aff=r.affinities, calls=eng.calls, mismatch=mismatch,
                     traj_lengths=sorted(t["Length"] for t in r._traj_infos))
    dist.destroy_process_group()


def test_sync_runner_host_logic_two_ranks():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = out[0], out[1]
    assert (r0["seed"], r1["seed"]) == (5, 105)                       # multigpu_rl_base.py:28
    assert r0["n_itr"] == r1["n_itr"] == 16                            # from sample_size * n_runners
    assert r0["sample_size"] == 2 * 32768
    assert np.array_equal(r0["params"], r1["params"]) and r0["params"][0] == 1.0   # rank-0 broadcast
    assert r0["handles"] == r1["handles"] == [0, 1]                    # handles gathered in rank order
    assert r0["aff"] == dict(gpu=0) and r1["aff"] == dict(gpu=1)
    assert ("comm_init", 1, 2) in r1["calls"]
    assert r0["traj_lengths"] == r1["traj_lengths"] == [10, 11]
    assert r0["mismatch"] and r1["mismatch"]                           # replica checksum guard (runners/multigpu_rl.py)


def test_oracle_virtual_ranks_average_equals_concatenated_batch():
    """sync DP semantics (sync_ppo_optimizer.py:13-78, optimizers/util.py:63-67): mean of per-rank minibatch
    gradients == gradient of the concatenated batch (losses are means over equal-size minibatches)."""
    from oracle import net as onet
    spec = onet.CNN_SPECS[0]
    rng = np.random.RandomState(0)
    flat = onet.init_params(spec, (4, 104, 80), 4, np.random.RandomState(1), np.random.RandomState(2))
    n = 6
    obs = rng.randint(0, 256, (2 * n, 4, 104, 80), dtype=np.uint8)
    act = rng.randint(0, 4, 2 * n).astype(np.uint8)
    adv = rng.randn(2 * n).astype(np.float32); ret = rng.randn(2 * n).astype(np.float32)
    oldp = rng.dirichlet(np.ones(4), 2 * n).astype(np.float32)
    gs = []
    for r in range(2):
        sl = slice(r * n, (r + 1) * n)
        _, g, _ = onet.loss_and_grad(flat, obs[sl], act[sl], adv[sl], ret[sl], oldp[sl], spec, 4, "ppo")
        gs.append(g)
    _, g_all, _ = onet.loss_and_grad(flat, obs, act, adv, ret, oldp, spec, 4, "ppo")
    avg = 0.5 * (gs[0] + gs[1])
    assert np.linalg.norm(avg - g_all) / np.linalg.norm(g_all) < 1e-4


def _async_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.synchronize = lambda *a, **k: None          # no device in this test
    from accel_rl_b200.runners.multigpu_rl import AccelRLAsync
    from accel_rl_b200.optimizers.async_.base import BaseAsyncOptimizer

    class Eng(_FakeEngine):
        def async_init(self, rank, world, n_update_chunks, exchange):
            self.handles = exchange(bytes([10 + rank]) * 64)
            self.calls.append(("async_init", rank, world, n_update_chunks))
            return 148

    class Opt(BaseAsyncOptimizer):
        def __init__(self, eng):
            self._engine = eng
            self.n_update_chunks = 3

    class Algo(object):
        need_extra_obs = True
        opt_info_keys = ["GradNorm"]

    class Pol(object):
        def __init__(self, eng):
            self.engine = eng

        def get_param_values(self):
            return self.engine.params.numpy().copy()

    eng = Eng()
    eng.params += float(rank + 1)
    algo = Algo()
    algo.optimizer = Opt(eng)
    r = AccelRLAsync(algo=algo, policy=Pol(eng), sampler=None, n_steps=1e6, seed=7, log_interval_steps=1e5,
                     affinities=[dict(gpu=0), dict(gpu=1)])
    r.init_comm()
    out[rank] = dict(tag=r.parallelism_tag, params=eng.params.numpy().copy(), handles=[h[0] for h in eng.handles],
                     calls=eng.calls, regions=algo.optimizer.n_lock_regions, rank=algo.optimizer._rank,
                     n=algo.optimizer._n_runners)
    dist.destroy_process_group()


def test_async_runner_host_logic_two_ranks():
    """AccelRLAsync.init_comm (multigpu_rl_base.py:161-208, optimizers/async/base.py:13-41): rank 0's parameters are
    broadcast, the IPC handles are all-gathered in rank order (rank 0's entry is the central store), the optimizer
    learns its rank / world size / chunk count"""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_async_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = out[0], out[1]
    assert r0["tag"] == r1["tag"] == "asynchronous"
    assert np.array_equal(r0["params"], r1["params"]) and r0["params"][0] == 1.0
    assert r0["handles"] == r1["handles"] == [10, 11]
    assert ("async_init", 0, 2, 3) in r0["calls"] and ("async_init", 1, 2, 3) in r1["calls"]
    assert ("pack",) in r1["calls"]
    assert (r0["regions"], r1["rank"], r1["n"]) == (148, 1, 2)


@pytest.mark.parametrize("world", [2, 3])
def test_one_script_launch_forks_and_joins_the_ranks(world):
    """AccelRLSync.launch_workers (reference runners/multigpu_rl_base.py:20-45): one plain python process forks ranks 1..,
    all join one group over 127.0.0.1; rank r gets affinities[r] and seed base + 100 r (one base seed for all, drawn by the
    master when none is given); a failing forked runner takes the job down (tests/launch_stub.py, gloo)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "launch_stub.py"), str(world)], capture_output=True,
                       text=True, timeout=300, cwd=root)
    line = [l for l in r.stdout.splitlines() if l.startswith("LAUNCH ")]
    assert r.returncode == 0 and len(line) == 1, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[0][len("LAUNCH "):])
    assert out["world"] == world and out["total"] == world * (world + 1) / 2 and out["exit"] == [0] * (world - 1)
    base = out["seeds"][0][0]
    assert [tuple(s) for s in out["seeds"]] == [(base + 100 * k, k) for k in range(world)]
    if world == 2:
        assert base == 7
        r = subprocess.run([sys.executable, os.path.join(root, "tests", "launch_stub.py"), "2", "1"], capture_output=True,
                           text=True, timeout=300, cwd=root)
        # the forked runner reports its exception and dies; the master's next collective fails instead of hanging
        assert r.returncode != 0 and "injected failure" in r.stderr and "LAUNCH " not in r.stdout
