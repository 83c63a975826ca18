"""Synchronous data-parallel path on >= 2 GPUs of one box (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sync_allreduce_adam_n_gpus(world):
    """fused P2P all-reduce + clip + Adam/RMSProp against the oracle on the averaged gradient and bit-identical parameters
    on every rank, at every world size the kernel specialises (reduce_slice<2,..>, <4,..>, <8,..>); bench.py --gpus N
    repeats the essential part inside every multi-GPU bench run (bench.sync_parity)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
           "127.0.0.1", "--master-port", str(29570 + world), os.path.join(ROOT, "tests", "sync_worker.py")]
    import json
    import numpy as np
    res = {}
    for overlap in ("1", "0"):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, ARL_SYNC_OVERLAP=overlap))
        assert r.returncode == 0 and "SYNC_OK world=%d" % world in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("SYNC_DIGEST ")][-1]
        res[overlap] = json.loads(line[len("SYNC_DIGEST "):])
    # the overlapped step (FC slice exchange beside the conv gradient chain, every rank updating its own replica of the
    # small tensors) and the monolithic all-reduce + update kernel produce bit-identical parameters
    assert res["1"]["params"] == res["0"]["params"]
    np.testing.assert_allclose(res["1"]["norms"], res["0"]["norms"], rtol=2e-6)


def _selflaunch(mode, world, torchrun, log_dir=None):
    import json
    script = os.path.join(ROOT, "tests", "selflaunch_worker.py")
    extra = [str(log_dir)] if log_dir is not None else []
    if torchrun:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
               "127.0.0.1", "--master-port", str(29590 + world), script, mode, str(world)] + extra
    else:
        cmd = [sys.executable, script, mode, str(world)] + extra
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("SELFLAUNCH_DIGEST ")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(lines[0][len("SELFLAUNCH_DIGEST "):])


@pytest.mark.parametrize("world", [2, 8])
def test_one_script_launch_trains_like_torchrun(world):
    """AccelRLSync(affinities=[gpu 0, gpu 1, ...]).train() from ONE plain python process forks its per-GPU runners
    (runners/multigpu_rl_base.py:20-45) and ends with the same parameters, bit for bit, as the same script launched one
    process per GPU by torch.distributed.run; AccelRLAsync launches the same way (its result depends on learner timing,
    so it is only required to finish with finite parameters)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    own = _selflaunch("sync", world, torchrun=False)
    tr = _selflaunch("sync", world, torchrun=True)
    assert own["n_itr"] == tr["n_itr"] == 3 and not own["torchrun"] and tr["torchrun"]
    assert own["params"] == tr["params"]
    _selflaunch("async", world, torchrun=False)


def test_multi_gpu_runner_with_offline_evaluation(tmp_path):
    """AccelRLEvalSync (runners/multigpu_rl.py:29-37, log mixins multigpu_rl_base.py:233-250): synchronous learners where only
    the master runs the evaluation episodes and writes the log while the others wait; same parameters from the one-script
    launch and from torchrun; progress.csv holds one evaluation row per log point with the evaluation columns."""
    import csv
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    own = _selflaunch("evalsync", 2, torchrun=False, log_dir=tmp_path / "own")
    tr = _selflaunch("evalsync", 2, torchrun=True, log_dir=tmp_path / "tr")
    assert own["n_itr"] == tr["n_itr"] == 3 and own["params"] == tr["params"]
    rows = list(csv.DictReader(open(tmp_path / "own" / "progress.csv")))
    assert len(rows) == 3 and all(int(r["TrajsInEval"]) > 0 and int(r["StepsInEval"]) > 0 for r in rows)
    assert [int(r["Iteration"]) for r in rows] == [0, 1, 2]
