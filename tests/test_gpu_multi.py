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
