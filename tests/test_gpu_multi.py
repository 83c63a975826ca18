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
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "SYNC_OK world=%d" % world in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
