"""Data-parallel update over NCCL on 2 GPUs (skipped on a single-GPU box): scripts/multigpu_check.py under
torchrun -- all-reduce, shard-vs-full trajectories, sharded TRPO/critic update == full-batch update."""
import os
import subprocess
import sys

import pytest

from relearn_b200 import _lib as L

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_update_matches_single_gpu():
    if L.lib().rl_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
