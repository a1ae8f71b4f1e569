"""Multi-GPU checks (skipped on a single-GPU box): the peer-memory sharded step over NVLink-mapped peer tables."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_sharded_step_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "PEER_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
