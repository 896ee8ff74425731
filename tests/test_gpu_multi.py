"""Multi-GPU tests (need >= 2 GPUs on the node; skipped otherwise): the NVLink peer-memory all-reduce kernel against
NCCL, and the data-parallel TrainStep with both backends.  One process per GPU through torch.distributed.run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_and_data_parallel_step():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "tests", "_peer_allreduce_worker.py")]
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PEER_ALLREDUCE_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
