"""Multi-GPU checks that need two real GPUs (skipped on a one-GPU box): statistics of train-mode BatchNorm shared over
NCCL (the only exchange step of the optional training path)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sync_batchnorm_over_nccl_matches_single_gpu_whole_batch():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "_syncbn_worker.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "SYNCBN OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
