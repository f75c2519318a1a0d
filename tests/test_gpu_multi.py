"""GPU, needs >= 2 devices (skipped otherwise): owner-partitioned node memory over NCCL."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partitioned_engine_matches_single_gpu():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(REPO, "tests", "dist_partition_check.py")],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "partition check OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_embedding_eval_matches_one_replica():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29537",
                        os.path.join(REPO, "tests", "dist_eval_check.py")],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "sharded eval check OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
