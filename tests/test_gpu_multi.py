"""Multi-GPU (torchrun) test of the row-partitioned path: halo exchange over peer memory + GCNConv on a PartitionedGraph.
Skipped on a box with fewer than two GPUs (the one-GPU parts of the path are covered by test_gpu_dist.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["ce", "sm"])
def test_partitioned_aggregation_and_gcn_on_two_gpus(cuda, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if torch.cuda.device_count() < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, STG_HALO_MODE=mode), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist_worker ok" in res.stdout
