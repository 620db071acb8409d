"""Multi-rank path on real GPUs: row shards + the peer-memory exchange (or NCCL) + remote rows.

  * test_sharded_problem_ranks_share_one_gpu runs everywhere a GPU exists (the driver's 1-GPU test box included): two processes
    on cuda:0, shards and exchange arenas mapped through CUDA IPC between the processes.
  * the two-GPU tests need `gpurun --gpus 2` and are skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch(world, port, **env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, **env})
    assert res.returncode == 0 and "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    return res.stdout


def test_sharded_problem_ranks_share_one_gpu():
    out = launch(2, 29731, CIAO_TEST_SHARE_GPU="1", CIAO_P2P_TIMEOUT_MS="60000")
    assert "share_gpu p2p" in out


def test_sharded_problem_three_ranks_share_one_gpu():
    out = launch(3, 29732, CIAO_TEST_SHARE_GPU="1", CIAO_P2P_TIMEOUT_MS="60000")
    assert "world 3 share_gpu p2p" in out


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_problem_two_gpus(exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = launch(2, 29733, CIAO_TEST_EXCHANGE=exchange)
    assert f"one_gpu_per_rank {exchange}" in out
