"""Multi-GPU path on real GPUs (skipped unless ≥ 2 are visible): row shards + NCCL allreduce + remote rows over NVLink."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_problem_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29733", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
