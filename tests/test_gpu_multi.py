"""Multi-GPU parity (needs >= 2 GPUs; skipped on the 1-GPU box): tools/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_sweep_and_epoch_two_ranks():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
