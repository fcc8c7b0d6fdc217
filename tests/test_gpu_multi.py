"""Launch the multi-rank device checks under torchrun when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.multigpu
def test_two_rank_fabric_and_fused_step():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(REPO, "tests", "mp_device_check.py")]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MP_DEVICE_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.multigpu
def test_two_rank_distributed_topk():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29618",
           os.path.join(REPO, "tests", "mp_topk_check.py")]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MP_TOPK_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
