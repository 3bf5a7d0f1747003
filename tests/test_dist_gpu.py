"""Multi-GPU paths on real GPUs: spawns tools/dist_check.py under torchrun when the box has >= 2 GPUs
(the fused transpose+exchange kernel needs one GPU per rank: ranks must never share a device)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_distributed_adi_and_transpose_two_gpus():
    env = dict(os.environ, DIST_QUICK="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29577", os.path.join(ROOT, "tools", "dist_check.py"), "16", "12"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "transpose err 0.00e+00" in out.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_element_decomposed_adi_two_gpus():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29578", os.path.join(ROOT, "tools", "dd_check.py"), "16", "12"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ADI (element decomposition) vs single-GPU" in out.stdout
