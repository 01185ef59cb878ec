"""Partitioned (multi-GPU) operator, V-cycle and time step against the single-GPU result of the same global problem
(scripts/mgpu_check.py under torchrun, NCCL).  Needs >= 2 visible GPUs; the host-side partition logic is covered on CPU
by tests/test_partition_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except (OSError, subprocess.TimeoutExpired):
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_run_matches_single_gpu(world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29520 + world), os.path.join(ROOT, "scripts", "mgpu_check.py"), "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "FAIL" not in r.stdout
