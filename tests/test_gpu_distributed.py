"""Row-sharded solve on real GPUs (needs >= 2 devices; the 1-GPU driver run skips it)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        from structurepreservingiterativesolvers_b200 import _native as nat
        return nat.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
def test_two_gpu_sharded_solve_matches_single_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "dist_gpu_check.py"), "300000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "-> OK" in res.stdout


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
def test_two_gpu_soak_of_the_flag_protocol():
    """200 back-to-back sharded solves through fresh sessions on 2 ranks (persistent NVLink communicator, cached
    sharding plan, LL-protocol reductions and halo exchanges): every solve must return the bits of the first one."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "dist_gpu_check.py"), "30000", "auto", "lkdv", "soak", "200"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "-> OK" in res.stdout and "mismatching solves summed over ranks: 0" in res.stdout
