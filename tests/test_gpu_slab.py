"""Multi-GPU parity of the slab decomposition (needs >= 2 GPUs; skipped on a single-GPU box):
N ranks over NCCL == one device, bit for bit, on every array."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("prog,nx,ny,npml,ns,tblock,extra", [
    ("3_2", 1024, 1536, 40, 61, 6, []),
    ("3_3", 900, 1200, 24, 80, 4, []),       # TFSF: every rank replicates the incident line
    ("3_2", 517, 640, 16, 33, 1, []),        # uneven split, exchange every step
    ("3_2", 1024, 1536, 40, 75, 6, ["24"]),  # 24 ghost rows: four 6-step passes per exchange (communication-avoiding)
    ("3_2", 1200, 1280, 24, 70, 6, ["30", "streamed"]),   # first 30 steps through run_streamed: no exchange at all
])
@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_slab_equals_single_device(prog, nx, ny, npml, ns, tblock, extra, halo):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "slab_nccl_worker.py"), prog, str(nx), str(ny), str(npml), str(ns), str(tblock)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, FDTD_SLAB_HALO=halo))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
