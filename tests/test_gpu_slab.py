"""Multi-GPU parity of the slab decomposition (needs >= 2 GPUs; skipped on a single-GPU box):
N ranks over NCCL == one device, bit for bit, on every array.

Rank counts 2, 3 (uneven split), 4 and 8 -- as many as the box has.  The full case matrix runs at 2 and 4 ranks in both
halo modes; 3 and 8 ranks run the fused peer-store exchange (the mode the bench uses) on the cases that stress its
epoch-flag handshake: many epochs, the deep passes, a long run, several passes per exchange.  FDTD_SLAB_WORLDS=8,3
restricts the rank counts (GPU-time budgeting); the log of a run on an 8-GPU box is committed under profiles/."""
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


def _torchrun(world, args, env, timeout):
    """torch.distributed.run on a free local port; the port can be taken between the probe and the rendezvous (or sit
    in TIME_WAIT from the previous case): retry on EADDRINUSE with another one."""
    for attempt in range(4):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + args
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        if r.returncode == 0 or "EADDRINUSE" not in (r.stdout + r.stderr):
            return r
    return r


CASES = [
    ("3_2", 1024, 1536, 40, 61, 6, []),
    ("3_3", 900, 1200, 24, 80, 4, []),       # TFSF: every rank replicates the incident line
    ("3_2", 517, 640, 16, 33, 1, []),        # uneven split, exchange every step
    ("3_2", 1024, 1536, 40, 75, 6, ["24"]),  # 24 ghost rows: four 6-step passes per exchange (communication-avoiding)
    ("3_2", 1200, 1280, 24, 70, 6, ["30", "streamed"]),   # first 30 steps through run_streamed: no exchange at all
    ("3_2", 2048, 1536, 40, 67, 12, ["12", "plain", "v4"]),   # deep passes (depth 12, 4-wide): the ring careful kernel carries the handshake
    ("3_2", 2048, 1536, 40, 247, 6, ["6", "plain", "v4"]),    # long run: 41 epochs of the flag handshake, the bench's kernels
    ("3_3", 1600, 1024, 24, 131, 8, ["8", "plain", "v4"]),    # TFSF + deep depth-8 passes
]
STRESS = (0, 5, 6, 7)           # what 3 and 8 ranks run (fused exchange only)


def _worlds():
    sel = os.environ.get("FDTD_SLAB_WORLDS")
    return [int(x) for x in sel.split(",")] if sel else [2, 3, 4, 8]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("halo", ["p2p", "nccl"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_slab_equals_single_device(case, halo, world):
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs")
    if world not in _worlds():
        pytest.skip("rank count deselected by FDTD_SLAB_WORLDS")
    if world in (3, 8) and (halo != "p2p" or case not in STRESS):
        pytest.skip("3 and 8 ranks run the fused-exchange stress cases")
    prog, nx, ny, npml, ns, tblock, extra = CASES[case]
    r = _torchrun(world, [os.path.join(ROOT, "tests", "slab_nccl_worker.py"), prog, str(nx), str(ny), str(npml), str(ns), str(tblock)] + extra,
                  dict(os.environ, FDTD_SLAB_HALO=halo), 600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "")


def test_missing_neighbour_is_reported_not_hung():
    """One rank skips an advance() call: its neighbour's pass gives up after the (shortened) bound and the host raises."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, [os.path.join(ROOT, "tests", "slab_nccl_worker.py"), "3_2", "1024", "1536", "40", "24", "6", "6", "skip"],
                  dict(os.environ, FDTD_SLAB_HALO="p2p"), 300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "gave up waiting" in r.stdout
