"""Multi-rank (world_size 2 and 3, gloo, CPU) test of the slab decomposition logic in
simulation_b200/slab.py: N-rank result == monolithic result, bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(world, args, timeout):
    """torch.distributed.run on a free local port, retried with another port when the rendezvous finds it taken"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for attempt in range(4):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + args
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        if r.returncode == 0 or "EADDRINUSE" not in (r.stdout + r.stderr):
            return r
    return r


@pytest.mark.parametrize("world,nx,ny,npml,ns,ghost,dtype", [
    (2, 51, 36, 5, 41, 3, "float64"),       # uneven split, ghost not dividing the step count
    (3, 64, 32, 6, 44, 4, "float32"),       # a middle rank with two neighbours
])                                          # (even splits and the per-step exchange: the emulated-engine cases below)
def test_slab_equals_monolithic(world, nx, ny, npml, ns, ghost, dtype):
    r = _torchrun(world, [os.path.join(ROOT, "tests", "slab_gloo_worker.py"), str(nx), str(ny), str(npml), str(ns), str(ghost), dtype], 300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,nx,ny,npml,ns,ghost,dtype", [
    (2, 96, 80, 8, 37, 6, "float32"),
    (3, 130, 72, 6, 29, 4, "float64"),      # a middle rank; uneven split
    (2, 64, 48, 6, 11, 1, "float32"),       # exchange every step
])
def test_slab_of_emulated_engines_equals_monolithic(world, nx, ny, npml, ns, ghost, dtype):
    """The same protocol with the PRODUCT's stepper on every rank: fd2d.Fdtd2D (host code) driving the kernels' own
    source on the CPU CTA emulator (tests/emu), ghost rows exchanged by slab.py over gloo."""
    from tests.emu import build_emu
    build_emu.build_library()               # once, before the ranks race for the build directory
    r = _torchrun(world, [os.path.join(ROOT, "tests", "slab_gloo_worker.py"), str(nx), str(ny), str(npml), str(ns), str(ghost), dtype, "emu"], 600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]


def test_partition_covers_all_rows():
    from simulation_b200 import slab
    for nx, world in ((32768 * 8, 8), (51, 2), (100, 3), (7, 7)):
        parts = [slab.partition(nx, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == nx
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
