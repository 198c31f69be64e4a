"""CPU check of the fused 1D kernels' OWN SOURCE (simulation_b200/csrc/fd1d.cu compiled for the host against the
fiber-based CTA emulator of tests/emu/) together with the product's host class simulation_b200.fd1d.Fdtd1D, which
tests/emu/device.py points at the emulated library with CPU tensors as device memory.  Bit-for-bit against the numpy
oracle: segment ownership, halo depth, edge masks, the ABC delay line, source injection, the packed-arithmetic
interior warps, the DFT accumulators' ownership and ordering.  Test infrastructure only: the product path is the
sm_100a build and has no CPU fallback (tests/test_gpu_fd1d.py runs the same cases on the device)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from simulation_b200 import _lib
from tests import cases
from tests.emu import device
from tests.test_gpu_fd1d import _sim_for

torch = pytest.importorskip("torch")


@pytest.fixture
def emu(monkeypatch):
    return device.install(monkeypatch)


def _run(emu, prog, nx, ns, dtype, tblock, dft, parts=None, seed=1, fused_dft=True):
    p, src = cases.line_program(prog, nx, ns, dtype)
    if not dft:
        p.freqs = None
    sim = _sim_for(prog, nx, dtype, tblock=tblock, device="cpu", **({"freqs": p.freqs} if dft else {}))
    names = ["ex", "hy"] + (["dx", "ix"] if p.form == "flux" else []) + (["sx"] if sim.debye else []) + (["bc"] if p.abc else [])
    acc = ["r_pt", "i_pt", "r_in", "i_in"] if dft else []
    if seed is not None:
        # a non-zero state everywhere, so that EVERY segment boundary carries signal (the programs' own pulses cover
        # a few dozen cells in these few steps and would leave the halo logic of most warps untested)
        rng = np.random.default_rng(seed)
        for n in names + acc:
            a = getattr(p, n)
            a[...] = rng.uniform(-1, 1, a.shape).astype(dtype)
            (getattr(sim.ft, n) if n in acc else sim.tensor(n)).copy_(torch.from_numpy(a))
    before = emu.emu_launches()
    for part in (parts or (ns,)):
        sim.advance(part, fused_dft=fused_dft)
    assert sim.t == ns and emu.emu_launches() > before
    orc.advance_1d(p, src)
    for n in names + acc:
        got, want = sim.get(n), getattr(p, n)
        if got.tobytes() != np.ascontiguousarray(want).tobytes():
            bad = np.argwhere(got.reshape(want.shape) != want)
            raise AssertionError(f"{prog} nx={nx} T={tblock} {np.dtype(dtype).name} {n}: {len(bad)} cells differ, first {bad[:6].tolist()}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("prog,nx,tblock", [("1_1", 200, 32), ("1_2", 1300, 5), ("1_3", 2500, 64), ("1_5", 1999, 32),
                                            ("2_1", 777, 16), ("2_3", 1210, 7)])
def test_emulated_advance_matches_oracle(emu, prog, nx, tblock, dtype):
    _run(emu, prog, nx, 90, dtype, tblock, dft=False, parts=(13, 77))


def test_emulated_advance_from_rest_as_the_programs_start(emu):
    _run(emu, "1_2", 600, 120, np.float32, 32, dft=False, seed=None)
    _run(emu, "2_2", 300, 120, np.float64, 8, dft=True, seed=None)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("prog,nx,tblock", [("2_2", 397, 16), ("2_3", 1203, 8), ("2_2", 12, 3), ("2_3", 700, 1), ("2_2", 1000, 32)])
def test_emulated_advance_carrying_the_running_dft(emu, prog, nx, tblock, dtype):
    ns = 40 if tblock == 1 else 84
    _run(emu, prog, nx, ns, dtype, tblock, dft=True, parts=(7, ns - 8, 1))


def test_emulated_per_step_path_with_the_fourier_kernel(emu):
    """reference-named kernels (dxfield, exfield, fourier, hyfield: one launch each) = the fused pass = the oracle"""
    _run(emu, "2_3", 300, 25, np.float32, 8, dft=True, fused_dft=False)


@pytest.mark.parametrize("dtype,tblock", [(np.float32, 32), (np.float64, 16), (np.float32, 4), (np.float64, 10)])
def test_emulated_last_warp_owning_one_cell(emu, dtype, tblock):
    """Found by tools/fuzz_emulated_1d.py: when nx-1 is a multiple of the segment length the last warp owns cell nx-1
    alone and the right-hand ABC's ex[nx-2] is its innermost halo cell -- one sub-step short of a pass whose depth is a
    whole number of vectors.  (nx below: 3 / 6 / 2 / 5 whole segments + 1.)"""
    vec, w = (4, 512) if dtype == np.float32 else (2, 256)
    halo = -(-tblock // vec) * vec
    nx = (w - 2 * halo) * {32: 3, 16: 6, 4: 2, 10: 5}[tblock] + 1
    for prog in ("1_2", "1_5", "2_3"):
        _run(emu, prog, nx, 2 * tblock + 7, dtype, tblock, dft=False)


def test_emulated_tiny_lines(emu):
    for nx in (3, 4, 17):
        _run(emu, "1_2", nx, 40, np.float64, 5, dft=False)


def test_emulated_library_rejects_what_the_device_library_rejects(emu):
    q = _lib.Problem1D()
    out = C.c_int(0)
    assert emu.fdtd1d_advance(C.byref(q), 0, 1, None, 1, None, C.byref(out)) == -1
    sim = _sim_for("2_2", 64, np.float32, device="cpu", freqs=np.array((1e8, 2e8, 3e8), dtype=np.float32))
    q = sim._problem()
    q.nf = 4
    assert emu.fdtd1d_advance(C.byref(q), 0, 1, None, 1, None, C.byref(out)) == -1
    assert b"nf=4" in emu.fdtd_last_error()
