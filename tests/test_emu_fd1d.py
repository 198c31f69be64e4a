"""CPU check of the fused 1D kernels' OWN SOURCE: simulation_b200/csrc/fd1d.cu compiled for the host against the
lockstep warp emulator (tests/emu/) and driven through the same C entry point (fdtd1d_advance) on numpy arrays,
bit-for-bit against the numpy oracle.  This covers what can go wrong without a GPU in sight -- segment ownership,
halo depth, edge masks, the ABC delay line, source injection, the DFT accumulators' ownership and ordering -- and is
test infrastructure only: the product path is the sm_100a build and has no CPU fallback (tests/test_gpu_fd1d.py
runs the same cases on the device)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from simulation_b200 import _lib, surface
from simulation_b200.fd2d import _phases
from tests import cases
from tests.emu import build_emu


@pytest.fixture(scope="module")
def emu():
    h = C.CDLL(build_emu.build("fd1d.cu"))
    h.fdtd1d_advance.restype = C.c_int
    h.fdtd1d_advance.argtypes = [C.POINTER(_lib.Problem1D), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p,
                                 C.POINTER(C.c_int)]
    h.fdtd_last_error.restype = C.c_char_p
    h.emu_launches.restype = C.c_longlong
    return h


def _aligned(shape, dtype):
    n = int(np.prod(shape))
    raw = np.zeros(n * np.dtype(dtype).itemsize + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype).reshape(shape)


def _copy(a, dtype):
    out = _aligned(a.shape, dtype)
    out[...] = a
    return out


class HostLine:
    """What Fdtd1D owns, on numpy arrays, for the emulated library (mirrors Fdtd1D._problem / advance)."""

    def __init__(self, p: orc.Line1D, dft: bool):
        self.p, self.dtype, self.nx, self.t, self.cur = p, np.dtype(p.dtype), p.nx, 0, 0
        flux = p.form == "flux"
        self.debye = flux and bool(np.any(p.ncx != 0) or np.any(p.ndx != 0))
        names = ["ex", "hy"] + (["dx", "ix"] if flux else []) + (["sx"] if self.debye else [])
        self.sets = [{n: _aligned((p.nx,), self.dtype) for n in names} for _ in range(2)]
        self.bc = [_aligned((4,), self.dtype), _aligned((4,), self.dtype)]
        self.coef = {n: _copy(getattr(p, n), self.dtype) for n in (("nax", "nbx", "ncx", "ndx") if flux else ("ca", "cb"))}
        self.freqs = p.freqs if dft else None
        if self.freqs is not None:
            nf = len(self.freqs)
            self.ft = {n: _aligned((nf, p.nx if n.endswith("pt") else 1), self.dtype) for n in ("r_pt", "i_pt", "r_in", "i_in")}

    def problem(self):
        p, q = self.p, _lib.Problem1D()
        q.dtype, q.nx = _lib.dtype_code(self.dtype), self.nx
        flux = p.form == "flux"
        q.flags = (_lib.ABC if p.abc else 0) | (_lib.FLUX if flux else 0) | (_lib.DEBYE if self.debye else 0)
        ptr = lambda a: a.ctypes.data
        if flux:
            q.md = _lib.Medium1D(ptr(self.coef["nax"]), ptr(self.coef["nbx"]),
                                 ptr(self.coef["ncx"]) if self.debye else None, ptr(self.coef["ndx"]) if self.debye else None)
        else:
            q.ca, q.cb = ptr(self.coef["ca"]), ptr(self.coef["cb"])
        for s in range(2):
            for k, n in enumerate(("ex", "hy", "dx", "ix", "sx")):
                q.state[s][k] = ptr(self.sets[s][n]) if n in self.sets[s] else None
            q.bc[s] = ptr(self.bc[s])
        q.src_field = 1 if flux else 0
        q.src_index, q.src_hard = int(p.src_index), int(p.src_hard)
        return q

    def advance(self, emu, nsteps, src, tblock):
        q = self.problem()
        keep = []
        if self.freqs is not None:
            ph = [_phases(self.freqs, self.p.dt, self.t + 1 + k, True, self.dtype) for k in range(nsteps)]
            cos_t = np.ascontiguousarray(np.stack([c for c, _ in ph]).reshape(-1), dtype=np.float64)
            sin_t = np.ascontiguousarray(np.stack([s for _, s in ph]).reshape(-1), dtype=np.float64)
            keep = [cos_t, sin_t]
            D = C.POINTER(C.c_double)
            q.nf, q.dft_sample = len(self.freqs), 10
            q.ft = _lib.FTrans(*[self.ft[n].ctypes.data for n in ("r_pt", "i_pt", "r_in", "i_in")])
            q.dft_cos, q.dft_sin = cos_t.ctypes.data_as(D), sin_t.ctypes.data_as(D)
        part = np.ascontiguousarray(src[self.t:self.t + nsteps], dtype=np.float64)
        out = C.c_int(-1)
        rc = emu.fdtd1d_advance(C.byref(q), self.cur, nsteps, part.ctypes.data_as(C.POINTER(C.c_double)), tblock, None, C.byref(out))
        assert rc == 0, emu.fdtd_last_error().decode()
        del keep
        self.cur, self.t = out.value, self.t + nsteps

    def get(self, n):
        if n == "bc":
            return self.bc[self.cur]
        if n in ("r_pt", "i_pt", "r_in", "i_in"):
            return self.ft[n]
        return self.sets[self.cur][n]


def _run(emu, prog, nx, ns, dtype, tblock, dft, parts=None, seed=1):
    p, src = cases.line_program(prog, nx, ns, dtype)
    if not dft:
        p.freqs = None
    host = HostLine(p, dft)
    if seed is not None:
        # a non-zero state everywhere, so that EVERY segment boundary carries signal (the programs' own pulses cover
        # a few dozen cells in these few steps and would leave the halo logic of most warps untested)
        rng = np.random.default_rng(seed)
        for n in list(host.sets[0]) + ["bc"] + (["r_pt", "i_pt", "r_in", "i_in"] if dft else []):
            a = getattr(p, n)
            a[...] = rng.uniform(-1, 1, a.shape).astype(dtype)
            host.get(n)[...] = a
    before = emu.emu_launches()
    for part in (parts or (ns,)):
        host.advance(emu, part, src, tblock)
    assert host.t == ns and emu.emu_launches() > before
    orc.advance_1d(p, src)
    names = ["ex", "hy"] + (["dx", "ix"] if p.form == "flux" else []) + (["sx"] if host.debye else [])
    names += ["bc"] if p.abc else []
    names += ["r_pt", "i_pt", "r_in", "i_in"] if dft else []
    for n in names:
        got, want = host.get(n), getattr(p, n)
        if got.tobytes() != np.ascontiguousarray(want).tobytes():
            bad = np.argwhere(got.reshape(want.shape) != want)
            raise AssertionError(f"{prog} nx={nx} T={tblock} {np.dtype(dtype).name} {n}: {len(bad)} cells differ, first {bad[:6].tolist()}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("prog,nx,tblock", [("1_1", 200, 32), ("1_2", 1300, 5), ("1_5", 1999, 32), ("2_1", 777, 16), ("2_3", 1210, 7)])
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


def test_emulated_library_rejects_what_the_device_library_rejects(emu):
    q = _lib.Problem1D()
    out = C.c_int(0)
    assert emu.fdtd1d_advance(C.byref(q), 0, 1, None, 1, None, C.byref(out)) == -1
    p, src = cases.line_program("2_2", 64, 4, np.float32)
    host = HostLine(p, True)
    q = host.problem()
    q.nf = 4
    assert emu.fdtd1d_advance(C.byref(q), 0, 1, None, 1, None, C.byref(out)) == -1
    assert b"nf=4" in emu.fdtd_last_error()
