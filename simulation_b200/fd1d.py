"""1D Ex/Hy host side: the reference's step functions under their own names (fd1d/program/fd1d_2_1.py:43-61,
fd1d_2_3.py:73-94; programs 1_1-1_5 inline the same statements in ``main()``, fd1d_1_5.py:63-72) on CUDA
tensors, and :class:`Fdtd1D`, which owns the arrays of one line and replaces the reference time loop by a
fused, temporally blocked ``advance``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .fd2d import _NP_DT, _TORCH_DT, _code, _phase_tables, _phases, _ptr, _require_cuda, _stream, ftrans


class medium(NamedTuple):
    nax: torch.Tensor
    nbx: torch.Tensor
    ncx: Optional[torch.Tensor] = None
    ndx: Optional[torch.Tensor] = None


@dataclass(frozen=True)
class LineSource:
    """``field[index] = w(t)`` (hard) or ``field[index] += w(t)`` (soft); field is 'ex' or 'dx'."""
    index: int
    waveform: object
    hard: bool = False
    field: str = "ex"


def _src(target, source, t):
    if source is None:
        return None
    return _lib.Source(target.data_ptr(), int(source.index), int(bool(source.hard)),
                       float(source.waveform.table(int(t), 1)[0]))


def exfield(t: int, nx: int, ex, hy, *, ca=None, cb=None, source: Optional[LineSource] = None) -> None:
    """FDTD-form E update ``ex[1:nx] = ca*ex + cb*(hy[i-1]-hy[i])`` then the source of step ``t``."""
    _require_cuda(ex, hy, ca, cb)
    s = _src(ex, source, t)
    check(lib().fdtd1d_exfield(_code(ex), nx, _ptr(ca), _ptr(cb), _ptr(ex), _ptr(hy),
                               C.byref(s) if s is not None else None, _stream()), "exfield")


def dxfield(t: int, nx: int, dx, hy, *, source: Optional[LineSource] = None) -> None:
    _require_cuda(dx, hy)
    s = _src(dx, source, t)
    check(lib().fdtd1d_dxfield(_code(dx), nx, _ptr(dx), _ptr(hy), C.byref(s) if s is not None else None, _stream()),
          "dxfield")


def exfield_flux(nx: int, md: medium, dx, ix, ex, sx=None) -> None:
    """Flux-form ``ex = nax*(dx-ix[-ncx*sx]); ix += nbx*ex; [sx = ncx*sx + ndx*ex]``."""
    _require_cuda(dx, ix, ex, sx, md.nax, md.nbx, md.ncx, md.ndx)
    ms = _lib.Medium1D(md.nax.data_ptr(), md.nbx.data_ptr(),
                       None if md.ncx is None else md.ncx.data_ptr(), None if md.ndx is None else md.ndx.data_ptr())
    check(lib().fdtd1d_exfield_flux(_code(ex), nx, C.byref(ms), _ptr(dx), _ptr(ix), _ptr(sx), _ptr(ex), _stream()),
          "exfield_flux")


def fourier(t: int, nf: int, nx: int, dt: float, freq, ex, ft: ftrans) -> None:
    """Running DFT of Ex and of the source sample ``ex[10]`` (fd1d/program/fd1d_2_2.py:65-71, same argument order)."""
    _require_cuda(ex, *ft)
    c, s = _phases(freq, dt, t, True, _NP_DT[ex.dtype])
    fs = ft.as_struct()
    D = C.POINTER(C.c_double)
    check(lib().fdtd1d_fourier(_code(ex), nf, nx, c.ctypes.data_as(D), s.ctypes.data_as(D), _ptr(ex), 10, C.byref(fs),
                               _stream()), "fourier")


def hyfield(nx: int, ex, hy, bc=None) -> None:
    """ABC (when ``bc`` is given) then the H update."""
    _require_cuda(ex, hy, bc)
    check(lib().fdtd1d_hyfield(_code(ex), nx, _ptr(ex), _ptr(hy), _ptr(bc), int(bc is not None), _stream()), "hyfield")


class Fdtd1D:
    """One 1D line.  ``form='fdtd'`` (``ca``/``cb``; programs 1_1-1_5) or ``'flux'`` (``nax``..``ndx``; 2_1-2_3)."""

    FIELDS = ("ex", "hy", "dx", "ix", "sx")
    DFT_SAMPLE = 10              # the running DFT's source sample is ex[10] (fd1d/program/fd1d_2_2.py:70-71)

    def __init__(self, nx: int, dtype=np.float32, *, form: str = "fdtd", abc: bool = True, source: Optional[LineSource] = None,
                 ca=None, cb=None, nax=None, nbx=None, ncx=None, ndx=None, device=None, tblock: int = 32, freqs=None,
                 dt: float = 0.01 / 6e8):
        if not torch.cuda.is_available():
            raise _lib.FdtdError("Fdtd1D needs a CUDA device: the product has no CPU path")
        lib()
        if form not in ("fdtd", "flux"):
            raise ValueError(form)
        self.nx, self.form, self.abc, self.source = int(nx), form, bool(abc), source
        self.np_dtype = np.dtype(dtype)
        self.dtype = _TORCH_DT[self.np_dtype]
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.tblock = int(tblock)
        self.debye = form == "flux" and ncx is not None
        self.t = 0
        up = lambda a: None if a is None else torch.from_numpy(
            np.ascontiguousarray(np.asarray(a, dtype=self.np_dtype))).to(self.device)
        self.ca, self.cb = up(ca), up(cb)
        if form == "flux":
            one = np.full(nx, 1.0, dtype=self.np_dtype)
            zero = np.zeros(nx, dtype=self.np_dtype)
            self.md = medium(up(one if nax is None else nax), up(zero if nbx is None else nbx), up(ncx), up(ndx))
        else:
            self.md = None
        names = self.FIELDS[:2] if form == "fdtd" else (self.FIELDS if self.debye else self.FIELDS[:4])
        z = lambda n: torch.zeros(n, dtype=self.dtype, device=self.device)
        self._sets = [{n: z(self.nx) for n in names} for _ in range(2)]
        self._bc = [z(4), z(4)]
        self._cur = 0
        self.freqs, self.dt = (None if freqs is None else np.asarray(freqs, dtype=self.np_dtype)), float(dt)
        if self.freqs is not None:
            nf = len(self.freqs)
            zz = lambda *shape: torch.zeros(shape, dtype=self.dtype, device=self.device)
            self.ft = ftrans(zz(nf, self.nx), zz(nf, self.nx), zz(nf, 1), zz(nf, 1))
        else:
            self.ft = None
        if source is not None and source.field == "dx" and form != "flux":
            raise _lib.FdtdError("a dx source needs the flux form")

    def tensor(self, name: str) -> torch.Tensor:
        return self._bc[self._cur] if name == "bc" else self._sets[self._cur][name]

    def get(self, name: str) -> np.ndarray:
        if name in ("r_pt", "i_pt", "r_in", "i_in"):
            return getattr(self.ft, name).cpu().numpy()
        return self.tensor(name).cpu().numpy()

    def set(self, name: str, host) -> None:
        self.tensor(name).copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(host, dtype=self.np_dtype))))

    def checkpoint(self) -> dict:
        """State arrays, ABC delay line, running-DFT accumulators and step counter as host arrays (see Fdtd2D.checkpoint)."""
        out = {n: self.get(n) for n in self._sets[self._cur]}
        out["bc"] = self.get("bc")
        if self.ft is not None:
            out.update({n: self.get(n) for n in ("r_pt", "i_pt", "r_in", "i_in")})
        out["t"] = np.int64(self.t)
        return out

    def restore(self, ckpt: dict) -> None:
        need = list(self._sets[self._cur]) + ["bc"] + (["r_pt", "i_pt", "r_in", "i_in"] if self.ft is not None else [])
        missing = [n for n in need if n not in ckpt]
        if missing:
            raise _lib.FdtdError(f"restore: checkpoint lacks {missing}")
        for n in list(self._sets[self._cur]) + ["bc"]:
            a = np.asarray(ckpt[n])
            if a.shape != tuple(self.tensor(n).shape) or a.dtype != self.np_dtype:
                raise _lib.FdtdError(f"restore: {n} has shape {a.shape} / {a.dtype}")
            self.set(n, a)
        if self.ft is not None:
            for n in ("r_pt", "i_pt", "r_in", "i_in"):
                getattr(self.ft, n).copy_(torch.from_numpy(np.ascontiguousarray(ckpt[n], dtype=self.np_dtype)).reshape(getattr(self.ft, n).shape))
        self.t = int(ckpt["t"])

    def _problem(self) -> _lib.Problem1D:
        p = _lib.Problem1D()
        p.dtype, p.nx = _lib.dtype_code(self.np_dtype), self.nx
        p.flags = (_lib.ABC if self.abc else 0) | (_lib.FLUX if self.form == "flux" else 0) | (_lib.DEBYE if self.debye else 0)
        p.ca = None if self.ca is None else self.ca.data_ptr()
        p.cb = None if self.cb is None else self.cb.data_ptr()
        if self.md is not None:
            p.md = _lib.Medium1D(self.md.nax.data_ptr(), self.md.nbx.data_ptr(),
                                 None if self.md.ncx is None else self.md.ncx.data_ptr(),
                                 None if self.md.ndx is None else self.md.ndx.data_ptr())
        for s in range(2):
            for k, n in enumerate(self.FIELDS):
                t = self._sets[s].get(n)
                p.state[s][k] = None if t is None else t.data_ptr()
            p.bc[s] = self._bc[s].data_ptr()
        if self.source is None:
            p.src_field, p.src_index, p.src_hard = 0, -1, 0
        else:
            p.src_field = 1 if self.source.field == "dx" else 0
            p.src_index, p.src_hard = int(self.source.index), int(self.source.hard)
        return p

    FUSED_DFT_MAX_FREQS = 3      # frequencies the fused pass carries in registers (library limit)

    def advance(self, nsteps: int, tblock: Optional[int] = None, fused_dft: bool = True) -> None:
        if nsteps <= 0:
            return
        if self.ft is not None and (not fused_dft or len(self.freqs) > self.FUSED_DFT_MAX_FREQS):
            # the DFT samples Ex between the E update and the ABC of every step (fd1d_2_2.py:138-142): unfused steps
            for _ in range(int(nsteps)):
                self.step()
            return
        src = None
        if self.source is not None:
            src = np.ascontiguousarray(self.source.waveform.table(self.t + 1, nsteps), dtype=np.float64)
        p = self._problem()
        if self.ft is not None:
            # running DFT carried through the passes: phase factors of every step of this call, evaluated with the
            # reference's expression (numpy programs: products in the array dtype until the np.int32 step counter)
            cos_t, sin_t = _phase_tables(self.freqs, self.dt, self.t + 1, nsteps, True, self.np_dtype)
            D = C.POINTER(C.c_double)
            p.nf, p.dft_sample, p.ft = len(self.freqs), self.DFT_SAMPLE, self.ft.as_struct()
            p.dft_cos, p.dft_sin = cos_t.ctypes.data_as(D), sin_t.ctypes.data_as(D)
        out = C.c_int(-1)
        with torch.cuda.device(self.device):
            check(lib().fdtd1d_advance(C.byref(p), self._cur, int(nsteps),
                                       None if src is None else src.ctypes.data_as(C.POINTER(C.c_double)),
                                       int(tblock or self.tblock), _stream(), C.byref(out)), "fdtd1d_advance")
        self._cur = out.value
        self.t += int(nsteps)

    def step(self) -> None:
        """One step through the reference-named functions in the reference order."""
        t = self.t + 1
        s = self._sets[self._cur]
        bc = self._bc[self._cur] if self.abc else None
        with torch.cuda.device(self.device):
            if self.form == "fdtd":
                exfield(t, self.nx, s["ex"], s["hy"], ca=self.ca, cb=self.cb, source=self.source)
            else:
                dxfield(t, self.nx, s["dx"], s["hy"], source=self.source)
                exfield_flux(self.nx, self.md, s["dx"], s["ix"], s["ex"], s.get("sx"))
            if self.ft is not None:
                fourier(t, len(self.freqs), self.nx, self.dt, self.freqs, s["ex"], self.ft)
            hyfield(self.nx, s["ex"], s["hy"], bc)
        self.t = t

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.device)
