"""2D TM (Dz/Ez/Hx/Hy + PML + TFSF + lossy medium) host side.

Two layers, both calling the sm_100a kernels of ``csrc/libfdtd_b200.so`` through ctypes:

* the reference's module-level step-function protocol, same names and argument order as
  fd2d/program/fd2d_3_3.py:60-110 / fd2d/python/fd2d_3_4.py:104-170 -- ``ezinct, dfield, inctdz, efield,
  hxinct, hfield, incthx, incthy`` -- on CUDA tensors the caller owns, mutated in place;
* :class:`Fdtd2D`, which owns the arrays the reference's ``main()`` allocates (fd2d_3_3.py:132-161) and
  replaces its ``for t in ...`` loop (:166-174) by ``advance(nsteps)``: the fused, temporally blocked kernel.

torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib, surface
from ._lib import check, lib

_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
_NP_DT = {torch.float32: np.float32, torch.float64: np.float64}
FIELD_NAMES = ("dz", "ez", "hx", "hy", "ihx", "ihy", "iz")


def _require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
            raise _lib.FdtdError("arguments must be contiguous CUDA tensors (there is no CPU path)")


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise _lib.FdtdError(f"unsupported tensor dtype {t.dtype}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(None if t is None else t.data_ptr())


class pmlayer(NamedTuple):
    """Device-resident PML vectors, reference field order."""
    fx1: torch.Tensor
    fx2: torch.Tensor
    fx3: torch.Tensor
    fy1: torch.Tensor
    fy2: torch.Tensor
    fy3: torch.Tensor
    gx2: torch.Tensor
    gx3: torch.Tensor
    gy2: torch.Tensor
    gy3: torch.Tensor

    def as_struct(self) -> _lib.PmlLayer:
        return _lib.PmlLayer(*[t.data_ptr() for t in self])


class medium(NamedTuple):
    naz: torch.Tensor
    nbz: Optional[torch.Tensor] = None


class ftrans(NamedTuple):
    """Running-DFT accumulators, reference field order (fd2d/python/fd2d_3_4.py:72-75)."""
    r_pt: torch.Tensor
    i_pt: torch.Tensor
    r_in: torch.Tensor
    i_in: torch.Tensor

    def as_struct(self) -> _lib.FTrans:
        return _lib.FTrans(*[t.data_ptr() for t in self])


def _phases(freq, dt, t, numpy_style, np_dtype):
    """float64 phase factors cos/sin(2*pi*f*dt*t) evaluated the way the reference evaluates them: the numpy
    programs (fd1d_2_2.py:68) multiply in the array dtype until the np.int32 step counter promotes to float64;
    the numba program (fd2d_3_4.py:94) widens freq[n] to float64 first."""
    t = np.int32(t)
    if numpy_style:
        f = np.asarray(freq, dtype=np_dtype).reshape(-1, 1)
        arg = 2 * np.pi * f * dt * t
        return (np.ascontiguousarray(np.cos(arg).astype(np.float64).ravel()),
                np.ascontiguousarray(np.sin(arg).astype(np.float64).ravel()))
    f = np.asarray(freq, dtype=np_dtype)
    arg = np.array([2 * np.pi * np.float64(x) * dt * t for x in f], dtype=np.float64)
    return np.cos(arg), np.sin(arg)


def _phase_tables(freq, dt, t_first, n, numpy_style, np_dtype):
    """``_phases`` of steps ``t_first .. t_first+n-1`` as two contiguous float64 tables [step][frequency].  Evaluated
    in one vectorised expression and probed against the per-step evaluation (numpy's array and scalar cos / sin are
    not guaranteed to round alike); any difference falls back to the per-step loop."""
    n = int(n)
    t = np.arange(t_first, t_first + n).astype(np.int32).reshape(n, 1)
    if numpy_style:
        arg = 2 * np.pi * np.asarray(freq, dtype=np_dtype).reshape(1, -1) * dt * t
        cos_t, sin_t = np.cos(arg).astype(np.float64), np.sin(arg).astype(np.float64)
    else:
        arg = 2 * np.pi * np.asarray(freq, dtype=np_dtype).astype(np.float64).reshape(1, -1) * dt * t
        cos_t, sin_t = np.cos(arg), np.sin(arg)
    probe = sorted(set(range(min(n, 8))) | set(range(max(n - 8, 0), n)) | set(range(0, n, max(1, n // 16))))
    for k in probe:
        c, s = _phases(freq, dt, t_first + k, numpy_style, np_dtype)
        if c.tobytes() != cos_t[k].tobytes() or s.tobytes() != sin_t[k].tobytes():
            ph = [_phases(freq, dt, t_first + j, numpy_style, np_dtype) for j in range(n)]
            cos_t, sin_t = np.stack([c for c, _ in ph]), np.stack([s for _, s in ph])
            break
    return (np.ascontiguousarray(cos_t.reshape(-1), dtype=np.float64),
            np.ascontiguousarray(sin_t.reshape(-1), dtype=np.float64))


def fourier(t: int, nf: int, nx: int, ny: int, dt: float, freq, ezi, ez, ft: ftrans) -> None:
    """Running DFT of Ez and of the source sample ``ezi[6]`` (reference argument order)."""
    _require_cuda(ez, ezi, *ft)
    c, s = _phases(freq, dt, t, False, _NP_DT[ez.dtype])
    fs = ft.as_struct()
    D = C.POINTER(C.c_double)
    check(lib().fdtd2d_fourier(_code(ez), nf, nx, ny, c.ctypes.data_as(D), s.ctypes.data_as(D), _ptr(ezi), 6, _ptr(ez),
                               C.byref(fs), _stream()), "fourier")


def pmlparam(nx: int, ny: int, npml: int, dtype=np.float32, device=None, where: str = "host",
             host_cubes: bool = True) -> pmlayer:
    """The ten PML vectors on the device.  ``where="host"`` (default): the reference formulas evaluated by
    surface.pmlparam and uploaded; ``where="device"``: the library's kernel (``fdtd2d_pmlparam``, float64 evaluation)
    writes the nx- and ny-long vectors, no host arrays (profiles/r2_pmlparam_host_vs_device.txt).  The reference's
    ``x ** 3`` is the host libm's pow(), within one ulp but not correctly rounded, so the 2*npml cubes are taken from the
    host (Python ``**``, as the reference) and the kernel expands them: bit-identical in float32 and float64.
    ``host_cubes=False``: nothing from the host, the kernel cubes in double-double -- float32 still bit-identical
    (every layer of every npml <= 300 tested), float64 to one ulp of the cube."""
    if where == "host":
        host = surface.pmlparam(nx, ny, npml, dtype)
        return pmlayer(*[torch.from_numpy(a).to(device or "cuda") for a in host])
    if where != "device":
        raise ValueError(where)
    if npml < 0 or 2 * npml > min(nx, ny):
        raise ValueError(f"npml={npml} does not fit a {nx}x{ny} grid")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    tdt = _TORCH_DT[np.dtype(dtype)]
    out = pmlayer(*[torch.empty(n, dtype=tdt, device=dev) for n in (nx, nx, nx, ny, ny, ny, nx, nx, ny, ny)])
    ps = out.as_struct()
    cubes = None
    if host_cubes and npml > 0:
        table = [((npml - n) / npml) ** 3 for n in range(npml)] + [((npml - n - 0.5) / npml) ** 3 for n in range(npml)]
        cubes = torch.tensor(table, dtype=torch.float64).to(dev)
    with torch.cuda.device(dev):
        check(lib().fdtd2d_pmlparam(_lib.dtype_code(dtype), int(nx), int(ny), int(npml),
                                    _ptr(cubes) if cubes is not None else None, C.byref(ps), _stream()), "pmlparam")
    return out


def dielectric(nx: int, ny: int, npml: int, rgrid: int, dt: float, epsr: float, sigma: float, dtype=np.float32,
               device=None, rows=None) -> medium:
    """``naz, nbz`` of the lossy dielectric cylinder (reference ``dielectric``, fd2d/python/fd2d_3_4.py:173-194),
    rasterised ON THE DEVICE -- the host version needs minutes at 32768^2.  ``rows=(lo, hi)``: a slab."""
    lo, hi = (0, nx) if rows is None else rows
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    tdt = _TORCH_DT[np.dtype(dtype)]
    naz = torch.empty((hi - lo, ny), dtype=tdt, device=dev)
    nbz = torch.empty((hi - lo, ny), dtype=tdt, device=dev)
    with torch.cuda.device(dev):
        check(lib().fdtd2d_dielectric_cylinder(_lib.dtype_code(dtype), nx, ny, npml, int(rgrid), float(dt), float(epsr),
                                               float(sigma), int(lo), int(hi), _ptr(naz), _ptr(nbz), _stream()),
              "dielectric")
    return medium(naz, nbz)


def plan_depths(nx: int, ny: int, dtype, nsteps: int, tblock: int = 0, rows=None, lossy: bool = False, nf: int = 0):
    """Pass depths the library will use for ``nsteps`` steps on rows ``rows`` (default: the whole grid) of an
    ``nx x ny`` problem -- asked from the library (``fdtd2d_plan``), never re-derived here.  ``tblock = 0``: the library's
    own choice by grid size and step count."""
    p = _lib.Problem2D()
    p.dtype = _lib.dtype_code(dtype)
    p.nx, p.ny = int(nx), int(ny)
    p.row_lo, p.row_hi = (0, int(nx)) if rows is None else (int(rows[0]), int(rows[1]))
    p.flags = _lib.LOSSY if lossy else 0
    p.nf = int(nf)
    return _plan(p, nsteps, tblock)[0]


def _plan(prob: "_lib.Problem2D", nsteps: int, tblock) -> tuple:
    """-> (pass depths, vector width, rows per chunk) of ``fdtd2d_advance(prob, ..., nsteps, ..., tblock)``"""
    nsteps, tb = int(nsteps), int(tblock or 0)
    cap = max(nsteps, 1)
    depths, v, chunk = (C.c_int * cap)(), C.c_int(0), C.c_int(0)
    n = lib().fdtd2d_plan(C.byref(prob), nsteps, tb, depths, cap, C.byref(v), C.byref(chunk))
    if n < 0:
        check(n, "fdtd2d_plan")
    return [int(depths[k]) for k in range(n)], int(v.value), int(chunk.value)


@dataclass(frozen=True)
class PointSource:
    """``dz[i, j] = waveform(t)`` (hard) or ``+=`` (soft) after the D update (fd2d_3_1.py:48, fd2d_3_2.py:66)."""
    i: int
    j: int
    waveform: object
    hard: bool = True


@dataclass(frozen=True)
class IncidentWave:
    """TFSF plane wave: the waveform drives ``ezi[3]`` of the incident line (fd2d_3_3.py:72)."""
    waveform: object


# ------------------------------------------------------------- reference-named step functions (in place)
def _src_struct(target, index, value, hard):
    if target is None:
        return None
    return _lib.Source(target.data_ptr(), int(index), int(bool(hard)), float(value))


def ezinct(ny: int, ezi, hxi, bc) -> None:
    _require_cuda(ezi, hxi, bc)
    check(lib().fdtd2d_ezinct(_code(ezi), ny, _ptr(ezi), _ptr(hxi), _ptr(bc), _stream()), "ezinct")


def dfield(t: int, nx: int, ny: int, *args, source=None, ezi=None) -> None:
    """D update; then the source sample of step ``t``: an :class:`IncidentWave` sets ``ezi[3]``, a
    :class:`PointSource` sets / adds to ``dz[i, j]``.  (The reference hard-codes the waveform here.)

    All three reference argument lists are accepted positionally: ``dfield(t, nx, ny, dz, hx, hy)`` (free space,
    fd2d/program/fd2d_3_1.py:44), ``dfield(t, nx, ny, pml, dz, hx, hy)`` (3_2, fd2d_3_2.py:61) and
    ``dfield(t, nx, ny, pml, ezi, dz, hx, hy)`` (3_3 / 3_4, fd2d_3_3.py:68)."""
    if args and isinstance(args[0], pmlayer):
        pml, arrays = args[0], args[1:]
    else:                                             # program 3_1: no PML -- the identity coefficient set
        pml, arrays = None, args
    if len(arrays) == 4:
        if ezi is not None:
            raise TypeError("dfield: ezi given both positionally and by keyword")
        ezi, dz, hx, hy = arrays
    elif len(arrays) == 3:
        dz, hx, hy = arrays
    else:
        raise TypeError(f"dfield takes (dz, hx, hy) or (ezi, dz, hx, hy) after pml, got {len(arrays)} arrays")
    _require_cuda(dz, hx, hy, ezi)
    if pml is None:
        pml = pmlparam(nx, ny, 0, _NP_DT[dz.dtype], dz.device)
    src = None
    if isinstance(source, IncidentWave):
        src = _src_struct(ezi, 3, source.waveform.table(int(t), 1)[0], True)
    elif isinstance(source, PointSource):
        src = _src_struct(dz, source.i * ny + source.j, source.waveform.table(int(t), 1)[0], source.hard)
    ps = pml.as_struct()
    check(lib().fdtd2d_dfield(_code(dz), nx, ny, C.byref(ps), _ptr(dz), _ptr(hx), _ptr(hy),
                              C.byref(src) if src is not None else None, _stream()), "dfield")


def inctdz(nx: int, ny: int, npml: int, hxi, dz) -> None:
    _require_cuda(hxi, dz)
    check(lib().fdtd2d_inctdz(_code(dz), nx, ny, npml, _ptr(hxi), _ptr(dz), _stream()), "inctdz")


def efield(nx: int, ny: int, md, dz, *rest, iz=None, ez=None) -> None:
    """Reference argument order, both forms: ``efield(nx, ny, naz, dz, ez)`` (programs 3_1-3_3,
    fd2d/program/fd2d_3_3.py:81) and ``efield(nx, ny, md, dz, iz, ez)`` with a lossy :class:`medium` (3_4,
    fd2d/python/fd2d_3_4.py:131).  ``iz`` / ``ez`` may also be given by keyword."""
    if len(rest) == 2:
        if iz is not None or ez is not None:
            raise TypeError("efield: iz / ez given both positionally and by keyword")
        iz, ez = rest
    elif len(rest) == 1:
        if ez is not None:
            raise TypeError("efield: ez given both positionally and by keyword")
        ez = rest[0]
    elif len(rest) != 0:
        raise TypeError(f"efield takes (dz, ez) or (dz, iz, ez) after the medium, got {1 + len(rest)} arrays")
    if ez is None:
        raise TypeError("efield: ez is missing")
    md = md if isinstance(md, medium) else medium(md)
    if (md.nbz is None) != (iz is None):
        raise _lib.FdtdError("efield: a lossy medium (nbz) and iz go together -- efield(nx, ny, md, dz, iz, ez)")
    _require_cuda(md.naz, md.nbz, dz, ez, iz)
    ms = _lib.Medium2D(md.naz.data_ptr(), None if md.nbz is None else md.nbz.data_ptr())
    check(lib().fdtd2d_efield(_code(dz), nx, ny, C.byref(ms), _ptr(dz), _ptr(iz), _ptr(ez), _stream()), "efield")


def hxinct(ny: int, ezi, hxi) -> None:
    _require_cuda(ezi, hxi)
    check(lib().fdtd2d_hxinct(_code(ezi), ny, _ptr(ezi), _ptr(hxi), _stream()), "hxinct")


def hfield(nx: int, ny: int, pml: pmlayer, ez, ihx, ihy, hx, hy) -> None:
    _require_cuda(ez, ihx, ihy, hx, hy)
    ps = pml.as_struct()
    check(lib().fdtd2d_hfield(_code(ez), nx, ny, C.byref(ps), _ptr(ez), _ptr(ihx), _ptr(ihy), _ptr(hx), _ptr(hy),
                              _stream()), "hfield")


def incthx(nx: int, ny: int, npml: int, ezi, hx) -> None:
    _require_cuda(ezi, hx)
    check(lib().fdtd2d_incthx(_code(hx), nx, ny, npml, _ptr(ezi), _ptr(hx), _stream()), "incthx")


def incthy(nx: int, ny: int, npml: int, ezi, hy) -> None:
    _require_cuda(ezi, hy)
    check(lib().fdtd2d_incthy(_code(hy), nx, ny, npml, _ptr(ezi), _ptr(hy), _stream()), "incthy")


# ------------------------------------------------------------------------------------------ Fdtd2D
class Fdtd2D:
    """Owns the device arrays of one 2D TM problem and advances them.

    ``rows=(lo, hi)`` with ``ghost=g`` makes this object one row slab of a larger grid (see slab.py): the
    arrays then hold global rows ``[lo-g, hi+g)`` clipped to the grid and ``advance`` may take at most ``g``
    steps between ghost exchanges.
    """

    def __init__(self, nx: int, ny: int, npml: int = 0, dtype=np.float32, *, source=None, naz=None, nbz=None,
                 device=None, tblock: Optional[int] = None, rows=None, ghost: int = 0, freqs=None, dt: float = surface.DT,
                 ipc: bool = False):
        if not torch.cuda.is_available():
            raise _lib.FdtdError("Fdtd2D needs a CUDA device: the product has no CPU path")
        lib()
        self.nx, self.ny, self.npml = int(nx), int(ny), int(npml)
        self.np_dtype = np.dtype(dtype)
        self.dtype = _TORCH_DT[self.np_dtype]
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.source = source
        self.tfsf = isinstance(source, IncidentWave)
        self.lossy = nbz is not None
        self.row_lo, self.row_hi = (0, self.nx) if rows is None else (int(rows[0]), int(rows[1]))
        self.ghost = int(ghost)
        self.row_base = max(self.row_lo - self.ghost, 0)
        self.rows_alloc = min(self.row_hi + self.ghost, self.nx) - self.row_base
        self.t = 0                                        # steps taken so far (next step is t+1)
        self._lossy_box_cache = None
        self._int_dtype = torch.int32 if self.np_dtype == np.float32 else torch.int64    # bit view: -0.0 is not +0
        code = _lib.dtype_code(self.np_dtype)
        self.max_tblock = lib().fdtd2d_max_tblock(code, self.ny)
        self.tblock = int(tblock) if tblock else 0       # 0: the library picks depth / vector width / chunking by grid size

        with torch.cuda.device(self.device):
            shape = (self.rows_alloc, self.ny)
            names = FIELD_NAMES if self.lossy else FIELD_NAMES[:-1]
            # ipc=True: state arrays come from the library's allocator so that the neighbour ranks can map them
            # (fused halo exchange, slab.py); torch wraps them in place
            self._buffers = []

            def new_state():
                if not ipc:
                    return torch.zeros(shape, dtype=self.dtype, device=self.device)
                buf = _lib.DeviceBuffer(shape, self.np_dtype)
                self._buffers.append(buf)
                return torch.as_tensor(buf, device=self.device)
            self._sets = [{n: new_state() for n in names} for _ in range(2)]
            self._ipc_handles = None if not ipc else [
                {n: self._buffers[s * len(names) + k].ipc_handle() for k, n in enumerate(names)} for s in range(2)]
            self._cur = 0
            self.naz = self._to_dev_rows(naz, fill=1.0)
            self.nbz = self._to_dev_rows(nbz, fill=0.0) if self.lossy else None
            self.pml = pmlparam(self.nx, self.ny, self.npml, self.np_dtype, self.device)
            if self.tfsf:
                z1 = lambda n: torch.zeros(n, dtype=self.dtype, device=self.device)
                self.ezi, self.hxi, self.bc = z1(self.ny), z1(self.ny), z1(4)
                self._ezi_hist = z1(self.max_tblock * self.ny)
                self._hxi_hist = z1(self.max_tblock * 2)
            else:
                self.ezi = self.hxi = self.bc = self._ezi_hist = self._hxi_hist = None
            check(lib().fdtd2d_preload(code, self.ny, int(self.lossy) | (2 if freqs is not None else 0)), "fdtd2d_preload")
            if self.check_identity() != 0:
                raise _lib.FdtdError("PML vectors violate the identity-coefficient promise outside the layer")
            if self.lossy and self.check_lossless_outside() != 0:
                raise _lib.FdtdError("nbz / iz violate the lossless-outside promise computed from them")
            # running DFT (program 3_4): accumulators over the stored rows, updated after every step
            self.freqs, self.dt = (None if freqs is None else np.asarray(freqs, dtype=self.np_dtype)), float(dt)
            if self.freqs is not None:
                nf = len(self.freqs)
                z = lambda *shape: torch.zeros(shape, dtype=self.dtype, device=self.device)
                self.ft = ftrans(z(nf, self.rows_alloc, self.ny), z(nf, self.rows_alloc, self.ny), z(nf), z(nf))
            else:
                self.ft = None

    # ---- array access -----------------------------------------------------------------------------
    def _to_dev_rows(self, host, fill):
        """Upload a coefficient array given for the whole grid, the owned rows, or the stored rows."""
        if host is None:
            return torch.full((self.rows_alloc, self.ny), fill, dtype=self.dtype, device=self.device)
        if isinstance(host, torch.Tensor) and host.is_cuda:           # already on the device (fd2d.dielectric)
            if tuple(host.shape) == (self.nx, self.ny):
                host = host[self.row_base:self.row_base + self.rows_alloc]
            if tuple(host.shape) != (self.rows_alloc, self.ny) or host.dtype != self.dtype:
                raise _lib.FdtdError(f"device coefficient array has shape {tuple(host.shape)} / {host.dtype}")
            return host.to(self.device).contiguous()
        if isinstance(host, torch.Tensor):
            host = host.detach().cpu().numpy()
        a = np.asarray(host, dtype=self.np_dtype)
        if a.shape == (self.nx, self.ny):
            a = a[self.row_base:self.row_base + self.rows_alloc]
        if a.shape != (self.rows_alloc, self.ny):
            raise _lib.FdtdError(f"coefficient array has shape {a.shape}, expected {(self.rows_alloc, self.ny)}")
        t = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        return t if t.data_ptr() % 16 == 0 else t.clone()      # (a row slice of a host array shared without a copy)

    def _owned(self, t: torch.Tensor) -> torch.Tensor:
        o = self.row_lo - self.row_base
        return t[o:o + (self.row_hi - self.row_lo)]

    def tensor(self, name: str, stored: bool = False) -> torch.Tensor:
        """Current device tensor of a field (owned rows; ``stored=True`` includes ghost rows)."""
        t = self._sets[self._cur][name]
        return t if stored else self._owned(t)

    def get(self, name: str) -> np.ndarray:
        if name in ("ezi", "hxi", "bc"):
            return getattr(self, name).cpu().numpy()
        if name in ("r_pt", "i_pt"):
            o = self.row_lo - self.row_base
            return getattr(self.ft, name)[:, o:o + (self.row_hi - self.row_lo)].cpu().numpy()
        if name in ("r_in", "i_in"):
            return getattr(self.ft, name).cpu().numpy()
        return self.tensor(name).cpu().numpy()

    def set(self, name: str, host) -> None:
        """Upload a field (owned-rows shape or whole-grid shape).  Ghost rows are filled from a whole-grid array."""
        if name in ("ezi", "hxi", "bc"):
            getattr(self, name).copy_(torch.as_tensor(np.asarray(host, dtype=self.np_dtype)))
            return
        if name == "iz":
            self._lossy_box_cache = None
        a = np.asarray(host, dtype=self.np_dtype)
        if a.shape == (self.nx, self.ny) and (self.rows_alloc != self.nx):
            self.tensor(name, stored=True).copy_(torch.from_numpy(
                np.ascontiguousarray(a[self.row_base:self.row_base + self.rows_alloc])))
        else:
            self.tensor(name).copy_(torch.from_numpy(np.ascontiguousarray(a)))

    # ---- the fused path -----------------------------------------------------------------------------
    def _problem(self, lossy_box=None) -> _lib.Problem2D:
        p = _lib.Problem2D()
        p.dtype = _lib.dtype_code(self.np_dtype)
        p.nx, p.ny = self.nx, self.ny
        p.row_lo, p.row_hi, p.row_base, p.rows_alloc = self.row_lo, self.row_hi, self.row_base, self.rows_alloc
        p.npml = self.npml
        p.flags = (_lib.TFSF if self.tfsf else 0) | (_lib.LOSSY if self.lossy else 0)
        p.pml = self.pml.as_struct()
        p.md = _lib.Medium2D(self.naz.data_ptr(), None if self.nbz is None else self.nbz.data_ptr())
        for s in range(2):
            for k, n in enumerate(FIELD_NAMES):
                t = self._sets[s].get(n)
                p.state[s][k] = None if t is None else t.data_ptr()
        if self.tfsf:
            p.ezi, p.hxi, p.bc = self.ezi.data_ptr(), self.hxi.data_ptr(), self.bc.data_ptr()
            p.ezi_hist, p.hxi_hist = self._ezi_hist.data_ptr(), self._hxi_hist.data_ptr()
        if isinstance(self.source, PointSource):
            p.src_i, p.src_j, p.src_hard = self.source.i, self.source.j, int(self.source.hard)
        else:
            p.src_i, p.src_j, p.src_hard = -1, -1, 1
        # pmlparam leaves every coefficient at its identity value on [npml, N-1-npml): promise it to the kernel
        p.ident_row_lo, p.ident_row_hi = self._ident(self.nx)
        p.ident_col_lo, p.ident_col_hi = self._ident(self.ny)
        if self.lossy:
            p.lossy_row_lo, p.lossy_row_hi, p.lossy_col_lo, p.lossy_col_hi = lossy_box if lossy_box is not None else self._lossy_box()
        return p

    def _lossy_box(self):
        """Global rows / columns ``(r0, r1, c0, c1)`` outside which ``nbz == 0`` and ``iz == +0`` in both state sets (the
        lossless-outside promise of fdtd2d_problem), found on the device from the arrays themselves; recomputed after
        ``iz`` or ``nbz`` was written from outside (``set``, ``restore``, :meth:`invalidate_lossy_box`)."""
        if self._lossy_box_cache is None:
            with torch.cuda.device(self.device):
                live = (self.nbz != 0) | (self._sets[0]["iz"].view(self._int_dtype) != 0) | (self._sets[1]["iz"].view(self._int_dtype) != 0)
                rows = torch.nonzero(live.any(dim=1)).flatten()
                cols = torch.nonzero(live.any(dim=0)).flatten()
                if rows.numel() == 0:
                    box = (self.row_base, self.row_base + 1, 0, 1)          # nothing lossy here: a one-cell box at a corner
                else:
                    box = (self.row_base + int(rows[0]), self.row_base + int(rows[-1]) + 1, int(cols[0]), int(cols[-1]) + 1)
            self._lossy_box_cache = box
        return self._lossy_box_cache

    def invalidate_lossy_box(self) -> None:
        """Call after writing ``nbz`` or ``iz`` tensors directly (``tensor('iz')[...] = ...``)."""
        self._lossy_box_cache = None

    def check_lossless_outside(self) -> int:
        """Device-side verification of the lossless-outside promise (0 = holds)."""
        bad = C.c_longlong(-1)
        p = self._problem()
        with torch.cuda.device(self.device):
            check(lib().fdtd2d_check_lossless_outside(C.byref(p), C.byref(bad)), "fdtd2d_check_lossless_outside")
        return int(bad.value)

    def _ident(self, n):
        lo, hi = self.npml, n - 1 - self.npml
        return (lo, hi) if hi > lo else (0, 0)

    def check_identity(self) -> int:
        """Device-side verification of the identity-coefficient promise (0 = holds)."""
        bad = C.c_longlong(-1)
        p = self._problem()
        with torch.cuda.device(self.device):
            check(lib().fdtd2d_check_identity(C.byref(p), C.byref(bad)), "fdtd2d_check_identity")
        return int(bad.value)

    def advance(self, nsteps: int, tblock: Optional[int] = None, lazy_ez: bool = False, epoch: Optional[int] = None) -> None:
        """``nsteps`` full time steps through the fused, temporally blocked kernel (asynchronous).

        ``ez`` is an output only, so just the last pass stores it; ``lazy_ez=True`` skips even that (used by
        the slab driver between ghost exchanges) and leaves ``ez`` stale until a later non-lazy ``advance``."""
        if nsteps <= 0:
            return
        if self.ft is not None and len(self.freqs) > 3:
            # more frequencies than the fused kernels carry: one fused single-step pass WITHOUT the accumulators
            # attached, then the fourier kernel -- which alone updates them -- per step
            for _ in range(int(nsteps)):
                self._advance_fused(1, 1, False, None, carry_dft=False)
                self._fourier(self.t)
            return
        self._advance_fused(int(nsteps), tblock, lazy_ez, epoch)

    def _fourier(self, t: int) -> None:
        with torch.cuda.device(self.device):
            fourier(t, len(self.freqs), self.rows_alloc, self.ny, self.dt, self.freqs, self.ezi,
                    self.tensor("ez", stored=True), self.ft)

    def _advance_fused(self, nsteps: int, tblock, lazy_ez: bool, epoch: Optional[int] = None, carry_dft: bool = True) -> None:
        tb = int(tblock if tblock is not None else self.tblock)
        if tb > self.max_tblock:
            raise _lib.FdtdError(f"tblock {tb} exceeds the deepest supported time block {self.max_tblock}")
        src = None
        if self.source is not None:
            src = np.ascontiguousarray(self.source.waveform.table(self.t + 1, nsteps), dtype=np.float64)
        p = self._problem()
        if lazy_ez:
            p.flags |= _lib.LAZY_EZ
        if epoch is not None and getattr(self, "p2p", None) is not None:
            # halo exchange fused into the pass: peer-mapped neighbour arrays + sync words (see slab.py)
            x = self.p2p
            p.halo, p.epoch = int(x["halo"]), int(epoch)
            p.sync_local = x["sync"].ptr
            for side, key in (("up", "peer_up"), ("dn", "peer_dn")):
                nb = x[side]
                if nb is None:
                    continue
                for s in range(2):
                    for k, n in enumerate(FIELD_NAMES):
                        getattr(p, key)[s][k] = nb["sets"][s].get(n)          # peer-mapped raw pointers
                setattr(p, key + "_base", int(nb["row_base"]))
                setattr(p, "sync_" + side, nb["sync"])
        if self.ft is not None and carry_dft:
            # running DFT fused into the passes: per-step phase factors, evaluated as the reference evaluates them
            nf = len(self.freqs)
            cos_t, sin_t = _phase_tables(self.freqs, self.dt, self.t + 1, nsteps, False, self.np_dtype)
            D = C.POINTER(C.c_double)
            p.nf, p.ft = nf, self.ft.as_struct()
            if not self.tfsf:
                p.ft.r_in = p.ft.i_in = None
            p.dft_cos, p.dft_sin = cos_t.ctypes.data_as(D), sin_t.ctypes.data_as(D)
        out = C.c_int(-1)
        with torch.cuda.device(self.device):
            check(lib().fdtd2d_advance(C.byref(p), self._cur, int(nsteps),
                                       None if src is None else src.ctypes.data_as(C.POINTER(C.c_double)),
                                       tb, _stream(), C.byref(out)), "fdtd2d_advance")
        self._cur = out.value
        self.t += int(nsteps)

    # ---- streamed run: host medium in, host Ez out, PCIe overlapped with the time stepping ----------------
    def pass_depths(self, nsteps: int, tblock=None, rows=None):
        """Pass depths ``advance(nsteps, tblock)`` will use, asked from the library (``fdtd2d_plan``): it chooses them by
        the rows a call produces (``rows``: one block of a streamed run), the step count, dtype and features."""
        p = _lib.Problem2D()
        p.dtype = _lib.dtype_code(self.np_dtype)
        p.nx, p.ny = self.nx, self.ny
        p.row_lo, p.row_hi = (self.row_lo, self.row_hi) if rows is None else (int(rows[0]), int(rows[1]))
        p.flags = _lib.LOSSY if self.lossy else 0
        p.nf = 0 if self.ft is None else min(len(self.freqs), 3)
        return _plan(p, nsteps, tblock if tblock is not None else self.tblock)[0]

    _depths = pass_depths

    def run_streamed(self, nsteps: int, naz_host: torch.Tensor, ez_host: torch.Tensor, blocks: Optional[int] = None,
                     tblock=None, streams: int = 16, trace: Optional[list] = None, block_rows=None,
                     schedule: str = "skewed", window: Optional[int] = 2, priorities: bool = False,
                     nbz_host: Optional[torch.Tensor] = None) -> None:
        """The whole job a reference ``main()`` does -- medium from the host, ``nsteps`` steps from zero fields,
        Ez back on the host -- with the PCIe transfers hidden behind the kernels.

        The grid is cut into row blocks (``block_rows``: one height or a list of heights top to bottom; or ``blocks``
        equal blocks) that are uploaded in order; every pass (``depth`` time steps) is issued block by block, each pass
        level on its own stream, so the levels form a systolic pipeline whose small launches overlap.

        ``schedule="skewed"`` (default): pass level p works on the block boundaries shifted UP by (p+1)*depth rows
        (parallelogram tiling in space-time).  Level p of block b then reads level p-1 only over rows that level p-1
        produced for blocks b and b-1 -- never for b+1 -- so a block runs through ALL its passes as soon as it has
        arrived, whatever its height; the last block grows by the shift and the first shrinks.  The ping-pong sets stay
        safe: what level p of block b overwrites was read by level p-1 of blocks <= b only (its dependencies).
        ``window`` blocks are in flight at most (block b+window starts once block b has finished its last pass): blocks
        then finish in order and evenly spaced, so their Ez leaves over PCIe while later blocks still step -- without
        it all 16 pass-level streams share the GPU, every block finishes near the end and the downloads trail the
        stepping (32768^2 x 96 steps: 157 ms without a window, 142 ms with 2; profiles/r1_streamed_schedules_k96.txt).
        ``priorities=True`` gives later pass levels higher stream priority instead (145 ms).
        ``schedule="wavefront"``: unshifted blocks; level p of block b must wait for level p-1 of block b+1, i.e. for
        the upload of block b+p+1 -- tall blocks starve the early passes (kept for comparison and for runs whose total
        shift would not fit the first block).

        Same kernels, same arithmetic, same result as ``set naz; advance(nsteps); get ez``.  ``naz_host``: pinned CPU
        tensor over the stored rows, ``ez_host`` over the owned rows (both (nx, ny) on a single device); a lossy
        medium also brings ``nbz_host`` (stored rows), uploaded with the same blocks.  With a TFSF source the incident
        line is advanced ONCE per pass level up front (``fdtd2d_incident_line``: one tiny launch per level, its history
        kept per level) and every block's pass reads that history (``FDTD_INCIDENT_READY``) -- the reference ``main()``
        this replaces is fd2d/python/fd2d_3_4.py:211-293.  No running DFT.

        On a slab (``rows=``, ``ghost=g``) the run is communication-avoiding: at most g steps, during which the
        ghost band is consumed one row per step instead of being exchanged (``FDTD_GHOST_DECAY``) -- the owned rows come
        out exact, the ghost rows must be refreshed before stepping on (``SlabFdtd2D.run_streamed`` does)."""
        if self.ft is not None:
            raise _lib.FdtdError("run_streamed: no running DFT (its accumulators would have to stream too)")
        if self.lossy != (nbz_host is not None):
            raise _lib.FdtdError("run_streamed: a lossy medium brings nbz_host (and only a lossy medium does)")
        if nbz_host is not None and tuple(nbz_host.shape) != (self.rows_alloc, self.ny):
            raise _lib.FdtdError("run_streamed: nbz_host must cover the stored rows (rows_alloc, ny)")
        slab = self.rows_alloc != self.nx
        rows_own = self.row_hi - self.row_lo
        if slab and int(nsteps) > self.ghost:
            raise _lib.FdtdError(f"run_streamed on a slab: {nsteps} steps without an exchange need {nsteps} ghost rows, have {self.ghost}")
        if tuple(naz_host.shape) != (self.rows_alloc, self.ny) or tuple(ez_host.shape) != (rows_own, self.ny):
            raise _lib.FdtdError("run_streamed: naz_host must cover the stored rows (rows_alloc, ny), ez_host the owned rows")
        plan = self.streamed_plan(nsteps, tblock=tblock, streams=streams, blocks=blocks, block_rows=block_rows,
                                  schedule=schedule, window=window)
        depths, edges, P, S = plan["depths"], plan["edges"], plan["levels"], plan["streams"]
        B = len(edges) - 1
        lo_all = self.row_base
        src = None
        if self.source is not None:
            src = np.ascontiguousarray(self.source.waveform.table(self.t + 1, nsteps), dtype=np.float64)
        first_step = np.concatenate(([0], np.cumsum(depths)))          # step offset of every pass
        D = C.POINTER(C.c_double)
        with torch.cuda.device(self.device):
            caller = torch.cuda.current_stream()
            if getattr(self, "_lanes", None) is None or len(self._lanes) < S + 2 or getattr(self, "_lanes_prio", False) != bool(priorities):
                if priorities:
                    # later pass levels first: work close to completion overtakes fresh blocks, so blocks finish in
                    # order and their Ez leaves while the rest still steps
                    least_p, greatest_p = torch.cuda.Stream.priority_range()
                    span = max(1, least_p - greatest_p)
                    self._lanes = [torch.cuda.Stream(priority=least_p - min(span, (k * (span + 1)) // S)) for k in range(S)] \
                        + [torch.cuda.Stream(), torch.cuda.Stream()]
                else:
                    self._lanes = [torch.cuda.Stream() for _ in range(S + 2)]
                self._lanes_prio = bool(priorities)
            lanes, up, down = self._lanes[:S], self._lanes[S], self._lanes[S + 1]
            for st in lanes + [up, down]:
                st.wait_stream(caller)
            # nbz is still arriving block by block: no lossless-outside promise can be derived from it (the whole grid is
            # "the box": every interior warp runs the lossy kernel)
            whole = (0, self.nx, 0, self.ny) if self.lossy else None
            hist = None
            if self.tfsf:
                # the incident line does not depend on the 2D grid: its history for every pass level, computed once on
                # the caller's stream before any block starts (the lanes wait for the caller above... and again here)
                dmax = max(depths)
                if getattr(self, "_level_hist", None) is None or self._level_hist[0].shape[0] < P or self._level_hist[0].shape[1] < dmax * self.ny:
                    self._level_hist = (torch.zeros((P, dmax * self.ny), dtype=self.dtype, device=self.device),
                                        torch.zeros((P, dmax * 2), dtype=self.dtype, device=self.device))
                hist = self._level_hist
                base = self._problem(lossy_box=whole)
                for p_idx in range(P):
                    k0 = int(first_step[p_idx])
                    check(lib().fdtd2d_incident_line(C.byref(base), depths[p_idx], src[k0:].ctypes.data_as(D),
                                                     C.c_void_p(hist[0][p_idx].data_ptr()), C.c_void_p(hist[1][p_idx].data_ptr()),
                                                     C.c_void_p(caller.cuda_stream)), "fdtd2d_incident_line")
                for st in lanes:
                    st.wait_stream(caller)
            uploaded = []
            with torch.cuda.stream(up):
                for b in range(B):
                    a, z = edges[b] - lo_all, edges[b + 1] - lo_all
                    self.naz[a:z].copy_(naz_host[a:z], non_blocking=True)
                    if nbz_host is not None:
                        self.nbz[a:z].copy_(nbz_host[a:z], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(up)
                    uploaded.append(ev)
            cur0 = self._cur
            done = {}                                                      # (b, p) -> event
            for item in plan["items"]:
                b, p_idx = item["block"], item["level"]
                lane = lanes[item["lane"]]
                for kind, key in item["waits"]:
                    lane.wait_event(uploaded[key] if kind == "upload" else done[key])
                prob = self._problem(lossy_box=whole)
                prob.row_lo, prob.row_hi = item["rows"]
                if p_idx < P - 1:
                    prob.flags |= _lib.LAZY_EZ
                if slab:
                    prob.flags |= _lib.GHOST_DECAY          # the ghost band is consumed instead of exchanged
                if hist is not None:
                    prob.flags |= _lib.INCIDENT_READY       # this level's incident-line history is already there
                    prob.ezi_hist, prob.hxi_hist = hist[0][p_idx].data_ptr(), hist[1][p_idx].data_ptr()
                out = C.c_int(-1)
                k0 = int(first_step[p_idx])
                if trace is not None:                                  # timeline probe (tools/probe_streamed.py)
                    t0 = torch.cuda.Event(enable_timing=True)
                    t0.record(lane)
                check(lib().fdtd2d_advance(C.byref(prob), (cur0 + p_idx) % 2, depths[p_idx],
                                           None if src is None else src[k0:].ctypes.data_as(D),
                                           depths[p_idx], C.c_void_p(lane.cuda_stream), C.byref(out)),
                      "fdtd2d_advance (streamed)")
                ev = torch.cuda.Event(enable_timing=trace is not None)
                ev.record(lane)
                done[(b, p_idx)] = ev
                if trace is not None:
                    trace.append((b, p_idx, t0, ev))
                if p_idx == P - 1:
                    down.wait_event(ev)
                    a, z = max(item["rows"][0], self.row_lo), min(item["rows"][1], self.row_hi)   # owned rows of this block's last pass
                    if z > a:
                        with torch.cuda.stream(down):
                            ez_dev = self._sets[(cur0 + P) % 2]["ez"]
                            ez_host[a - self.row_lo:z - self.row_lo].copy_(ez_dev[a - lo_all:z - lo_all], non_blocking=True)
            for st in lanes + [up, down]:
                caller.wait_stream(st)
        self._cur = (cur0 + P) % 2
        self.t += int(nsteps)
        self._lossy_box_cache = None                # nbz was replaced

    def streamed_plan(self, nsteps: int, tblock=None, streams: int = 16, blocks: Optional[int] = None, block_rows=None,
                      schedule: str = "skewed", window: Optional[int] = 2) -> dict:
        """The launch plan of :meth:`run_streamed` as plain data (no device work): pass depths, block edges (global rows)
        and, in issue order, one item per (block, pass level) with the rows it produces, its stream and the events it
        waits for -- ``("upload", k)``: block k of the medium has arrived, ``("item", (b, p))``: that item has finished.
        Items of one stream run in issue order.  tests/test_streamed_plan.py proves on this data that any two items not
        ordered by these rules touch disjoint rows of every array set."""
        if schedule not in ("skewed", "wavefront"):
            raise ValueError(schedule)
        depths = self._depths(nsteps, tblock)
        P = len(depths)
        S = max(1, min(int(streams), P))
        dmax = max(depths)
        least = 4 * dmax                                                           # shortest block a pass may be given
        lo_all, hi_all = self.row_base, self.row_base + self.rows_alloc           # global rows stored here
        skew = P * dmax if schedule == "skewed" else 0                             # shift of the last pass level
        if skew and skew + least > self.rows_alloc // 2:
            schedule, skew = "wavefront", 0                                        # a long run on a short grid
        if blocks is None and block_rows is None:
            block_rows = self._default_block_rows(schedule, least)
        if isinstance(block_rows, (list, tuple)):        # explicit block heights, top to bottom (the last one is stretched
            edges = [lo_all]                             # or cut to end at the last stored row)
            for h in block_rows:
                if hi_all - edges[-1] < 2 * least:
                    break
                edges.append(min(edges[-1] + max(int(h), least), hi_all - least))
            if len(edges) == 1:
                edges.append(hi_all)                     # a grid too short to cut: one block
            edges[-1] = hi_all
        elif block_rows:                                 # one block height (the last block takes the remainder)
            edges = list(range(lo_all, hi_all, max(int(block_rows), least))) + [hi_all]
            if len(edges) > 2 and edges[-1] - edges[-2] < least:
                del edges[-2]                            # a remainder too short for a pass joins the block above
        else:
            B = max(1, min(int(blocks), self.rows_alloc // max(least, 1)))
            edges = [lo_all + self.rows_alloc * k // B for k in range(B + 1)]
        if skew:
            # the first block must keep `least` rows after the deepest shift: merge leading blocks until it does
            while len(edges) > 2 and edges[1] - lo_all < skew + least:
                del edges[1]
            if edges[1] - lo_all < skew + least:
                schedule, skew = "wavefront", 0
        B = len(edges) - 1

        def rows_of(b, p):
            """global rows pass level p produces for block b"""
            sh = (p + 1) * dmax if skew else 0
            return (lo_all if b == 0 else edges[b] - sh), (hi_all if b == B - 1 else edges[b + 1] - sh)

        if skew:                                          # block by block, every level of a block in a row
            order = [(b, p) for b in range(B) for p in range(P)]
        else:                                             # by wave, increasing p inside
            order = [(w - p, p) for w in range(B + P - 1) for p in range(P) if 0 <= w - p < B]
        items = []
        for b, p in order:
            if skew:
                # needs level p-1 of blocks b and b-1: (b, p-1) implies (b-1, p-1) -- same stream, earlier block.
                # Level 0 reads naz up to the end of block b, and nothing of block b+1.
                waits = [("upload", b)] if p == 0 else [("item", (b, p - 1))]
                if p == 0 and window and b - int(window) >= 0:
                    waits.append(("item", (b - int(window), P - 1)))       # at most `window` blocks in flight
            elif p == 0:
                waits = [("upload", min(b + 1, B - 1))]                    # the pass reads naz up to depth rows below
            else:
                waits = [("item", (min(b + 1, B - 1), p - 1))]             # (b+1, p-1) implies (b, p-1) and (b-1, p-1)
            items.append({"block": b, "level": p, "rows": rows_of(b, p), "lane": p % S, "waits": waits})
        return {"schedule": schedule, "depths": depths, "levels": P, "streams": S, "edges": edges, "items": items,
                "stored_rows": (lo_all, hi_all)}

    def _default_block_rows(self, schedule: str, least: int):
        """Block plan of run_streamed (profiles/r1_streamed_schedules_k96.txt, 32768^2 x 96 steps).  Wavefront: 1024-row
        blocks (tall blocks starve the early pass levels).  Skewed: short blocks first, so stepping starts after ~1 ms
        of upload, ~3072-row blocks in the middle (launches of several waves), short blocks last, so little is left to
        download once the last pass ends."""
        rows = self.rows_alloc
        if rows < 8192:
            return max(least, -(-rows // 8))
        if schedule == "wavefront" or rows < 16384:
            return 1024
        head, tail = [512, 1024, 2048], [2048, 1024, 512]
        left = rows - sum(head) - sum(tail)
        nb = max(1, round(left / 3072))
        return head + [left // nb] * nb + tail

    # ---- checkpoint / restore, snapshots --------------------------------------------------
    def checkpoint(self) -> dict:
        """Everything ``advance`` carries from step to step, as host arrays: the state arrays (owned rows), the
        incident line, the running-DFT accumulators and the step counter.  The reference keeps this state in
        local arrays of ``main()`` and has no checkpointing; ``restore`` on a problem built with the
        same setup continues bit-identically."""
        names = FIELD_NAMES if self.lossy else FIELD_NAMES[:-1]
        out = {n: self.get(n) for n in names}
        if self.tfsf:
            out.update({n: self.get(n) for n in ("ezi", "hxi", "bc")})
        if self.ft is not None:
            out.update({n: self.get(n) for n in ("r_pt", "i_pt", "r_in", "i_in")})
        out["t"] = np.int64(self.t)
        return out

    def restore(self, ckpt: dict) -> None:
        """Inverse of :meth:`checkpoint` (same grid, rows, dtype and features)."""
        names = FIELD_NAMES if self.lossy else FIELD_NAMES[:-1]
        need = list(names) + (["ezi", "hxi", "bc"] if self.tfsf else []) + \
            (["r_pt", "i_pt", "r_in", "i_in"] if self.ft is not None else [])
        missing = [n for n in need if n not in ckpt]
        if missing:
            raise _lib.FdtdError(f"restore: checkpoint lacks {missing}")
        rows = self.row_hi - self.row_lo
        for n in names:
            a = np.asarray(ckpt[n])
            if a.shape != (rows, self.ny) or a.dtype != self.np_dtype:
                raise _lib.FdtdError(f"restore: {n} has shape {a.shape} / {a.dtype}, expected {(rows, self.ny)} / {self.np_dtype}")
            self.set(n, a)
        if self.tfsf:
            for n in ("ezi", "hxi", "bc"):
                self.set(n, ckpt[n])
        if self.ft is not None:
            o = self.row_lo - self.row_base
            for n in ("r_pt", "i_pt"):
                getattr(self.ft, n)[:, o:o + rows].copy_(torch.from_numpy(np.ascontiguousarray(ckpt[n], dtype=self.np_dtype)))
            for n in ("r_in", "i_in"):
                getattr(self.ft, n).copy_(torch.from_numpy(np.ascontiguousarray(ckpt[n], dtype=self.np_dtype)))
        self.t = int(ckpt["t"])

    def advance_with_snapshots(self, nsteps: int, every: int, out_host: Optional[torch.Tensor] = None,
                               tblock: Optional[int] = None) -> torch.Tensor:
        """``nsteps`` steps, keeping Ez (owned rows) after every ``every``-th step in pinned host memory -- what the
        reference's animation scripts do with ``ez.copy()`` per frame (fd2d/animation/fd2d_3_3.py:156-166) -- with the
        device-to-host copies overlapped with the following steps: each frame is first copied on the device into one
        of two staging buffers (stream-ordered, so the passes never race with it) and leaves over PCIe on a second
        stream while the next ``every`` steps run.  Returns the ``(nsteps // every, rows, ny)`` pinned tensor; call
        :meth:`synchronize` (or sync the current stream) before reading it.  Steps beyond the last frame are taken too."""
        every, nsteps = int(every), int(nsteps)
        if every <= 0:
            raise _lib.FdtdError("advance_with_snapshots: every must be positive")
        rows, frames = self.row_hi - self.row_lo, nsteps // every
        if out_host is None:
            out_host = torch.empty((frames, rows, self.ny), dtype=self.dtype).pin_memory()
        if tuple(out_host.shape) != (frames, rows, self.ny) or out_host.dtype != self.dtype or not out_host.is_pinned():
            raise _lib.FdtdError("advance_with_snapshots: out_host must be a pinned (nsteps // every, rows, ny) tensor of the field dtype")
        with torch.cuda.device(self.device):
            compute = torch.cuda.current_stream()
            if getattr(self, "_snap", None) is None:
                self._snap = {"stream": torch.cuda.Stream(),
                              "stage": [torch.empty((rows, self.ny), dtype=self.dtype, device=self.device) for _ in range(2)],
                              "free": [None, None]}
            snap = self._snap
            for k in range(frames):
                self.advance(every, tblock=tblock)
                b = k & 1
                if snap["free"][b] is not None:
                    compute.wait_event(snap["free"][b])          # frame k-2 has left this staging buffer
                snap["stage"][b].copy_(self.tensor("ez"), non_blocking=True)
                staged = torch.cuda.Event()
                staged.record(compute)
                snap["stream"].wait_event(staged)
                with torch.cuda.stream(snap["stream"]):
                    out_host[k].copy_(snap["stage"][b], non_blocking=True)
                    snap["free"][b] = torch.cuda.Event()
                    snap["free"][b].record(snap["stream"])
            self.advance(nsteps - frames * every, tblock=tblock)
            compute.wait_stream(snap["stream"])                  # the frames are complete once the caller's stream is
        return out_host

    # ---- the unfused path: the reference loop body, one kernel per reference function ----------------
    def step(self) -> None:
        """One time step via the reference-named functions in the reference order (single device only)."""
        if self.rows_alloc != self.nx:
            raise _lib.FdtdError("step() runs the whole-grid reference functions; use advance() on a slab")
        t = self.t + 1
        nx, ny, n = self.nx, self.ny, self.npml
        s = self._sets[self._cur]
        with torch.cuda.device(self.device):
            if self.tfsf:
                ezinct(ny, self.ezi, self.hxi, self.bc)
            dfield(t, nx, ny, self.pml, s["dz"], s["hx"], s["hy"], source=self.source, ezi=self.ezi)
            if self.tfsf:
                inctdz(nx, ny, n, self.hxi, s["dz"])
            if self.lossy:
                efield(nx, ny, medium(self.naz, self.nbz), s["dz"], s["iz"], s["ez"])
            else:
                efield(nx, ny, self.naz, s["dz"], s["ez"])
            if self.ft is not None:
                fourier(t, len(self.freqs), nx, ny, self.dt, self.freqs, self.ezi, s["ez"], self.ft)
            if self.tfsf:
                hxinct(ny, self.ezi, self.hxi)
            hfield(nx, ny, self.pml, s["ez"], s["ihx"], s["ihy"], s["hx"], s["hy"])
            if self.tfsf:
                incthx(nx, ny, n, self.ezi, s["hx"])
                incthy(nx, ny, n, self.ezi, s["hy"])
        self.t = t

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.device)

    def state_bytes(self) -> int:
        per = self.rows_alloc * self.ny * self.np_dtype.itemsize
        return per * (2 * len(self._sets[0]) + 1 + (1 if self.lossy else 0))
