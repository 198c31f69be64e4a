"""CPU oracle for the FDTD hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``simulation_b200/``
imports it, and the product has no CPU fallback.

What it is: a numpy restatement of the Yee time-stepping of dsarvan/simulation for
the 1D (Ex/Hy, flux form Dx/Ex/Ix/Sx/Hy) and 2D TM (Dz/Ez/Hx/Hy + PML + TFSF +
lossy medium + running DFT) programs.  The restatement is *general* (one 1D stepper
and one 2D stepper driven by coefficient arrays and a host-computed source table)
while the reference is one closed script per book example; the restatement is
pinned bit-for-bit against every reference program in ``tests/test_oracle_pinning.py``
(executed against ``/root/reference`` in the build container) and against the
committed goldens in ``tests/golden/`` everywhere else.

Parity status: PINNED (bitwise vs. the reference numpy programs fd1d_1_1..2_3,
fd2d_3_1..3_3 in fp64 and fp32; <=1e-12 relative vs. the numba-only program
fd2d_3_4, which is compiled with fastmath and is not bit-stable itself).

Evaluation rules that make the match bitwise:
  * every array op is evaluated left to right in the array dtype, one rounding per
    op, no fused multiply-add;
  * source waveforms are float64 scalars; a hard source is rounded on assignment, a
    soft source is ADDED IN float64 and then rounded into the array dtype (numpy
    NEP-50: array-element + np.float64 scalar promotes to float64);
  * ``0.5 * x`` is exact, so ``0.5*(a-b)`` and ``0.5*a-0.5*b`` coincide.

Reference lines followed (relative to /root/reference):
  1D FDTD form        fd1d/program/fd1d_1_1.py:39-45, fd1d_1_2.py:41-50,
                      fd1d_1_3.py:53-63, fd1d_1_5.py:37-44,63-72
  1D flux form        fd1d/program/fd1d_2_1.py:43-72, fd1d_2_3.py:82-109
  1D running DFT      fd1d/program/fd1d_2_2.py:65-71,145-146
  2D free space       fd2d/program/fd2d_3_1.py:44-59
  2D PML              fd2d/program/fd2d_3_2.py:61-93
  2D PML + TFSF       fd2d/program/fd2d_3_3.py:60-122
  2D lossy + DFT      fd2d/python/fd2d_3_4.py:89-99,131-136,173-194,279-284
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import sqrt

import numpy as np

EPS0 = 8.854e-12   # vacuum permittivity used by every reference dielectric()
C_TIMES_2 = 6e8    # dt = ds / 6e8  (Courant number 0.5)


# --------------------------------------------------------------------------- sources
def gaussian_pulse(t, t0, spread):
    """fd1d/program/fd1d_1_1.py:27-28 -- ``t`` is an np.int32, result is np.float64."""
    return np.exp(-0.5 * ((t - t0) / spread) ** 2)


def sine_wave(t, ds, freq):
    """fd1d/program/fd1d_1_4.py:30-32."""
    dt = ds / C_TIMES_2
    return np.sin(2 * np.pi * freq * dt * t)


def step_indices(ns, t_first=1):
    """The reference loop variable: ``np.arange(1, ns+1).astype(np.int32)``."""
    return np.arange(t_first, t_first + ns).astype(np.int32)


def source_table(kind, ns, t_first=1, **kw):
    """float64 waveform samples for steps t_first .. t_first+ns-1 (entry k <-> step t_first+k)."""
    out = np.zeros(ns, dtype=np.float64)
    for k, t in enumerate(step_indices(ns, t_first)):
        if kind == "gaussian":
            out[k] = gaussian_pulse(t, kw["t0"], kw["spread"])
        elif kind == "sine":
            out[k] = sine_wave(t, kw.get("ds", 0.01), kw["freq"])
        elif kind == "none":
            out[k] = 0.0
        else:
            raise ValueError(kind)
    return out


def _inject(arr, idx, value, hard):
    """Hard source: overwrite (rounded on store).  Soft source: float64 add, then round."""
    if hard:
        arr[idx] = value
    else:
        arr[idx] = np.float64(arr[idx]) + np.float64(value)


# ------------------------------------------------------------------------------- 1D
@dataclass
class Line1D:
    """State + coefficients of one 1D problem.  ``form``:
    'fdtd'  ex = ca*ex + cb*(hy[i-1]-hy[i])               (programs 1_1 .. 1_5)
    'flux'  dx += 0.5*(..); ex = nax*(dx-ix-ncx*sx); ...   (programs 2_1 .. 2_3)
    """
    nx: int
    dtype: type = np.float64
    form: str = "fdtd"
    abc: bool = True
    src_index: int = 1
    src_hard: bool = False
    ca: np.ndarray = None
    cb: np.ndarray = None
    nax: np.ndarray = None
    nbx: np.ndarray = None
    ncx: np.ndarray = None
    ndx: np.ndarray = None
    freqs: np.ndarray = None          # running DFT frequencies (None = DFT off)
    dt: float = 0.01 / C_TIMES_2
    ex: np.ndarray = field(init=False)
    hy: np.ndarray = field(init=False)

    def __post_init__(self):
        z = lambda: np.zeros(self.nx, dtype=self.dtype)
        self.ex, self.hy, self.bc = z(), z(), np.zeros(4, dtype=self.dtype)
        if self.form == "fdtd":
            if self.ca is None:
                self.ca = np.full(self.nx, 1.0, dtype=self.dtype)
            if self.cb is None:
                self.cb = np.full(self.nx, 0.5, dtype=self.dtype)
        else:
            self.dx, self.ix, self.sx = z(), z(), z()
            for name, fill in (("nax", 1.0), ("nbx", 0.0), ("ncx", 0.0), ("ndx", 0.0)):
                if getattr(self, name) is None:
                    setattr(self, name, np.full(self.nx, fill, dtype=self.dtype))
        if self.freqs is not None:
            nf = len(self.freqs)
            self.r_pt = np.zeros((nf, self.nx), dtype=self.dtype)
            self.i_pt = np.zeros((nf, self.nx), dtype=self.dtype)
            self.r_in = np.zeros((nf, 1), dtype=self.dtype)
            self.i_in = np.zeros((nf, 1), dtype=self.dtype)


def e_update_1d(p: Line1D, src_value):
    """E (or D->E) half step + source.  fd1d_1_5.py:65-67; fd1d_2_3.py:73-86."""
    nx = p.nx
    curl = p.hy[0:nx - 1] - p.hy[1:nx]
    if p.form == "fdtd":
        p.ex[1:nx] = p.ca[1:nx] * p.ex[1:nx] + p.cb[1:nx] * curl
        _inject(p.ex, p.src_index, src_value, p.src_hard)
    else:
        p.dx[1:nx] += 0.5 * curl
        _inject(p.dx, p.src_index, src_value, p.src_hard)
        p.ex[1:nx] = p.nax[1:nx] * (p.dx[1:nx] - p.ix[1:nx] - p.ncx[1:nx] * p.sx[1:nx])
        p.ix[1:nx] += p.nbx[1:nx] * p.ex[1:nx]
        p.sx[1:nx] = p.ncx[1:nx] * p.sx[1:nx] + p.ndx[1:nx] * p.ex[1:nx]


def dft_1d(p: Line1D, t):
    """Running DFT of Ex at every cell + of the source sample ex[10].  fd1d_2_2.py:65-71."""
    nf = len(p.freqs)
    f = np.asarray(p.freqs, dtype=p.dtype).reshape(nf, 1)
    p.r_in[0:nf] += np.cos(2 * np.pi * f[0:nf] * p.dt * t) * p.ex[10]
    p.i_in[0:nf] -= np.sin(2 * np.pi * f[0:nf] * p.dt * t) * p.ex[10]
    p.r_pt[0:nf, 0:p.nx] += np.cos(2 * np.pi * f[0:nf] * p.dt * t) * p.ex[0:p.nx]
    p.i_pt[0:nf, 0:p.nx] -= np.sin(2 * np.pi * f[0:nf] * p.dt * t) * p.ex[0:p.nx]


def h_update_1d(p: Line1D):
    """Two-step-delay absorbing boundary, then the H half step.  fd1d_1_2.py:47-50."""
    nx, ex, bc = p.nx, p.ex, p.bc
    if p.abc:
        ex[0], bc[0], bc[1] = bc[0], bc[1], ex[1]
        ex[nx - 1], bc[3], bc[2] = bc[3], bc[2], ex[nx - 2]
    p.hy[0:nx - 1] += 0.5 * (ex[0:nx - 1] - ex[1:nx])


def advance_1d(p: Line1D, src, t_first=1):
    """Run len(src) steps; src[k] is the waveform sample of step t_first+k."""
    for k, t in enumerate(step_indices(len(src), t_first)):
        e_update_1d(p, src[k])
        if p.freqs is not None:
            dft_1d(p, t)
        h_update_1d(p)
    return p


def dft_amplitude_phase(r_pt, i_pt, r_in, i_in):
    """fd1d_2_2.py:145-146."""
    amp = 1 / np.hypot(r_in, i_in) * np.hypot(r_pt, i_pt)
    pha = np.arctan2(i_pt, r_pt) - np.arctan2(i_in, r_in)
    return amp, pha


def lossy_halfspace_fdtd(nx, dt, epsr, sigma, dtype, start=None, stop=None):
    """ca/cb of fd1d_1_5.py:37-44 (sigma=0 gives fd1d_1_3.py:31-34's cb)."""
    start = nx // 2 if start is None else start
    ca = 1.0 + np.zeros(nx, dtype=dtype)
    cb = 0.5 + np.zeros(nx, dtype=dtype)
    epsf = dt * sigma / (2 * EPS0 * epsr)
    ca[start:stop] = (1 - epsf) / (1 + epsf)
    cb[start:stop] = 0.5 / (epsr * (1 + epsf))
    return ca, cb


def lossy_halfspace_flux(nx, dt, epsr, sigma, dtype, chi=None, tau=None, start=None, stop=None):
    """nax/nbx[/ncx/ndx] of fd1d_2_1.py:64-72 and fd1d_2_3.py:97-109."""
    start = nx // 2 if start is None else start
    nax = np.full(nx, 1.0, dtype=dtype)
    nbx = np.full(nx, 0.0, dtype=dtype)
    ncx = np.full(nx, 0.0, dtype=dtype)
    ndx = np.full(nx, 0.0, dtype=dtype)
    if chi is None:
        nax[start:stop] = 1 / (epsr + sigma * dt / EPS0)
        nbx[start:stop] = sigma * dt / EPS0
    else:
        nax[start:stop] = 1 / (epsr + sigma * dt / EPS0 + chi * dt / tau)
        nbx[start:stop] = sigma * dt / EPS0
        ncx[start:stop] = np.exp(-dt / tau)
        ndx[start:stop] = chi * dt / tau
    return nax, nbx, ncx, ndx


# ------------------------------------------------------------------------------- 2D
PML_NAMES = ("fx1", "fx2", "fx3", "fy1", "fy2", "fy3", "gx2", "gx3", "gy2", "gy3")


def pml_vectors(nx, ny, npml, dtype):
    """fd2d/program/fd2d_3_3.py:113-122,147-158.  npml=0 gives the free-space identity set."""
    v = {k: np.full(nx if k[1] == "x" else ny, 0.0 if k[2] == "1" else 1.0, dtype=dtype)
         for k in PML_NAMES}
    for n in range(npml):
        xm = 0.33 * ((npml - n) / npml) ** 3
        xn = 0.33 * ((npml - n - 0.5) / npml) ** 3
        for ax, size in (("x", nx), ("y", ny)):
            f1, f2, f3 = v["f" + ax + "1"], v["f" + ax + "2"], v["f" + ax + "3"]
            g2, g3 = v["g" + ax + "2"], v["g" + ax + "3"]
            f1[n] = f1[size - 2 - n] = xn
            f2[n] = f2[size - 2 - n] = 1 / (1 + xn)
            f3[n] = f3[size - 2 - n] = (1 - xn) / (1 + xn)
            g2[n] = g2[size - 1 - n] = 1 / (1 + xm)
            g3[n] = g3[size - 1 - n] = (1 - xm) / (1 + xm)
    return v


def cylinder_medium(nx, ny, npml, rgrid, dt, epsr, sigma, dtype):
    """naz/nbz of a lossy dielectric cylinder, 3x3 sub-cell average.  fd2d/python/fd2d_3_4.py:173-194.
    Vectorised over (i, j); the nine sub-samples are accumulated in the reference's m, n order."""
    naz = np.full((nx, ny), 1.0, dtype=dtype)
    nbz = np.full((nx, ny), 0.0, dtype=dtype)
    ii = np.arange(npml, nx - npml, dtype=np.float64)[:, None]
    jj = np.arange(npml, ny - npml, dtype=np.float64)[None, :]
    epsn = np.full((ii.shape[0], jj.shape[1]), 1.0)
    cond = np.zeros_like(epsn)
    for m in range(-1, 2):
        for n in range(-1, 2):
            x = nx / 2 - 1 - ii + m / 3
            y = ny / 2 - 1 - jj + n / 3
            inside = np.sqrt(x ** 2 + y ** 2) <= rgrid
            epsn = np.where(inside, epsn + (epsr - 1) / 9, epsn)
            cond = np.where(inside, cond + sigma / 9, cond)
    naz[npml:nx - npml, npml:ny - npml] = 1 / (epsn + cond * dt / EPS0)
    nbz[npml:nx - npml, npml:ny - npml] = cond * dt / EPS0
    return naz, nbz


@dataclass
class Grid2D:
    """State + coefficients of one 2D TM problem (superset program 3_4)."""
    nx: int
    ny: int
    npml: int = 0
    dtype: type = np.float64
    tfsf: bool = False                 # incident line + TFSF corrections (3_3, 3_4)
    lossy: bool = False                # iz / nbz (3_4)
    point: tuple = None                # (i, j) of the point source on dz (3_1, 3_2)
    point_hard: bool = True
    freqs: np.ndarray = None           # running DFT (3_4)
    dt: float = 0.01 / C_TIMES_2
    naz: np.ndarray = None
    nbz: np.ndarray = None
    pml: dict = None

    def __post_init__(self):
        nx, ny, dt = self.nx, self.ny, self.dtype
        z2 = lambda: np.zeros((nx, ny), dtype=dt)
        self.dz, self.ez, self.hx, self.hy, self.ihx, self.ihy = (z2() for _ in range(6))
        if self.naz is None:
            self.naz = np.ones((nx, ny), dtype=dt)
        if self.lossy:
            self.iz = z2()
            if self.nbz is None:
                self.nbz = z2()
        if self.pml is None:
            self.pml = pml_vectors(nx, ny, self.npml, dt)
        if self.tfsf:
            self.ezi = np.zeros(ny, dtype=dt)
            self.hxi = np.zeros(ny, dtype=dt)
            self.bc = np.zeros(4, dtype=dt)
        if self.freqs is not None:
            nf = len(self.freqs)
            self.r_pt = np.zeros((nf, nx, ny), dtype=dt)
            self.i_pt = np.zeros((nf, nx, ny), dtype=dt)
            self.r_in = np.zeros(nf, dtype=dt)
            self.i_in = np.zeros(nf, dtype=dt)

    FIELDS = ("dz", "ez", "hx", "hy", "ihx", "ihy")


def incident_e(g: Grid2D):
    """ezinct: 1D incident Ez along j + both-end ABC.  fd2d_3_3.py:60-65."""
    ny, ezi, hxi, bc = g.ny, g.ezi, g.hxi, g.bc
    ezi[1:ny] += 0.5 * (hxi[0:ny - 1] - hxi[1:ny])
    ezi[0], bc[0], bc[1] = bc[0], bc[1], ezi[1]
    ezi[ny - 1], bc[3], bc[2] = bc[3], bc[2], ezi[ny - 2]


def incident_h(g: Grid2D):
    """hxinct.  fd2d_3_3.py:86-88."""
    ny = g.ny
    g.hxi[0:ny - 1] += 0.5 * (g.ezi[0:ny - 1] - g.ezi[1:ny])


def d_update_2d(g: Grid2D, src_value):
    """dfield (+ its embedded source assignment).  fd2d_3_2.py:61-66, fd2d_3_3.py:68-72.
    With the identity PML set this is bit-identical to the free-space form fd2d_3_1.py:44-48."""
    nx, ny, p = g.nx, g.ny, g.pml
    dz, hx, hy = g.dz, g.hx, g.hy
    dz[1:nx, 1:ny] = (p["gx3"][1:nx, None] * p["gy3"][1:ny] * dz[1:nx, 1:ny]
                      + p["gx2"][1:nx, None] * p["gy2"][1:ny] * 0.5
                      * (hy[1:nx, 1:ny] - hy[0:nx - 1, 1:ny] - hx[1:nx, 1:ny] + hx[1:nx, 0:ny - 1]))
    if g.tfsf:
        g.ezi[3] = src_value                       # hard source of the incident line
    elif g.point is not None:
        _inject(dz, g.point, src_value, g.point_hard)


def tfsf_d(g: Grid2D):
    """inctdz.  fd2d_3_3.py:75-78."""
    nx, ny, n = g.nx, g.ny, g.npml
    g.dz[n - 1:nx - n + 1, n - 1] += 0.5 * g.hxi[n - 2]
    g.dz[n - 1:nx - n + 1, ny - n] -= 0.5 * g.hxi[ny - n]


def e_update_2d(g: Grid2D):
    """efield over the full array.  fd2d_3_3.py:81-83; lossy form fd2d/python/fd2d_3_4.py:131-136."""
    if g.lossy:
        g.ez[:, :] = g.naz * (g.dz - g.iz)
        g.iz += g.nbz * g.ez
    else:
        g.ez[:, :] = g.naz * g.dz


def dft_2d(g: Grid2D, t):
    """fourier.  fd2d/python/fd2d_3_4.py:89-99 (numba: phase factors in float64, product promoted
    to float64, accumulated into the array dtype)."""
    for n, f in enumerate(np.asarray(g.freqs, dtype=g.dtype)):
        c = np.cos(2 * np.pi * np.float64(f) * g.dt * t)
        s = np.sin(2 * np.pi * np.float64(f) * g.dt * t)
        g.r_in[n] = np.float64(g.r_in[n]) + c * np.float64(g.ezi[6])
        g.i_in[n] = np.float64(g.i_in[n]) - s * np.float64(g.ezi[6])
        g.r_pt[n] = (g.r_pt[n].astype(np.float64) + c * g.ez.astype(np.float64)).astype(g.dtype)
        g.i_pt[n] = (g.i_pt[n].astype(np.float64) - s * g.ez.astype(np.float64)).astype(g.dtype)


def h_update_2d(g: Grid2D):
    """hfield with the PML integrals.  fd2d_3_3.py:91-98 (identity PML set == fd2d_3_1.py:56-59)."""
    nx, ny, p = g.nx, g.ny, g.pml
    ez, hx, hy, ihx, ihy = g.ez, g.hx, g.hy, g.ihx, g.ihy
    curl_m = ez[0:nx - 1, 0:ny - 1] - ez[0:nx - 1, 1:ny]
    curl_n = ez[0:nx - 1, 0:ny - 1] - ez[1:nx, 0:ny - 1]
    ihx[0:nx - 1, 0:ny - 1] += curl_m
    ihy[0:nx - 1, 0:ny - 1] += curl_n
    hx[0:nx - 1, 0:ny - 1] = (p["fy3"][0:ny - 1] * hx[0:nx - 1, 0:ny - 1]
                              + p["fy2"][0:ny - 1] * (0.5 * curl_m + p["fx1"][0:nx - 1, None] * ihx[0:nx - 1, 0:ny - 1]))
    hy[0:nx - 1, 0:ny - 1] = (p["fx3"][0:nx - 1, None] * hy[0:nx - 1, 0:ny - 1]
                              - p["fx2"][0:nx - 1, None] * (0.5 * curl_n + p["fy1"][0:ny - 1] * ihy[0:nx - 1, 0:ny - 1]))


def tfsf_h(g: Grid2D):
    """incthx + incthy.  fd2d_3_3.py:101-110."""
    nx, ny, n, ezi = g.nx, g.ny, g.npml, g.ezi
    g.hx[n - 1:nx - n + 1, n - 2] += 0.5 * ezi[n - 1]
    g.hx[n - 1:nx - n + 1, ny - n] -= 0.5 * ezi[ny - n]
    g.hy[n - 2, n - 1:ny - n + 1] -= 0.5 * ezi[n - 1:ny - n + 1]
    g.hy[nx - n, n - 1:ny - n + 1] += 0.5 * ezi[n - 1:ny - n + 1]


def step_2d(g: Grid2D, t, src_value):
    """One full time step in the reference order (fd2d/python/fd2d_3_4.py:268-277)."""
    if g.tfsf:
        incident_e(g)
    d_update_2d(g, src_value)
    if g.tfsf:
        tfsf_d(g)
    e_update_2d(g)
    if g.freqs is not None:
        dft_2d(g, t)
    if g.tfsf:
        incident_h(g)
    h_update_2d(g)
    if g.tfsf:
        tfsf_h(g)


def advance_2d(g: Grid2D, src, t_first=1):
    for k, t in enumerate(step_indices(len(src), t_first)):
        step_2d(g, t, src[k])
    return g
