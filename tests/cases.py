"""Oracle-side descriptions of the reference programs (sizes, media, sources) shared by the tests.

Every builder returns ``(problem, src)``: an ``oracle.fdtd_oracle`` Line1D / Grid2D in its initial
state and the float64 source table for ``ns`` steps.  Literals are the reference's own:
fd1d/program/fd1d_1_1.py:33-43 ... fd1d_2_3.py:112-131, fd2d/program/fd2d_3_1.py:62-80,
fd2d_3_2.py:98-132, fd2d_3_3.py:125-164, fd2d/python/fd2d_3_4.py:211-266.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import fdtd_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = 0.01 / 6e8


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def line_program(prog, nx, ns, dtype):
    """1D programs 1_1 .. 2_3 at arbitrary nx / ns / dtype."""
    if prog in ("1_1", "1_2"):
        p = orc.Line1D(nx, dtype, form="fdtd", abc=(prog == "1_2"), src_index=nx // 2, src_hard=True)
        src = orc.source_table("gaussian", ns, t0=40, spread=12.0)
    elif prog in ("1_3", "1_4", "1_5"):
        sigma = 0.04 if prog == "1_5" else 0.0
        ca, cb = orc.lossy_halfspace_fdtd(nx, DT, 4.0, sigma, dtype)
        p = orc.Line1D(nx, dtype, form="fdtd", ca=ca, cb=cb)
        src = (orc.source_table("gaussian", ns, t0=40, spread=12.0) if prog == "1_3"
               else orc.source_table("sine", ns, freq=700e6))
    elif prog == "2_1":
        nax, nbx, _, _ = orc.lossy_halfspace_flux(nx, DT, 4.0, 0.04, dtype)
        p = orc.Line1D(nx, dtype, form="flux", nax=nax, nbx=nbx)
        src = orc.source_table("sine", ns, freq=700e6)
    elif prog == "2_2":
        nax, nbx, _, _ = orc.lossy_halfspace_flux(nx, DT, 4.0, 0.0, dtype)
        p = orc.Line1D(nx, dtype, form="flux", nax=nax, nbx=nbx,
                       freqs=np.array((100e6, 200e6, 500e6), dtype=dtype))
        src = orc.source_table("gaussian", ns, t0=50, spread=10.0)
    elif prog == "2_3":
        nax, nbx, ncx, ndx = orc.lossy_halfspace_flux(nx, DT, 2.0, 0.01, dtype, chi=2.0, tau=0.001e-6)
        p = orc.Line1D(nx, dtype, form="flux", nax=nax, nbx=nbx, ncx=ncx, ndx=ndx,
                       freqs=np.array((50e6, 200e6, 500e6), dtype=dtype))
        src = orc.source_table("gaussian", ns, t0=50, spread=10.0)
    else:
        raise KeyError(prog)
    return p, src


LINE_MAIN = {"1_1": (512, 300), "1_2": (512, 570), "1_3": (512, 740), "1_4": (512, 740),
             "1_5": (512, 740), "2_1": (512, 740), "2_2": (512, 740), "2_3": (512, 740)}


def grid_program(prog, nx, ny, ns, dtype, npml=8, naz=None, radius=0.15, dft=True, freqs=(50e6, 300e6, 700e6)):
    """2D programs 3_1 .. 3_4 at arbitrary size / dtype."""
    if prog == "3_1":
        g = orc.Grid2D(nx, ny, 0, dtype, point=(nx // 2, ny // 2), naz=naz)
        src = orc.source_table("gaussian", ns, t0=20, spread=6.0)
    elif prog == "3_2":
        g = orc.Grid2D(nx, ny, npml, dtype, point=(nx // 2 - 5, ny // 2 - 5), naz=naz)
        src = orc.source_table("sine", ns, freq=1500e6)
    elif prog == "3_3":
        g = orc.Grid2D(nx, ny, npml, dtype, tfsf=True, naz=naz)
        src = orc.source_table("gaussian", ns, t0=20, spread=8.0)
    elif prog == "3_4":
        rgrid = int(radius / 0.01 - 1)
        md_naz, md_nbz = orc.cylinder_medium(nx, ny, npml, rgrid, DT, 30.0, 0.30, dtype)
        g = orc.Grid2D(nx, ny, npml, dtype, tfsf=True, lossy=True, naz=md_naz, nbz=md_nbz,
                       freqs=np.array(freqs, dtype=dtype) if dft else None)
        src = orc.source_table("gaussian", ns, t0=20, spread=8.0)
    else:
        raise KeyError(prog)
    return g, src


GRID_MAIN = {"3_1": (60, 60, 70), "3_2": (60, 60, 100), "3_3": (60, 60, 115)}
GRID_MAIN_NUMBA = {"3_1": (100, 100, 90), "3_2": (100, 100, 120), "3_3": (100, 100, 120),
                   "3_4": (100, 100, 120)}


def amplitude_row(g, n=2):
    """Post-processing of program 3_4: amplitude along row nx//2-1 (fd2d/python/fd2d_3_4.py:279-284)."""
    amp = np.zeros(g.ny, dtype=g.dtype)
    i = g.nx // 2 - 1
    for j in range(g.npml - 1, g.ny - g.npml + 1):
        amp[j] = 1 / np.hypot(g.r_in[n], g.i_in[n]) * np.hypot(g.r_pt[n, i, j], g.i_pt[n, i, j])
    return amp
