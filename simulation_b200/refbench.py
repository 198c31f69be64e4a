"""The reference's own benchmark definitions -- its ``test_*`` twins -- on the GPU path.

Every ``test_*`` program of the reference is a closed script: hard-coded sizes, a timed ``for t in ...`` loop, then
``print`` of the compute time and of a slice of the final field (no asserts; a human compares the lines across the six
language variants).  This module keeps each program's setup literals and replaces only the loop (``advance``), then
returns / prints the same lines, so a user of the reference can run "the same benchmark" here:

    1D  fd1d/program/test_1_1.py .. test_2_3.py   nx=38000, ns=40000, fp32;  prints ex[0:50]
    2D  fd2d/python/test_3_1.py  .. test_3_4.py   1024 x 1024, ns=5000, npml=80 (3_1: no PML), fp32;  prints ez[2][0:50];
        3_4 also prints amplt[2][0:ny-50] (fd2d/python/test_3_4.py:289-293)

Sizes can be overridden (the tests run reduced ones).  Post-processing of the running DFT (amplitude) is the
reference's own float32 numpy expression on the downloaded accumulators (fd1d/program/test_2_2.py:149,
fd2d/python/test_3_4.py:283-287).
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from . import fd1d, fd2d, surface

DT = surface.DT
PROGRAMS_1D = ("1_1", "1_2", "1_3", "1_4", "1_5", "2_1", "2_2", "2_3")
PROGRAMS_2D = ("3_1", "3_2", "3_3", "3_4")


def _sync(device):
    if torch.cuda.is_available():
        torch.cuda.synchronize(device)


def line_1d(prog: str, nx: int = 38000, dtype=np.float32, device=None, tblock: int = 64) -> fd1d.Fdtd1D:
    """The problem of fd1d/program/test_<prog>.py (setup literals as shipped)."""
    S = fd1d.LineSource
    kw = dict(device=device, tblock=tblock)
    if prog == "1_1":       # free space, hard Gaussian at nx//2, no ABC (test_1_1.py:41-48)
        return fd1d.Fdtd1D(nx, dtype, abc=False, source=S(nx // 2, surface.Gaussian(40, 12.0), hard=True), **kw)
    if prog == "1_2":       # + the two-step-delay ABC (test_1_2.py:44-52)
        return fd1d.Fdtd1D(nx, dtype, abc=True, source=S(nx // 2, surface.Gaussian(40, 12.0), hard=True), **kw)
    if prog in ("1_3", "1_4"):    # dielectric half space eps_r = 4, soft Gaussian / 700 MHz sinusoid at ex[1]
        _, cb = surface.dielectric_fdtd(nx, DT, 4.0, 0.0, dtype)
        wave = surface.Gaussian(40, 12.0) if prog == "1_3" else surface.Sinusoid(700e6)
        return fd1d.Fdtd1D(nx, dtype, source=S(1, wave), cb=cb, **kw)
    if prog == "1_5":       # lossy half space eps_r = 4, sigma = 0.04 (test_1_5.py:37-45)
        ca, cb = surface.dielectric_fdtd(nx, DT, 4.0, 0.04, dtype)
        return fd1d.Fdtd1D(nx, dtype, source=S(1, surface.Sinusoid(700e6)), ca=ca, cb=cb, **kw)
    if prog == "2_1":       # flux form, same medium (test_2_1.py:65-73)
        nax, nbx, _, _ = surface.dielectric_flux(nx, DT, 4.0, 0.04, dtype)
        return fd1d.Fdtd1D(nx, dtype, form="flux", source=S(1, surface.Sinusoid(700e6), field="dx"), nax=nax, nbx=nbx, **kw)
    if prog == "2_2":       # eps_r = 4, sigma = 0, Gaussian, running DFT at 100 / 200 / 500 MHz (test_2_2.py:121-126)
        nax, nbx, _, _ = surface.dielectric_flux(nx, DT, 4.0, 0.0, dtype)
        return fd1d.Fdtd1D(nx, dtype, form="flux", source=S(1, surface.Gaussian(50, 10.0), field="dx"), nax=nax, nbx=nbx,
                           freqs=[100e6, 200e6, 500e6], **kw)
    if prog == "2_3":       # Debye medium, running DFT at 50 / 200 / 500 MHz (test_2_3.py:128-135)
        nax, nbx, ncx, ndx = surface.dielectric_flux(nx, DT, 2.0, 0.01, dtype, chi=2.0, tau=0.001e-6)
        return fd1d.Fdtd1D(nx, dtype, form="flux", source=S(1, surface.Gaussian(50, 10.0), field="dx"), nax=nax, nbx=nbx,
                           ncx=ncx, ndx=ndx, freqs=[50e6, 200e6, 500e6], **kw)
    raise KeyError(prog)


def grid_2d(prog: str, nx: int = 1024, ny: int = 1024, npml: int = 80, radius: float = 1.50, dtype=np.float32,
            device=None, dft: bool = True) -> fd2d.Fdtd2D:
    """The problem of fd2d/python/test_<prog>.py (setup literals as shipped)."""
    if prog == "3_1":       # free space, hard Gaussian at the centre (test_3_1.py:54)
        return fd2d.Fdtd2D(nx, ny, 0, dtype, source=fd2d.PointSource(nx // 2, ny // 2, surface.Gaussian(20, 6.0)), device=device)
    if prog == "3_2":       # PML, hard 1500 MHz sinusoid (test_3_2.py:70)
        return fd2d.Fdtd2D(nx, ny, npml, dtype, source=fd2d.PointSource(nx // 2 - 5, ny // 2 - 5, surface.Sinusoid(1500e6)),
                           device=device)
    if prog == "3_3":       # PML + TFSF plane wave (test_3_3.py:78)
        return fd2d.Fdtd2D(nx, ny, npml, dtype, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), device=device)
    if prog == "3_4":       # + lossy dielectric cylinder eps_r = 30, sigma = 0.3, running DFT at 50 / 300 / 700 MHz
        rgrid = int(radius / surface.DS - 1)
        if device is None or torch.device(device).type == "cuda":
            md = fd2d.dielectric(nx, ny, npml, rgrid, DT, 30.0, 0.30, dtype, device=device)
            naz, nbz = md.naz, md.nbz
        else:               # (CPU tensors only under the test suite's emulated device)
            naz, nbz = surface.dielectric_cylinder(nx, ny, npml, rgrid, DT, 30.0, 0.30, dtype)
        return fd2d.Fdtd2D(nx, ny, npml, dtype, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=naz, nbz=nbz,
                           freqs=[50e6, 300e6, 700e6] if dft else None, device=device)
    raise KeyError(prog)


def amplitude_1d(sim: fd1d.Fdtd1D) -> np.ndarray:
    """``amplt = 1/hypot(r_in, i_in) * hypot(r_pt, i_pt)`` (fd1d/program/test_2_2.py:149), float32 numpy as there."""
    r_in, i_in, r_pt, i_pt = (sim.get(n) for n in ("r_in", "i_in", "r_pt", "i_pt"))
    with np.errstate(divide="ignore", invalid="ignore"):
        return 1 / np.hypot(r_in, i_in) * np.hypot(r_pt, i_pt)


def amplitude_2d(sim: fd2d.Fdtd2D) -> np.ndarray:
    """``amplt[n, j]`` along row nx//2-1 for j in [npml-1, ny-npml] (fd2d/python/test_3_4.py:283-287); zero elsewhere."""
    nf, nx, ny, npml = len(sim.freqs), sim.nx, sim.ny, sim.npml
    r_in, i_in, r_pt, i_pt = (sim.get(n) for n in ("r_in", "i_in", "r_pt", "i_pt"))
    amplt = np.zeros((nf, ny), dtype=sim.np_dtype)
    i, ja, jb = nx // 2 - 1, npml - 1, ny - npml + 1
    with np.errstate(divide="ignore", invalid="ignore"):
        for n in range(nf):
            amplt[n, ja:jb] = 1 / np.hypot(r_in[n], i_in[n]) * np.hypot(r_pt[n, i, ja:jb], i_pt[n, i, ja:jb])
    return amplt


def run(prog: str, ns: Optional[int] = None, warm: bool = True, **size) -> dict:
    """Run one benchmark definition.  -> {"prog", "seconds", "cells", "ns", "lines": the arrays the reference prints, in
    order, "ex" | "ez": the final field, "amplt": amplitude where the program computes one}.
    ``warm``: one untimed short run of the same problem first (module load, clocks), as every timing here does."""
    one_d = prog in PROGRAMS_1D
    make = (lambda: line_1d(prog, **size)) if one_d else (lambda: grid_2d(prog, **size))
    ns = int(ns if ns is not None else (40000 if one_d else 5000))
    if warm:
        w = make()
        w.advance(min(ns, 128 if one_d else 24))
        _sync(w.device)
        del w
    sim = make()
    _sync(sim.device)
    t0 = time.perf_counter()
    sim.advance(ns)
    _sync(sim.device)
    amplt = None
    if sim.ft is not None:      # the reference keeps the amplitude post-processing inside its timed region too
        amplt = amplitude_1d(sim) if one_d else amplitude_2d(sim)
    seconds = time.perf_counter() - t0
    out = {"prog": prog, "seconds": seconds, "ns": ns, "amplt": amplt}
    if one_d:
        ex = sim.get("ex")
        out.update(cells=sim.nx, ex=ex, lines=[ex[0:50]])
    else:
        ez = sim.get("ez")
        out.update(cells=sim.nx * sim.ny, ez=ez, lines=[ez[2][0:50]] + ([amplt[2][0:sim.ny - 50]] if amplt is not None else []))
    return out


def report(result: dict) -> list:
    """The lines the reference program prints (``Total compute time on GPU: ... s`` as fd2d/cuda/test_3_4.cu:361 words
    it, then the printed slices in numpy's default formatting)."""
    out = [f"Total compute time on GPU: {result['seconds']:.3f} s"]
    out += [str(a) for a in result["lines"]]
    return out
