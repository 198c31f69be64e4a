"""ctypes driver of the REFERENCE's own C/OpenMP step functions (oracle/_ref/libref_fd2d_*.so, built by
oracle/Makefile from /root/reference/fd2d/clang/test_3_*.c with -Dmain=ref_program_main).

TEST INFRASTRUCTURE ONLY: used by tests to cross-check the numpy oracle against a second reference
implementation, and by ``bench.py --impl reference`` as the reference arm (all host threads).  The
product never imports this.

The C programs hard-code their source waveform inside dfield in float32 (expf/sinf); callers that want
bit-parity with the numpy programs overwrite the (hard) source cell after the call with the float64-evaluated
sample -- see :func:`step_3_2` / :func:`step_3_3`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FP = C.POINTER(C.c_float)


class CPml(C.Structure):           # fd2d/clang/test_3_3.c:14-19
    _fields_ = [(n, FP) for n in ("fx1", "fx2", "fx3", "fy1", "fy2", "fy3", "gx2", "gx3", "gy2", "gy3")]


def available(prog="3_2") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libref_fd2d_{prog}.so"))


def load(prog):
    return C.CDLL(os.path.join(HERE, "_ref", f"libref_fd2d_{prog}.so"))


def _p(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(FP)


def pml_struct(pml: dict) -> CPml:
    return CPml(*[_p(pml[k]) for k in ("fx1", "fx2", "fx3", "fy1", "fy2", "fy3", "gx2", "gx3", "gy2", "gy3")])


def step_3_2(lib, t, g, src_value):
    """One step of the reference C program 3_2 on an oracle Grid2D (fp32); point source forced to src_value."""
    nx, ny = g.nx, g.ny
    ps = pml_struct(g.pml)
    lib.dfield(C.c_int(int(t)), nx, ny, C.byref(ps), _p(g.dz), _p(g.hx), _p(g.hy))
    g.dz[nx // 2 - 5, ny // 2 - 5] = src_value
    lib.efield(nx, ny, _p(g.naz), _p(g.dz), _p(g.ez))
    lib.hfield(nx, ny, C.byref(ps), _p(g.ez), _p(g.ihx), _p(g.ihy), _p(g.hx), _p(g.hy))


def step_3_3(lib, t, g, src_value):
    """One step of the reference C program 3_3 (TFSF) on an oracle Grid2D (fp32)."""
    nx, ny, n = g.nx, g.ny, g.npml
    ps = pml_struct(g.pml)
    lib.ezinct(ny, _p(g.ezi), _p(g.hxi), _p(g.bc))
    lib.dfield(C.c_int(int(t)), nx, ny, C.byref(ps), _p(g.ezi), _p(g.dz), _p(g.hx), _p(g.hy))
    g.ezi[3] = src_value
    lib.inctdz(nx, ny, n, _p(g.hxi), _p(g.dz))
    lib.efield(nx, ny, _p(g.naz), _p(g.dz), _p(g.ez))
    lib.hxinct(ny, _p(g.ezi), _p(g.hxi))
    lib.hfield(nx, ny, C.byref(ps), _p(g.ez), _p(g.ihx), _p(g.ihy), _p(g.hx), _p(g.hy))
    lib.incthx(nx, ny, n, _p(g.ezi), _p(g.hx))
    lib.incthy(nx, ny, n, _p(g.ezi), _p(g.hy))


class CMedium(C.Structure):        # fd2d/clang/test_3_4.c: typedef struct { float *naz, *nbz; } medium;
    _fields_ = [("naz", FP), ("nbz", FP)]


def step_3_4(lib, t, g, src_value):
    """One step of the reference C program 3_4 (TFSF + lossy medium, DFT skipped) on an oracle Grid2D (fp32)."""
    nx, ny, n = g.nx, g.ny, g.npml
    ps = pml_struct(g.pml)
    md = CMedium(_p(g.naz), _p(g.nbz))
    lib.ezinct(ny, _p(g.ezi), _p(g.hxi), _p(g.bc))
    lib.dfield(C.c_int(int(t)), nx, ny, C.byref(ps), _p(g.ezi), _p(g.dz), _p(g.hx), _p(g.hy))
    g.ezi[3] = src_value
    lib.inctdz(nx, ny, n, _p(g.hxi), _p(g.dz))
    lib.efield(nx, ny, C.byref(md), _p(g.dz), _p(g.iz), _p(g.ez))
    lib.hxinct(ny, _p(g.ezi), _p(g.hxi))
    lib.hfield(nx, ny, C.byref(ps), _p(g.ez), _p(g.ihx), _p(g.ihy), _p(g.hx), _p(g.hy))
    lib.incthx(nx, ny, n, _p(g.ezi), _p(g.hx))
    lib.incthy(nx, ny, n, _p(g.ezi), _p(g.hy))
