"""Second pin of the numpy oracle: the REFERENCE's own C/OpenMP step functions (compiled from the sources
under /root/reference into oracle/_ref/, see oracle/Makefile) run next to the oracle in fp32 on grids too
large for quick numpy-vs-numpy checks.  Once the float32-evaluated source of the C program is replaced by
the float64-evaluated sample the Python programs use, the two agree exactly on every value above 1e-12 (peak ~1);
differences are born inside the SUBNORMAL range (|x| < 2**-126), where the C form ``0.5f*a - 0.5f*b`` and the
numpy form ``0.5*(a-b)`` round differently.  The numpy programs are the oracle."""
import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from oracle import ref_c
from tests import cases

needs_ref = pytest.mark.skipif(not ref_c.available(), reason="oracle/_ref not built (no /root/reference at build time)")


@needs_ref
@pytest.mark.parametrize("prog,nx,ny,npml,ns", [("3_2", 200, 328, 24, 120), ("3_3", 256, 192, 20, 150)])
def test_reference_c_functions_equal_oracle_fp32(prog, nx, ny, npml, ns):
    lib = ref_c.load(prog)
    a, src = cases.grid_program(prog, nx, ny, ns, np.float32, npml=npml)
    b, _ = cases.grid_program(prog, nx, ny, ns, np.float32, npml=npml)
    step = ref_c.step_3_2 if prog == "3_2" else ref_c.step_3_3
    for k, t in enumerate(orc.step_indices(ns)):
        step(lib, t, a, src[k])
    orc.advance_2d(b, src)
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.abs(x - y).max() <= 1e-20, name         # subnormal-born differences (leading edge, deep PML) stay tiny
        normal = np.abs(y) >= 1e-12
        assert np.array_equal(x[normal], y[normal]), name
    assert np.abs(b.ez).max() > 1e-3
