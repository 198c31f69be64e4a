"""Pin the numpy oracle: (a) against the committed goldens (made by executing the reference, see
tests/golden/make_golden.py) -- runs everywhere; (b) against the reference executed live -- build
container only.  Bitwise for the numpy programs; 1e-12 relative for the numba/fastmath ones."""
import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from tests import cases


def same_bits(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


# ------------------------------------------------------------------ 1D
@pytest.mark.parametrize("prog", sorted(cases.LINE_MAIN))
def test_line_main_golden_fp64(prog):
    nx, ns = cases.LINE_MAIN[prog]
    p, src = cases.line_program(prog, nx, ns, np.float64)
    orc.advance_1d(p, src)
    g = cases.golden(f"main_fd1d_{prog}")
    assert same_bits(p.ex, g["ex"])
    if p.freqs is not None:
        amp, _ = orc.dft_amplitude_phase(p.r_pt, p.i_pt, p.r_in, p.i_in)
        assert same_bits(amp[2], g["amplt2"])


@pytest.mark.parametrize("prog", sorted(cases.LINE_MAIN))
def test_line_twin_golden_fp32(prog):
    g = cases.golden(f"twin_fd1d_{prog}")
    nx, ns = int(g["nx"]), int(g["ns"])
    p, src = cases.line_program(prog, nx, ns, np.float32)
    orc.advance_1d(p, src)
    assert same_bits(p.ex, g["ex"])
    if p.freqs is not None:
        amp, _ = orc.dft_amplitude_phase(p.r_pt, p.i_pt, p.r_in, p.i_in)
        assert same_bits(amp[2], g["amplt2"])


def test_line_split_advance_is_identical():
    """advance(a) then advance(b) == advance(a+b): the source table is indexed by absolute step."""
    p, src = cases.line_program("1_5", 300, 400, np.float32)
    q, _ = cases.line_program("1_5", 300, 400, np.float32)
    orc.advance_1d(p, src)
    orc.advance_1d(q, src[:170], t_first=1)
    orc.advance_1d(q, src[170:], t_first=171)
    assert same_bits(p.ex, q.ex) and same_bits(p.hy, q.hy)


# ------------------------------------------------------------------ 2D
@pytest.mark.parametrize("prog", sorted(cases.GRID_MAIN))
def test_grid_main_golden_fp64(prog):
    nx, ny, ns = cases.GRID_MAIN[prog]
    g, src = cases.grid_program(prog, nx, ny, ns, np.float64)
    orc.advance_2d(g, src)
    assert np.array_equal(g.ez, cases.golden(f"main_fd2d_{prog}")["ez"])     # +-0 tolerant for 3_1
    if prog != "3_1":
        assert same_bits(g.ez, cases.golden(f"main_fd2d_{prog}")["ez"])


@pytest.mark.parametrize("prog", sorted(cases.GRID_MAIN_NUMBA))
def test_grid_main_numba_golden(prog):
    nx, ny, ns = cases.GRID_MAIN_NUMBA[prog]
    g, src = cases.grid_program(prog, nx, ny, ns, np.float64)
    orc.advance_2d(g, src)
    ref = cases.golden(f"main_numba_fd2d_{prog}")
    peak = np.abs(ref["ez"]).max()
    assert np.abs(g.ez - ref["ez"]).max() <= 1e-12 * peak
    if prog == "3_4":
        amp = cases.amplitude_row(g)
        assert np.abs(amp - ref["amplt2"]).max() <= 1e-10 * np.abs(ref["amplt2"]).max()


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("prog", ["3_1", "3_2", "3_3"])
def test_grid_drive_golden_all_fields(prog, tag, dtype):
    ref = cases.golden(f"drive_{prog}_{tag}")
    nx, ny, ns = int(ref["nx"]), int(ref["ny"]), int(ref["ns"])
    npml = int(ref["npml"]) if "npml" in ref else 0
    g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml)
    orc.advance_2d(g, src)
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy", "ezi", "hxi", "bc"):
        if name in ref:
            if prog == "3_1":
                assert np.array_equal(getattr(g, name), ref[name]), name
            else:
                assert same_bits(getattr(g, name), ref[name]), name


def test_grid_drive_random_medium_fp32():
    ref = cases.golden("drive_3_2_randnaz_f32")
    nx, ny, ns, npml = (int(ref[k]) for k in ("nx", "ny", "ns", "npml"))
    g, src = cases.grid_program("3_2", nx, ny, ns, np.float32, npml=npml, naz=ref["naz"].copy())
    orc.advance_2d(g, src)
    for name in orc.Grid2D.FIELDS:
        assert same_bits(getattr(g, name), ref[name]), name


def test_grid_3_4_drive_golden_fp64():
    ref = cases.golden("drive_3_4_f64")
    nx, ny, ns, npml = (int(ref[k]) for k in ("nx", "ny", "ns", "npml"))
    g, src = cases.grid_program("3_4", nx, ny, ns, np.float64, npml=npml, radius=0.12)
    assert int(ref["rgrid"]) == 11
    # numba's fastmath dielectric() differs from the Python semantics of its own source by <= 1 ulp
    assert np.abs(g.naz - ref["naz"]).max() <= 2.3e-16 and np.abs(g.nbz - ref["nbz"]).max() <= 2.3e-16
    orc.advance_2d(g, src)
    peak = np.abs(ref["ez"]).max()
    for name in ("dz", "ez", "iz", "hx", "hy", "ihx", "ihy", "ezi", "hxi"):
        scale = max(np.abs(ref[name]).max(), peak)
        assert np.abs(getattr(g, name) - ref[name]).max() <= 1e-12 * scale, name
    for name in ("r_pt", "i_pt", "r_in", "i_in"):
        assert np.abs(getattr(g, name) - ref[name]).max() <= 1e-11 * np.abs(ref[name]).max(), name


def test_pml_vectors_nonsquare_and_identity():
    v = orc.pml_vectors(40, 56, 0, np.float32)
    assert all((v[k] == (0.0 if k[2] == "1" else 1.0)).all() for k in orc.PML_NAMES)
    v = orc.pml_vectors(40, 56, 7, np.float64)
    assert v["fx1"].shape == (40,) and v["gy3"].shape == (56,)
    assert v["fx1"][0] == v["fx1"][40 - 2] == v["fy1"][56 - 2] and v["gx2"][0] == v["gy2"][56 - 1]
    assert v["fx1"][40 - 1] == 0.0 and v["gx3"][7] == 1.0


# ------------------------------------------------------------------ live reference (build container)
@pytest.mark.reference
@pytest.mark.parametrize("prog", ["1_2", "1_5", "2_3"])
def test_live_reference_line_main(prog):
    from oracle import refload
    seen = refload.run_main(f"fd1d/program/fd1d_{prog}.py")
    ex = [a for a in seen["visualize"] if isinstance(a, np.ndarray)][-1]
    nx, ns = cases.LINE_MAIN[prog]
    p, src = cases.line_program(prog, nx, ns, np.float64)
    orc.advance_1d(p, src)
    assert same_bits(p.ex, ex)


@pytest.mark.reference
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_live_reference_grid_3_3_functions(dtype):
    """Drive the reference's own step functions next to the oracle's on a fresh odd-sized grid."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    nx, ny, npml, ns = 53, 67, 9, 160
    ref = make_golden.drive_3_3(nx, ny, npml, ns, dtype)
    g, src = cases.grid_program("3_3", nx, ny, ns, dtype, npml=npml)
    orc.advance_2d(g, src)
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy", "ezi", "hxi", "bc"):
        assert same_bits(getattr(g, name), ref[name]), name
