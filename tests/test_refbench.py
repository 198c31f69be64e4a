"""The drop-in runner of the reference's benchmark definitions (simulation_b200/refbench.py): all eight
1D and four 2D ``test_*`` programs, at reduced sizes, print the lines the reference programs print.

Goldens (tests/golden/make_golden.py, produced by EXECUTING the reference programs with only their size literals
reduced): ``twin_fd1d_*`` (numpy, bit-stable: the printed ``ex[0:50]`` text must be identical), ``benchdef_numpy_3_*``
(numpy 2D twins: identical text) and ``benchdef_numba_3_*`` (the numba twins BASELINE.md quotes: fastmath, float32 arrays
evaluated through float64 literals -- compared within the north star's 1e-5 of peak |Ez|, incl. the amplitude line of
3_4).  Each case runs twice: on the CPU emulator of the kernels (always) and on the GPU (``-m gpu``)."""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")


@pytest.fixture(params=["emulated", pytest.param("cuda", marks=pytest.mark.gpu)])
def where(request, monkeypatch):
    if request.param == "emulated":
        from tests.emu import device
        device.install(monkeypatch)
        return "cpu"
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return "cuda"


@pytest.mark.parametrize("prog", ["1_1", "1_2", "1_3", "1_4", "1_5", "2_1", "2_2", "2_3"])
def test_1d_definitions_print_the_reference_lines(prog, where):
    from simulation_b200 import refbench
    ref = cases.golden(f"twin_fd1d_{prog}")
    nx, ns = int(ref["nx"]), int(ref["ns"])
    r = refbench.run(prog, ns, nx=nx, device=where, warm=False)
    lines = refbench.report(r)
    assert lines[0].startswith("Total compute time on GPU: ") and lines[0].endswith(" s") and len(lines) == 2
    assert lines[1] == str(ref["ex"][0:50])                      # what fd1d/program/test_<prog>.py prints
    assert r["ex"].tobytes() == ref["ex"].tobytes()
    if prog in ("2_2", "2_3"):                                   # amplitude(..., amplt[2]) of the reference's post-processing
        assert np.array_equal(r["amplt"][2], ref["amplt2"], equal_nan=True)


@pytest.mark.parametrize("prog", ["3_1", "3_2", "3_3"])
def test_2d_numpy_definitions_print_the_reference_lines(prog, where):
    from simulation_b200 import refbench
    ref = cases.golden(f"benchdef_numpy_{prog}")
    nx, ny, ns, npml = (int(ref[k]) for k in ("nx", "ny", "ns", "npml"))
    r = refbench.run(prog, ns, nx=nx, ny=ny, npml=npml, device=where, warm=False)
    lines = refbench.report(r)
    assert len(lines) == 2 and lines[0].startswith("Total compute time on GPU: ")
    if prog == "3_1":            # free space: the reference's `+=` form and the PML form differ in the sign of zeros only
        assert np.array_equal(r["ez"], ref["ez"]) and np.array_equal(r["lines"][0], ref["printed_ez"])
    else:
        assert lines[1] == str(ref["text_ez"])                   # what fd2d/program/test_<prog>.py prints, character for character
        assert r["ez"].tobytes() == ref["ez"].tobytes()


@pytest.mark.parametrize("prog", ["3_1", "3_2", "3_3", "3_4"])
def test_2d_numba_definitions_within_tolerance(prog, where):
    """fd2d/python/test_3_*.py (the definitions BASELINE.md quotes; the only home of 3_4)."""
    from simulation_b200 import refbench
    ref = cases.golden(f"benchdef_numba_{prog}")
    nx, ny, ns, npml = (int(ref[k]) for k in ("nx", "ny", "ns", "npml"))
    kw = {"radius": float(ref["radius"])} if prog == "3_4" else {}
    r = refbench.run(prog, ns, nx=nx, ny=ny, npml=npml, device=where, warm=False, **kw)
    peak = float(np.abs(ref["ez"]).max())
    assert peak > 1e-3
    assert float(np.abs(r["ez"].astype(np.float64) - ref["ez"]).max()) <= 1e-5 * peak
    assert float(np.abs(r["lines"][0].astype(np.float64) - ref["printed_ez"]).max()) <= 1e-5 * peak     # ez[2][0:50]
    assert len(refbench.report(r)) == (3 if prog == "3_4" else 2)
    if prog == "3_4":            # print(amplt[2][0:ny-50]), fd2d/python/test_3_4.py:293
        got, want = r["lines"][1], ref["printed_amplt"]
        assert got.shape == want.shape == (ny - 50,)
        assert float(np.abs(got.astype(np.float64) - want).max()) <= 1e-4 * float(np.abs(want).max())


@pytest.mark.reference
def test_1d_runner_against_the_reference_program_run_live(monkeypatch):
    """In the build container: execute fd1d/program/test_1_5.py itself (size literals reduced) and compare its stdout
    line with the runner's, character for character."""
    from tests.emu import device
    from tests.golden import make_golden
    from simulation_b200 import refbench
    device.install(monkeypatch)
    nx, ns = 400, 700
    _, printed = make_golden.run_twin("fd1d/program/test_1_5.py", nx, ns)
    r = refbench.run("1_5", ns, nx=nx, device="cpu", warm=False)
    assert refbench.report(r)[1] == str(printed[1][0])
