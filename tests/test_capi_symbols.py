"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every
symbol include/fdtd_b200.h declares; the ctypes prototypes cover them; struct layouts agree with the C
compiler's; argument validation answers without touching a device."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fdtd_b200.h")


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\s*\*|int)\s*(fdtd\w*)\s*\(", text, flags=re.M)))


def test_header_declares_the_reference_protocol():
    names = declared_functions()
    for fn in ("ezinct", "dfield", "inctdz", "efield", "hxinct", "hfield", "incthx", "incthy"):
        assert f"fdtd2d_{fn}" in names
    for fn in ("exfield", "hyfield", "dxfield", "exfield_flux", "advance"):
        assert f"fdtd1d_{fn}" in names
    assert "fdtd2d_advance" in names and "fdtd_last_error" in names


def test_library_exports_every_declared_symbol():
    from simulation_b200 import _lib
    h = _lib.lib()
    for name in declared_functions():
        assert hasattr(h, name), f"{name} declared in fdtd_b200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from simulation_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "fdtd_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(fdtd_pmlayer),sizeof(fdtd_medium2d),sizeof(fdtd_medium1d),sizeof(fdtd_source),"
                   "sizeof(fdtd1d_problem),sizeof(fdtd2d_problem));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    mine = [C.sizeof(t) for t in (_lib.PmlLayer, _lib.Medium2D, _lib.Medium1D, _lib.Source, _lib.Problem1D, _lib.Problem2D)]
    assert sizes == mine


def test_argument_validation_without_a_device():
    from simulation_b200 import _lib
    h = _lib.lib()
    out = C.c_int(0)
    assert h.fdtd2d_advance(None, 0, 1, None, 1, None, C.byref(out)) == -1
    assert b"null" in h.fdtd_last_error()
    p = _lib.Problem2D()
    p.nx = p.ny = 64
    p.row_hi = p.rows_alloc = 64
    assert h.fdtd2d_advance(C.byref(p), 0, 1, None, 13, None, C.byref(out)) == -1
    assert b"tblock" in h.fdtd_last_error()
    q = _lib.Problem1D()
    q.nx = 2
    assert h.fdtd1d_advance(C.byref(q), 0, 1, None, 1, None, C.byref(out)) == -1
    assert h.fdtd2d_max_tblock(_lib.F32, 1024) >= 1 and h.fdtd2d_max_tblock(7, 1024) == 0
    assert h.fdtd_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from simulation_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfdtd_b200.so")
    with pytest.raises(_lib.FdtdError, match="no CPU fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_host_classes_refuse_to_run_without_cuda():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from simulation_b200 import _lib, fd1d, fd2d
    with pytest.raises(_lib.FdtdError):
        fd2d.Fdtd2D(32, 32, 4)
    with pytest.raises(_lib.FdtdError):
        fd1d.Fdtd1D(32)


def _build_c_example(tmp_path):
    exe = tmp_path / "c_host_3_3"
    subprocess.run(["gcc", "-Wall", "-O1", os.path.join(ROOT, "examples", "c_host_3_3.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", os.path.join(ROOT, "simulation_b200", "csrc"), "-lfdtd_b200", "-lm", "-o", str(exe)], check=True)
    return exe


def test_c_host_example_compiles_and_links(tmp_path):
    """A plain C program shaped like the reference's fd2d/cuda/test_3_3.cu main() links against the C ABI."""
    assert _build_c_example(tmp_path).exists()


@pytest.mark.gpu
def test_c_host_example_runs(tmp_path):
    exe = _build_c_example(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "simulation_b200", "csrc") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe), "600", "760", "150", "24"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical bytes" in r.stdout
