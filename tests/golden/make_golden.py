"""Generate the golden vectors in this directory by EXECUTING THE REFERENCE (dsarvan/simulation).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box has no /root/reference; it uses the committed .npz files.

Three families:
  main_*      the program's own ``main()`` as shipped (fp64 book-size demo); the arrays handed to its
              plot helpers are recorded.  sha256 prefixes agree with SURVEY.md 8(c).
  twin_*      the fp32 ``test_*`` benchmark twin of a 1D program whose loop is inlined in main():
              the module source is executed with ONLY the ``nx`` / ``ns`` literals reduced.
  drive_*     the reference's module-level step functions called in the program's own order on
              small non-square grids, fp32 and fp64, recording ALL state arrays.
"""
from __future__ import annotations

import hashlib
import os
import re
import sys
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refload  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    digest = {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()[:16]
              for k, v in arrays.items() if isinstance(v, np.ndarray) and v.ndim}
    print(f"{name}: {os.path.getsize(path)} B  {digest}")


# ------------------------------------------------------------------ main_* (as shipped)
def golden_mains():
    for prog in ("1_1", "1_2", "1_3", "1_4", "1_5", "2_1", "2_2", "2_3"):
        seen = refload.run_main(f"fd1d/program/fd1d_{prog}.py")
        out = {"ex": [a for a in seen["visualize"] if isinstance(a, np.ndarray)][-1]}
        if "amplitude" in seen:
            out["amplt2"] = [a for a in seen["amplitude"] if isinstance(a, np.ndarray)][-1]
        save(f"main_fd1d_{prog}", **out)
    for prog in ("3_1", "3_2", "3_3"):
        seen = refload.run_main(f"fd2d/program/fd2d_{prog}.py")
        save(f"main_fd2d_{prog}", ez=seen["surfaceplot"][-1])
    for prog in ("3_1", "3_2", "3_3", "3_4"):            # numba, fastmath: tolerance goldens
        seen = refload.run_main(f"fd2d/python/fd2d_{prog}.py")
        out = {"ez": seen["surfaceplot"][-1]}
        if "amplitudeplot" in seen:
            out["amplt2"] = seen["amplitudeplot"][-1]
        save(f"main_numba_fd2d_{prog}", **out)


# ------------------------------------------------------------------ twin_* (fp32, shrunk literals)
def run_twin(relpath, nx, ns):
    src = open(os.path.join(refload.REFERENCE_ROOT, relpath)).read()
    src, n1 = re.subn(r"nx: int = \d+", f"nx: int = {nx}", src)
    src, n2 = re.subn(r"ns: int = \d+", f"ns: int = {ns}", src)
    assert n1 == 1 and n2 == 1, (relpath, n1, n2)
    refload.load("fd1d/program/fd1d_1_1.py")             # installs the matplotlib stubs
    glb = {"__name__": "twin"}
    exec(compile(src, relpath, "exec"), glb)
    seen = {}
    for fn in ("visualize", "amplitude"):
        if fn in glb:
            glb[fn] = (lambda name: (lambda *a, **k: seen.__setitem__(name, a)))(fn)
    printed = []
    glb["print"] = lambda *a, **k: printed.append(a)
    glb["main"]()
    return seen, printed


def golden_twins():
    for prog in ("1_1", "1_2", "1_3", "1_4", "1_5", "2_1", "2_2", "2_3"):
        nx, ns = 600, 1500
        seen, printed = run_twin(f"fd1d/program/test_{prog}.py", nx, ns)
        ex = [a for a in seen["visualize"] if isinstance(a, np.ndarray)][-1]
        assert ex.dtype == np.float32 and ex.shape == (nx,)
        out = {"ex": ex, "nx": np.int64(nx), "ns": np.int64(ns)}
        if "amplitude" in seen:
            out["amplt2"] = [a for a in seen["amplitude"] if isinstance(a, np.ndarray)][-1]
        save(f"twin_fd1d_{prog}", **out)


# ------------------------------------------------------------------ drive_* (all state arrays)
def _pml(mod, nx, ny, npml, dt):
    pml = mod.pmlayer(
        fx1=np.full(nx, 0.0, dtype=dt), fx2=np.full(nx, 1.0, dtype=dt), fx3=np.full(nx, 1.0, dtype=dt),
        fy1=np.full(ny, 0.0, dtype=dt), fy2=np.full(ny, 1.0, dtype=dt), fy3=np.full(ny, 1.0, dtype=dt),
        gx2=np.full(nx, 1.0, dtype=dt), gx3=np.full(nx, 1.0, dtype=dt),
        gy2=np.full(ny, 1.0, dtype=dt), gy3=np.full(ny, 1.0, dtype=dt))
    mod.pmlparam(nx, ny, npml, pml)
    return pml


def drive_3_1(nx, ny, ns, dt):
    m = refload.load("fd2d/program/fd2d_3_1.py")
    dz, ez, hx, hy = (np.zeros((nx, ny), dtype=dt) for _ in range(4))
    naz = np.ones((nx, ny), dtype=dt)
    for t in np.arange(1, ns + 1).astype(np.int32):
        m.dfield(t, nx, ny, dz, hx, hy)
        m.efield(nx, ny, naz, dz, ez)
        m.hfield(nx, ny, ez, hx, hy)
    return dict(dz=dz, ez=ez, hx=hx, hy=hy)


def drive_3_2(nx, ny, npml, ns, dt, naz=None):
    m = refload.load("fd2d/program/fd2d_3_2.py")
    dz, ez, hx, hy, ihx, ihy = (np.zeros((nx, ny), dtype=dt) for _ in range(6))
    naz = np.ones((nx, ny), dtype=dt) if naz is None else naz
    pml = _pml(m, nx, ny, npml, dt)
    for t in np.arange(1, ns + 1).astype(np.int32):
        m.dfield(t, nx, ny, pml, dz, hx, hy)
        m.efield(nx, ny, naz, dz, ez)
        m.hfield(nx, ny, pml, ez, ihx, ihy, hx, hy)
    return dict(dz=dz, ez=ez, hx=hx, hy=hy, ihx=ihx, ihy=ihy)


def drive_3_3(nx, ny, npml, ns, dt):
    m = refload.load("fd2d/program/fd2d_3_3.py")
    ezi, hxi = np.zeros(ny, dtype=dt), np.zeros(ny, dtype=dt)
    dz, ez, hx, hy, ihx, ihy = (np.zeros((nx, ny), dtype=dt) for _ in range(6))
    naz, bc = np.ones((nx, ny), dtype=dt), np.zeros(4, dtype=dt)
    pml = _pml(m, nx, ny, npml, dt)
    for t in np.arange(1, ns + 1).astype(np.int32):
        m.ezinct(ny, ezi, hxi, bc)
        m.dfield(t, nx, ny, pml, ezi, dz, hx, hy)
        m.inctdz(nx, ny, npml, hxi, dz)
        m.efield(nx, ny, naz, dz, ez)
        m.hxinct(ny, ezi, hxi)
        m.hfield(nx, ny, pml, ez, ihx, ihy, hx, hy)
        m.incthx(nx, ny, npml, ezi, hx)
        m.incthy(nx, ny, npml, ezi, hy)
    return dict(dz=dz, ez=ez, hx=hx, hy=hy, ihx=ihx, ihy=ihy, ezi=ezi, hxi=hxi, bc=bc)


def drive_3_4(nx, ny, npml, ns, dt, radius):
    """numba program 3_4 (lossy cylinder + TFSF + DFT); float64 only: numba specialises per dtype
    and its fp32 arithmetic promotes through float64 literals."""
    m = refload.load("fd2d/python/fd2d_3_4.py")
    ezi, hxi = np.zeros(ny, dtype=dt), np.zeros(ny, dtype=dt)
    dz, ez, iz, hx, hy, ihx, ihy = (np.zeros((nx, ny), dtype=dt) for _ in range(7))
    bc = np.zeros(4, dtype=dt)
    pml = _pml(m, nx, ny, npml, dt)
    ds = 0.01
    dtime = ds / 6e8
    rgrid = int(radius / ds - 1)
    md = m.dielectric(nx, ny, npml, rgrid, dtime, 30.0, 0.30)
    freq = np.array((50e6, 300e6, 700e6), dtype=dt)
    nf = len(freq)
    ft = m.ftrans(r_pt=np.zeros((nf, nx, ny), dtype=dt), i_pt=np.zeros((nf, nx, ny), dtype=dt),
                  r_in=np.zeros(nf, dtype=dt), i_in=np.zeros(nf, dtype=dt))
    for t in np.arange(1, ns + 1).astype(np.int32):
        m.ezinct(ny, ezi, hxi, bc)
        m.dfield(t, nx, ny, pml, ezi, dz, hx, hy)
        m.inctdz(nx, ny, npml, hxi, dz)
        m.efield(nx, ny, md, dz, iz, ez)
        m.fourier(t, nf, nx, ny, dtime, freq, ezi, ez, ft)
        m.hxinct(ny, ezi, hxi)
        m.hfield(nx, ny, pml, ez, ihx, ihy, hx, hy)
        m.incthx(nx, ny, npml, ezi, hx)
        m.incthy(nx, ny, npml, ezi, hy)
    return dict(dz=dz, ez=ez, iz=iz, hx=hx, hy=hy, ihx=ihx, ihy=ihy, ezi=ezi, hxi=hxi,
                naz=md.naz, nbz=md.nbz, r_pt=ft.r_pt, i_pt=ft.i_pt, r_in=ft.r_in, i_in=ft.i_in,
                rgrid=np.int64(rgrid))


def golden_drives():
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        save(f"drive_3_1_{tag}", nx=np.int64(40), ny=np.int64(56), ns=np.int64(60),
             **drive_3_1(40, 56, 60, dt))
        save(f"drive_3_2_{tag}", nx=np.int64(56), ny=np.int64(72), npml=np.int64(8), ns=np.int64(130),
             **drive_3_2(56, 72, 8, 130, dt))
        save(f"drive_3_3_{tag}", nx=np.int64(64), ny=np.int64(48), npml=np.int64(7), ns=np.int64(140),
             **drive_3_3(64, 48, 7, 140, dt))
    save("drive_3_4_f64", nx=np.int64(60), ny=np.int64(72), npml=np.int64(8), ns=np.int64(120),
         **drive_3_4(60, 72, 8, 120, np.float64, radius=0.12))
    # randomised medium through the reference 3_2 functions (seeded)
    rng = np.random.default_rng(0)
    naz = rng.uniform(0.25, 1.0, size=(48, 40)).astype(np.float32)
    save("drive_3_2_randnaz_f32", nx=np.int64(48), ny=np.int64(40), npml=np.int64(6), ns=np.int64(90),
         naz=naz, **{k: v for k, v in drive_3_2(48, 40, 6, 90, np.float32, naz=naz).items()})


# ------------------------------------------------------------------ benchdef_* (2D test_* twins, what they PRINT)
def run_twin_2d(relpath, nx, ny, ns, npml=None, radius=None):
    """A 2D ``test_*`` benchmark program executed with ONLY its size literals reduced; returns what it prints after the
    compute-time line (``ez[2][0:50]``, and ``amplt[2][0:ny-50]`` for 3_4) and the arrays it hands to its plot helpers."""
    src = open(os.path.join(refload.REFERENCE_ROOT, relpath)).read()
    subs = [(r"nx: int = \d+", f"nx: int = {nx}"), (r"ny: int = \d+", f"ny: int = {ny}"), (r"ns: int = \d+", f"ns: int = {ns}")]
    if npml is not None:
        subs.append((r"npml: int = \d+", f"npml: int = {npml}"))
    if radius is not None:
        subs.append((r"radius: float = [0-9.]+", f"radius: float = {radius}"))
    for pat, rep in subs:
        src, n = re.subn(pat, rep, src)
        assert n == 1, (relpath, pat, n)
    refload.load("fd1d/program/fd1d_1_1.py")             # installs the matplotlib stubs
    glb = {"__name__": "twin"}
    exec(compile(src, relpath, "exec"), glb)
    seen = {}
    for fn in ("surfaceplot", "contourplot", "amplitudeplot"):
        if fn in glb:
            glb[fn] = (lambda name: (lambda *a, **k: seen.__setitem__(name, a)))(fn)
    printed = []
    glb["print"] = lambda *a, **k: printed.append(a[0])
    glb["main"]()
    return seen, printed


def golden_benchdefs():
    """The reference's 2D benchmark definitions at reduced size: the numpy twins (fd2d/program/test_3_1..3_3.py, bit-stable)
    and the numba twins (fd2d/python/test_3_1..3_4.py: fastmath, float32 arrays evaluated through float64 literals --
    compared within the north-star tolerance)."""
    nx, ny, ns, npml = 168, 200, 260, 20
    for prog in ("3_1", "3_2", "3_3"):
        seen, printed = run_twin_2d(f"fd2d/program/test_{prog}.py", nx, ny, ns, None if prog == "3_1" else npml)
        ez = seen["surfaceplot"][-1]
        assert ez.dtype == np.float32 and ez.shape == (nx, ny) and np.array_equal(printed[1], ez[2][0:50])
        save(f"benchdef_numpy_{prog}", ez=ez, printed_ez=np.asarray(printed[1]), text_ez=np.array(str(printed[1])),
             nx=np.int64(nx), ny=np.int64(ny), ns=np.int64(ns), npml=np.int64(0 if prog == "3_1" else npml))
    for prog in ("3_1", "3_2", "3_3", "3_4"):
        seen, printed = run_twin_2d(f"fd2d/python/test_{prog}.py", nx, ny, ns, None if prog == "3_1" else npml,
                                    0.30 if prog == "3_4" else None)
        ez = seen["surfaceplot"][-1]
        out = dict(ez=ez, printed_ez=np.asarray(printed[1]), nx=np.int64(nx), ny=np.int64(ny), ns=np.int64(ns),
                   npml=np.int64(0 if prog == "3_1" else npml))
        if prog == "3_4":
            out.update(printed_amplt=np.asarray(printed[2]), radius=np.float64(0.30))
        save(f"benchdef_numba_{prog}", **out)


if __name__ == "__main__":
    assert refload.available(), "needs /root/reference"
    if len(sys.argv) > 1 and sys.argv[1] == "benchdefs":
        golden_benchdefs()
        sys.exit(0)
    golden_mains()
    golden_twins()
    golden_benchdefs()
    golden_drives()
