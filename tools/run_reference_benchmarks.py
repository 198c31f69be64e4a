"""Run the reference's own benchmark definitions (its `test_*` twins) on the GPU path and
print what they print: total compute time and the first 50 field values.

  1D: nx=38000, ns=40000, fp32, prints ex[0:50]        (fd1d/program/test_1_1.py:34-53 ... test_2_3.py)
  2D: 1024x1024, ns=5000, npml=80, fp32, prints ez[2][0:50]   (fd2d/python/test_3_3.py:145-199)

plus the BASELINE.json configs 2 and 4 (1e6-cell lossy line, 4096^2 TFSF + lossy cylinder) as timing lines.
    python tools/run_reference_benchmarks.py [--quick]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd1d, fd2d, surface  # noqa: E402

DT = surface.DT


def timed(sim, ns):
    sim.advance(min(ns, 64))                    # warm-up (module load, clocks); then restart from zero state
    torch.cuda.synchronize()
    return sim


def run_line(prog, nx, ns):
    S = fd1d.LineSource
    kw = {}
    if prog in ("1_1", "1_2"):
        kw = dict(abc=(prog == "1_2"), source=S(nx // 2, surface.Gaussian(40, 12.0), hard=True))
    elif prog in ("1_3", "1_4", "1_5"):
        ca, cb = surface.dielectric_fdtd(nx, DT, 4.0, 0.04 if prog == "1_5" else 0.0, np.float32)
        kw = dict(source=S(1, surface.Gaussian(40, 12.0) if prog == "1_3" else surface.Sinusoid(700e6)),
                  ca=ca if prog == "1_5" else None, cb=cb)
    elif prog == "2_1":
        nax, nbx, _, _ = surface.dielectric_flux(nx, DT, 4.0, 0.04, np.float32)
        kw = dict(form="flux", source=S(1, surface.Sinusoid(700e6), field="dx"), nax=nax, nbx=nbx)
    elif prog == "2_3":
        nax, nbx, ncx, ndx = surface.dielectric_flux(nx, DT, 2.0, 0.01, np.float32, chi=2.0, tau=0.001e-6)
        kw = dict(form="flux", source=S(1, surface.Gaussian(50, 10.0), field="dx"), nax=nax, nbx=nbx, ncx=ncx, ndx=ndx)
    warm = fd1d.Fdtd1D(nx, np.float32, tblock=64, **kw)
    warm.advance(128)
    torch.cuda.synchronize()
    sim = fd1d.Fdtd1D(nx, np.float32, tblock=64, **kw)
    t0 = time.perf_counter()
    sim.advance(ns)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"fd1d test_{prog}: nx={nx} ns={ns}  Total compute time on GPU: {dt:.3f} s  "
          f"({nx * ns / dt / 1e9:.2f} Gcell-updates/s)")
    return sim.get("ex")


def run_grid(prog, n, ns, npml=80, radius=1.5):
    if prog == "3_2":
        src = fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6))
        mk = lambda: fd2d.Fdtd2D(n, n, npml, np.float32, source=src)
    elif prog == "3_3":
        mk = lambda: fd2d.Fdtd2D(n, n, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)))
    else:
        rgrid = int(radius / 0.01 - 1)
        naz, nbz = surface.dielectric_cylinder(n, n, npml, rgrid, DT, 30.0, 0.30, np.float32)
        mk = lambda: fd2d.Fdtd2D(n, n, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=naz, nbz=nbz)
    warm = mk()
    warm.advance(24)
    torch.cuda.synchronize()
    del warm
    sim = mk()
    t0 = time.perf_counter()
    sim.advance(ns)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"fd2d test_{prog}: {n}x{n} ns={ns} npml={npml}  Total compute time on GPU: {dt:.3f} s  "
          f"({n * n * ns / dt / 1e9:.2f} Gcell-updates/s)")
    return sim.get("ez")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    np.set_printoptions(linewidth=120)
    ns1, ns2 = (4000, 500) if a.quick else (40000, 5000)
    for prog in ("1_1", "1_2", "1_5", "2_1", "2_3"):
        ex = run_line(prog, 38000, ns1)
        print(ex[0:50])
    for prog in ("3_2", "3_3", "3_4"):
        ez = run_grid(prog, 1024, ns2)
        print(ez[2][0:50])
    # BASELINE.json configs 2 and 4
    nx = 1_000_000
    ca, cb = surface.dielectric_fdtd(nx, DT, 4.0, 0.04, np.float32, start=nx // 2, stop=nx // 2 + nx // 4)
    sim = fd1d.Fdtd1D(nx, np.float32, tblock=64, source=fd1d.LineSource(1, surface.Sinusoid(700e6)), ca=ca, cb=cb)
    sim.advance(128); torch.cuda.synchronize()
    ns = 1000 if a.quick else 10000
    t0 = time.perf_counter(); sim.advance(ns); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"config 2: 1D lossy slab nx=1e6 ns={ns}: {dt:.3f} s  ({nx * ns / dt / 1e9:.2f} Gcell-updates/s, "
          f"{24 * nx * ns / dt / 1e9:.0f} GB/s algorithmic)")
    run_grid("3_4", 4096, 200 if a.quick else 2000, radius=6.0)


if __name__ == "__main__":
    main()
