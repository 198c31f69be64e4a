"""Run the reference's own benchmark definitions (its `test_*` twins) on the GPU path and
print what they print: total compute time and the first 50 field values (3_4: the amplitude line as well).

  1D: all eight programs 1_1 .. 2_3, nx=38000, ns=40000, fp32, prints ex[0:50]   (fd1d/program/test_1_1.py:34-53 ...)
  2D: all four programs 3_1 .. 3_4, 1024x1024, ns=5000, npml=80, fp32, prints ez[2][0:50]; 3_4 also amplt[2][0:ny-50]
      (fd2d/python/test_3_3.py:145-199, test_3_4.py:289-293)

plus the BASELINE.json configs 2 and 4 (1e6-cell lossy line, 4096^2 TFSF + lossy cylinder) as timing lines.
    python tools/run_reference_benchmarks.py [--quick] [--only 3_4,2_2]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd1d, refbench, surface  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="a tenth of the reference's step counts")
    ap.add_argument("--only", default="", help="comma-separated program ids (default: all twelve)")
    a = ap.parse_args()
    np.set_printoptions(linewidth=120)
    only = [x for x in a.only.split(",") if x]
    ns1, ns2 = (4000, 500) if a.quick else (40000, 5000)
    for prog in refbench.PROGRAMS_1D + refbench.PROGRAMS_2D:
        if only and prog not in only:
            continue
        one_d = prog in refbench.PROGRAMS_1D
        r = refbench.run(prog, ns1 if one_d else ns2)
        where = f"fd1d test_{prog}: nx=38000" if one_d else f"fd2d test_{prog}: 1024x1024 npml={0 if prog == '3_1' else 80}"
        print(f"== {where} ns={r['ns']}  ({r['cells'] * r['ns'] / r['seconds'] / 1e9:.2f} Gcell-updates/s)")
        for line in refbench.report(r):
            print(line)
    if only:
        return
    # BASELINE.json configs 2 and 4
    nx = 1_000_000
    ca, cb = surface.dielectric_fdtd(nx, surface.DT, 4.0, 0.04, np.float32, start=nx // 2, stop=nx // 2 + nx // 4)
    sim = fd1d.Fdtd1D(nx, np.float32, tblock=64, source=fd1d.LineSource(1, surface.Sinusoid(700e6)), ca=ca, cb=cb)
    sim.advance(128); torch.cuda.synchronize()
    ns = 1000 if a.quick else 10000
    t0 = time.perf_counter(); sim.advance(ns); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"config 2: 1D lossy slab nx=1e6 ns={ns}: {dt:.3f} s  ({nx * ns / dt / 1e9:.2f} Gcell-updates/s, "
          f"{24 * nx * ns / dt / 1e9:.0f} GB/s algorithmic)")
    r = refbench.run("3_4", 200 if a.quick else 2000, nx=4096, ny=4096, radius=6.0, dft=False)
    print(f"config 4: 4096x4096 TFSF + lossy cylinder ns={r['ns']}: {r['seconds']:.3f} s  "
          f"({r['cells'] * r['ns'] / r['seconds'] / 1e9:.2f} Gcell-updates/s)")


if __name__ == "__main__":
    main()
