"""Tuning sweep of the fused 2D kernel on one GPU: vector width, time-block depth, warps per CTA, rows per
chunk.  Prints one line per configuration (Gcell-updates/s, device-resident, CUDA events).
    python tools/sweep_march.py [--size 32768] [--steps 24]"""
import argparse
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import _lib, fd2d, surface  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--vs", default="4,2")
    ap.add_argument("--ts", default="4,3,2,1")
    ap.add_argument("--warps", default="4,2,8")
    ap.add_argument("--chunks", default="0,256,1024")
    ap.add_argument("--rings", default="4")
    ap.add_argument("--careful", type=int, default=0)
    ap.add_argument("--prog", default="3_2")
    ap.add_argument("--npml", type=int, default=80)
    ap.add_argument("--dtype", default="float32")
    a = ap.parse_args()
    n = a.size
    DT = np.float32 if a.dtype == "float32" else np.float64
    if a.prog == "3_2":
        sim = fd2d.Fdtd2D(n, n, a.npml, DT, source=fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6)))
    elif a.prog == "3_3":
        sim = fd2d.Fdtd2D(n, n, a.npml, DT, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)))
    else:
        naz, nbz = surface.dielectric_cylinder(n, n, a.npml, int(n * 0.15), surface.DT, 30.0, 0.30, DT)
        sim = fd2d.Fdtd2D(n, n, a.npml, DT, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=naz, nbz=nbz)
    lib = _lib.lib()
    ints = lambda s: [int(x) for x in s.split(",")]
    for T, V, W, C, R in itertools.product(ints(a.ts), ints(a.vs), ints(a.warps), ints(a.chunks), ints(a.rings)):
        lib.fdtd2d_tune(V, C, W, R, a.careful)
        steps = max(1, a.steps // T) * T
        sim.advance(2 * T, tblock=T)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.advance(steps, tblock=T)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"T={T} V={V} warps={W} chunk={C:5d} ring={R}  {n * n * steps / ms / 1e6:8.1f} Gcell/s  {ms / steps:7.3f} ms/step", flush=True)
    lib.fdtd2d_tune(0, 0, 0, 0, 0)


if __name__ == "__main__":
    main()
