"""BASELINE config 4 (4096^2 fp32, npml 80, TFSF plane wave, lossy dielectric cylinder eps_r=30 sigma=0.3 radius 6 m) and
larger grids with the same cylinder: fused advance with the lossless-outside split (interior warps outside the
cylinder's box run the lossless kernel) against the lossy kernel on every warp.   python tools/probe_c4_split.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import _lib, fd2d, surface  # noqa: E402

for n, ns in ((4096, 600), (8192, 240), (16384, 96)):
    rgrid = int(6.0 / 0.01 - 1)
    md = fd2d.dielectric(n, n, 80, rgrid, surface.DT, 30.0, 0.30, np.float32)
    res = {}
    for name, flag in (("lossy kernel everywhere", 2), ("lossless-outside split", 0)):
        _lib.lib().fdtd2d_tune(0, 0, 0, 0, flag)
        sim = fd2d.Fdtd2D(n, n, 80, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=md.naz, nbz=md.nbz)
        sim.advance(24)
        torch.cuda.synchronize()
        best = None
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sim.advance(ns)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        res[name] = sim
        print(f"{n}^2 TFSF + lossy cylinder (box {sim._lossy_box()}), {name:24s}: {best / ns * 1e3:8.1f} us/step  {n * n * ns / best / 1e6:7.1f} Gcell-updates/s", flush=True)
    _lib.lib().fdtd2d_tune(0, 0, 0, 0, 0)
    a, b = res.values()
    print("   bitwise equal:", all(torch.equal(a.tensor(f), b.tensor(f)) for f in ("dz", "ez", "iz", "hx", "hy", "ihx", "ihy")), flush=True)
    del res, a, b, sim, md
    torch.cuda.empty_cache()
