import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from simulation_b200 import fd2d, surface, _lib
n, npml, ns = 4096, 80, 300
rgrid = int(6.0 / 0.01 - 1)
naz, nbz = surface.dielectric_cylinder(n, n, npml, rgrid, surface.DT, 30.0, 0.30, np.float32)
mk = lambda: fd2d.Fdtd2D(n, n, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=naz, nbz=nbz)
a = mk(); a.advance(ns)
b = mk()
for _ in range(ns): b.step()
_lib.lib().fdtd2d_tune(0, 0, 0, 0, 1)
c = mk(); c.advance(ns); torch.cuda.synchronize()
_lib.lib().fdtd2d_tune(0, 0, 0, 0, 0)
for name in ("dz", "ez", "iz", "hx", "hy", "ihx", "ihy"):
    x, y, z = a.tensor(name), b.tensor(name), c.tensor(name)
    print(name, "fused==unfused", bool(torch.equal(x, y)), "max|d|", float((x - y).abs().max()),
          " careful==unfused", bool(torch.equal(z, y)), float((z - y).abs().max()), " peak", float(y.abs().max()))
    if not torch.equal(x, y):
        bad = (x != y).nonzero()
        print("   first bad", bad[:5].tolist(), "count", len(bad))
