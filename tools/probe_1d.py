import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd1d, surface
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 38000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 6400
ca, cb = surface.dielectric_fdtd(nx, surface.DT, 4.0, 0.04, np.float32)
sim = fd1d.Fdtd1D(nx, np.float32, tblock=64, source=fd1d.LineSource(1, surface.Sinusoid(700e6)), ca=ca, cb=cb)
sim.advance(128); torch.cuda.synchronize()
t0 = time.perf_counter(); sim.advance(ns); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"nx={nx} ns={ns}: {dt*1e3:.2f} ms, {dt/ns*1e6:.3f} us/step, {nx*ns/dt/1e9:.2f} Gcell/s")
