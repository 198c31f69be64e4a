import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd1d, surface
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 38000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 6400
ca, cb = surface.dielectric_fdtd(nx, surface.DT, 4.0, 0.04, np.float32)
sim = fd1d.Fdtd1D(nx, np.float32, tblock=64, source=fd1d.LineSource(1, surface.Sinusoid(700e6)), ca=ca, cb=cb)
sim.advance(128); torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); sim.advance(ns); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"nx={nx} ns={ns}: host call {1e3*(t1-t0):.2f} ms, until sync {1e3*(t2-t0):.2f} ms, GPU events {e0.elapsed_time(e1):.2f} ms "
          f"-> {e0.elapsed_time(e1)/ns*1e3:.3f} us/step")
t0 = time.perf_counter(); tab = sim.source.waveform.table(1, ns); print("source table", 1e3*(time.perf_counter()-t0), "ms")
