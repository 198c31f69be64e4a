"""Timing probe: advance(K) for several K on the BASELINE config-5 grid (is there a fixed cost per call?)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from simulation_b200 import fd2d, surface

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
src = fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6), hard=True)
sim = fd2d.Fdtd2D(n, n, 80, np.float32, source=src, tblock=6)
sim.advance(12); torch.cuda.synchronize()
for K in (6, 24, 96, 96, 384):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0.record(); sim.advance(K); e1.record()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"K={K}: {ms:.2f} ms  {ms / K:.4f} ms/step  host call {1e3 * (t1 - t0):.2f} ms  {n * n * K / ms / 1e6:.1f} Gcell/s", flush=True)
