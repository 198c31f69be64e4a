"""Where does run_streamed spend its time?  (a) plain advance, (b) the block-wavefront schedule with DEVICE
'host' tensors (no PCIe), (c) the real thing with pinned host tensors, (d) the two copies alone."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd2d, surface

n, K = 32768, 96
src = fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6))
sim = fd2d.Fdtd2D(n, n, 80, np.float32, source=src, tblock=6)

def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

sim.advance(12)
print("plain advance(96):            %.1f ms" % timed(lambda: sim.advance(K)))
dn = torch.ones((n, n), dtype=torch.float32, device="cuda"); de = torch.empty_like(dn)
for B in (8, 16):
    print("wavefront, device tensors B=%2d: %.1f ms" % (B, timed(lambda: sim.run_streamed(K, dn, de, blocks=B))))
del dn, de
hn = torch.ones((n, n), dtype=torch.float32).pin_memory(); he = torch.empty((n, n), dtype=torch.float32).pin_memory()
print("H2D 4 GiB alone:              %.1f ms" % timed(lambda: sim.naz.copy_(hn, non_blocking=True)))
print("D2H 4 GiB alone:              %.1f ms" % timed(lambda: he.copy_(sim.tensor("ez"), non_blocking=True)))
for B in (8, 12, 16, 20, 24):
    print("streamed, pinned host B=%2d:    %.1f ms" % (B, timed(lambda: sim.run_streamed(K, hn, he, blocks=B))))
