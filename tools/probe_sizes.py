"""Throughput of advance() with the library's own launch plan over grid sizes and programs (device-resident)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from simulation_b200 import fd2d, surface

def build(prog, n, npml):
    if prog == "3_2":
        return fd2d.Fdtd2D(n, n, npml, np.float32, source=fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6)))
    if prog == "3_3":
        return fd2d.Fdtd2D(n, n, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)))
    naz, nbz = surface.dielectric_cylinder(n, n, npml, int(n * 0.15), surface.DT, 30.0, 0.30, np.float32)
    freqs = np.array([50e6, 300e6, 700e6]) if prog.endswith("dft") else None       # program 3_4 with its running DFT
    return fd2d.Fdtd2D(n, n, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=naz, nbz=nbz, freqs=freqs)

for prog, n, npml, steps in [("3_2", 1024, 80, 1200), ("3_3", 1024, 80, 1200), ("3_4", 1024, 80, 1200), ("3_2", 2048, 80, 600),
                             ("3_2", 4096, 80, 240), ("3_4", 4096, 80, 240), ("3_4dft", 1024, 80, 600), ("3_4dft", 4096, 80, 120), ("3_2", 8192, 80, 96), ("3_2", 16384, 80, 48),
                             ("3_2", 32768, 80, 96)]:
    sim = build(prog, n, npml)
    sim.advance(24); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.advance(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{prog} {n:6d}^2 npml={npml}: {n * n * steps / ms / 1e6:8.1f} Gcell/s  {ms / steps * 1e3:9.1f} us/step", flush=True)
    del sim
    torch.cuda.empty_cache()
