"""1D running DFT: the DFT-carrying fused pass (k1_advance_dft) against the per-step path (dxfield, exfield, fourier,
hyfield kernels) on the same line -- time and bitwise equality.   python tools/probe_1d_dft.py [nx] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd1d, surface  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
freqs = np.array((100e6, 200e6, 500e6), dtype=np.float32)
nax, nbx, _, _ = surface.dielectric_flux(nx, surface.DT, 4.0, 0.0, np.float32)
mk = lambda: fd1d.Fdtd1D(nx, np.float32, form="flux", source=fd1d.LineSource(1, surface.Gaussian(50, 10.0), field="dx"),
                         nax=nax, nbx=nbx, freqs=freqs)
out = {}
for name, fused, steps in (("fused", True, ns), ("per-step", False, max(ns // 10, 1))):
    warm = mk(); warm.advance(32, fused_dft=fused); torch.cuda.synchronize(); del warm
    sim = mk()
    t0 = time.perf_counter()
    sim.advance(steps, fused_dft=fused)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"1D flux line + running DFT (3 frequencies), nx={nx}: {name:8s} {steps} steps in {dt:.4f} s = "
          f"{nx * steps / dt / 1e9:.1f} Gcell-updates/s")
    out[name] = sim
a, b = mk(), mk()
a.advance(200); b.advance(200, fused_dft=False)
print("bitwise equal after 200 steps:", all(a.get(n).tobytes() == b.get(n).tobytes() for n in ("ex", "hy", "dx", "ix", "r_pt", "i_pt", "r_in", "i_in")))
