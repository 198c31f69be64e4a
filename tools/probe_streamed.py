"""Timing probe of Fdtd2D.run_streamed on the BASELINE config-5 grid: blocks x streams."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from simulation_b200 import fd2d, surface

n, K = 32768, 96
src = fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6), hard=True)
sim = fd2d.Fdtd2D(n, n, 80, np.float32, source=src, tblock=6)
host_naz = torch.ones((n, n), dtype=torch.float32).pin_memory()
host_ez = torch.empty((n, n), dtype=torch.float32).pin_memory()
import os
if "FDTD_FORCE_V" in os.environ:
    from simulation_b200 import _lib
    _lib.lib().fdtd2d_tune(int(os.environ["FDTD_FORCE_V"]), int(os.environ.get("FDTD_CHUNK_ROWS", "0")), 0, 0, 0)
combos = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(24, 1), (24, 2), (24, 4)]
for blocks, streams in combos:
    for rep in range(2):
        for name in ("dz", "hx", "hy", "ihx", "ihy", "ez"):
            sim.tensor(name, stored=True).zero_()
        sim.t = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        h0 = time.perf_counter()
        trace = [] if os.environ.get("TRACE") else None
        sim.run_streamed(K, host_naz, host_ez, blocks=blocks, streams=streams, trace=trace, block_rows=(blocks if blocks > 100 else None))
        e1.record()
        host_ms = 1e3 * (time.perf_counter() - h0)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"blocks={blocks} streams={streams} rep={rep}: host issue {host_ms:.1f} ms; {ms:.1f} ms  {n * n * K / ms / 1e6:.1f} Gcell/s", flush=True)
        if trace and rep > 0:
            rows = [(b, p, e0.elapsed_time(a), e0.elapsed_time(z)) for b, p, a, z in trace]
            busy = sum(z - a for _, _, a, z in rows)
            print(f"  items {len(rows)}  sum of item durations {busy:.1f} ms  first start {min(r[2] for r in rows):.1f}  last end {max(r[3] for r in rows):.1f}")
            for b, p, a, z in rows[len(rows) // 2:len(rows) // 2 + 16:2]:
                print(f"  b={b:2d} p={p:2d}  start {a:7.2f}  end {z:7.2f}  dur {z - a:5.2f}")
