"""Sweep of row-block plans for Fdtd2D.run_streamed on the BASELINE config-5 grid (32768^2 fp32, K steps): uniform block
heights against ramped plans (small blocks first so stepping starts early, tall blocks in the middle for efficient
launches, small blocks last so the download tail is short).   python tools/probe_stream_plans.py [K]"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simulation_b200 import fd2d, surface  # noqa: E402

n = 32768
K = int(sys.argv[1]) if len(sys.argv) > 1 else 96
src = fd2d.PointSource(n // 2 - 5, n // 2 - 5, surface.Sinusoid(1500e6), hard=True)
sim = fd2d.Fdtd2D(n, n, 80, np.float32, source=src, tblock=6)
host_naz = torch.ones((n, n), dtype=torch.float32).pin_memory()
host_ez = torch.empty((n, n), dtype=torch.float32).pin_memory()


def ramp(head, body, tail):
    left = n - sum(head) - sum(tail)
    nb = max(1, round(left / body))
    return list(head) + [left // nb] * nb + list(tail)


plans = {
    "default": None,
    "uniform 1024": 1024,
    "uniform 512": 512,
    "uniform 2048": 2048,
    "uniform 4096": 4096,
    "ramp 512,1024,2048 | 4096 | 2048,1024,512": ramp((512, 1024, 2048), 4096, (2048, 1024, 512)),
    "ramp 512,1024 | 2048 | 1024,512": ramp((512, 1024), 2048, (1024, 512)),
    "ramp 512,1024,2048 | 3072 | 2048,1024,512": ramp((512, 1024, 2048), 3072, (2048, 1024, 512)),
    "ramp 256,512,1024,2048 | 4096 | 2048,1024,512,256": ramp((256, 512, 1024, 2048), 4096, (2048, 1024, 512, 256)),
    "ramp 1024,2048 | 4096 | 2048,1024": ramp((1024, 2048), 4096, (2048, 1024)),
    "ramp 512,1024,2048 | 6144 | 2048,1024,512": ramp((512, 1024, 2048), 6144, (2048, 1024, 512)),
    "ramp 512,1024,2048 | 4096 | 1024": ramp((512, 1024, 2048), 4096, (1024,)),
}

def timed(**kw):
    best = None
    sim._lanes = None                                  # fresh streams for this configuration
    for rep in range(3):
        for f in ("dz", "hx", "hy", "ihx", "ihy", "ez"):
            sim.tensor(f, stored=True).zero_()
        sim.t = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        sim.run_streamed(K, host_naz, host_ez, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or (rep > 0 and ms < best) else best
    return best


print("stream priority range", torch.cuda.Stream.priority_range())
if os.environ.get("PLANS"):
    ref = None
    for schedule, name, plan in [("wavefront", "uniform 1024", 1024)] + [("skewed", k, v) for k, v in plans.items()]:
        best = timed(block_rows=plan, schedule=schedule)
        if ref is None:
            ref = host_ez.clone()
        nblk = len(plan) if isinstance(plan, list) else (-(-n // plan) if plan else 0)
        print(f"{schedule:9s} {name:52s} blocks {nblk:3d}: {best:7.1f} ms  {n * n * K / best / 1e6:6.1f} Gcell/s  ez == wavefront ez: {torch.equal(ref, host_ez)}", flush=True)
for prio in (False, True):
    for S in (4, 8, 16):
        for window in (None, 2, 3, 4, 6):
            best = timed(streams=S, window=window, priorities=prio)
            print(f"skewed default plan, {S:2d} streams, window {window}, priorities {prio}: {best:7.1f} ms  {n * n * K / best / 1e6:6.1f} Gcell/s", flush=True)

# timeline of the default plan: where the time between the kernels' sum and the wall goes
for f in ("dz", "hx", "hy", "ihx", "ihy", "ez"):
    sim.tensor(f, stored=True).zero_()
sim.t = 0
trace = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
sim.run_streamed(K, host_naz, host_ez, trace=trace)
e1.record()
torch.cuda.synchronize()
rows = [(b, p, e0.elapsed_time(a), e0.elapsed_time(z)) for b, p, a, z in trace]
print(f"default plan traced: {e0.elapsed_time(e1):.1f} ms; items {len(rows)}, sum of item durations {sum(z - a for _, _, a, z in rows):.1f} ms, "
      f"first start {min(r[2] for r in rows):.2f}, last end {max(r[3] for r in rows):.2f}")
nb = max(r[0] for r in rows) + 1
for b in sorted({0, 1, 3, nb // 2, nb - 2, nb - 1}):
    it = [r for r in rows if r[0] == b]
    print(f"  block {b:2d}: level 0 starts {it[0][2]:7.2f}, last level ends {it[-1][3]:7.2f}; item durations "
          + " ".join(f"{z - a:.2f}" for _, _, a, z in it))
