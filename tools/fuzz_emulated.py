"""Fuzz of the launch logic on the CPU emulator (test infrastructure, tests/emu): random grid sizes, PML depths, programs,
pass depths, vector widths, row partitions and step counts through Fdtd2D.advance, bit-for-bit against the numpy oracle.
    python tools/fuzz_emulated.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fdtd_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402
from tests.emu import device  # noqa: E402
from tests.test_gpu_fd2d import _assert_same, _sim_for  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
mp = pytest.MonkeyPatch()
emu = device.install(mp)
t_end, n, bad = time.time() + budget, 0, 0
while time.time() < t_end:
    prog = str(rng.choice(["3_1", "3_2", "3_3", "3_4"]))
    big = rng.random() < 0.3
    nx = int(rng.integers(24, 420 if big else 120))
    ny = int(rng.integers(24, 900 if big else 160))
    npml = 0 if prog == "3_1" else int(rng.integers(2, max(3, min(nx, ny) // 2 - 2)))
    if prog in ("3_2",) and (nx // 2 - 5 < 1 or ny // 2 - 5 < 1):
        continue
    tblock = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 8]))
    force_v = int(rng.choice([0, 1, 2, 4]))
    chunk = int(rng.choice([0, 1, 3, 5, 16, 40, 100]))
    careful = int(rng.choice([0, 0, 0, 1, 2]))
    ns = int(rng.integers(1, 26))
    dtype = np.float32 if rng.random() < 0.7 else np.float64
    radius = float(rng.uniform(0.03, 0.4))
    cfg = dict(prog=prog, nx=nx, ny=ny, npml=npml, tblock=tblock, force_v=force_v, chunk=chunk, careful=careful, ns=ns,
               dtype=np.dtype(dtype).name, radius=round(radius, 3))
    try:
        emu.fdtd2d_tune(force_v, chunk, 0, 0, careful)
        sim = _sim_for(prog, nx, ny, dtype, npml=npml, radius=radius, device="cpu")
        parts = [ns] if rng.random() < 0.5 else [ns // 2, ns - ns // 2]
        for part in parts:
            if part <= 6 and rng.random() < 0.3:          # the reference-named kernels, one launch per function
                for _ in range(part):
                    sim.step()
            else:
                sim.advance(part, tblock=tblock or None)
        g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml, radius=radius, dft=False)
        orc.advance_2d(g, src)
        _assert_same(sim, g, prog, exact_zero_sign=(prog != "3_1"))
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", cfg, type(e).__name__, str(e)[:300], flush=True)
    finally:
        emu.fdtd2d_tune(0, 0, 0, 0, 0)
    n += 1
print(f"{n} random configurations, {bad} failures (seed {seed})")
mp.undo()
