"""Fuzz of the deep passes on the CPU emulator (test infrastructure, tests/emu): the warp-chain interior kernel in every
shape with its column / row variants, the shared-memory-accumulator kernels, the ring careful kernel, short edge chunks --
random grid sizes, PML depths, chunk heights, pass depths and step splits through Fdtd2D.advance at 4-wide vectors,
bit-for-bit against the numpy oracle.
    python tools/fuzz_emulated_deep.py [seconds] [seed]"""
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
    prog = str(rng.choice(["3_2", "3_2", "3_3", "3_1"]))
    nx = int(rng.integers(40, 520))
    ny = 4 * int(rng.integers(30, 330))
    npml = 0 if prog == "3_1" else int(rng.integers(2, max(3, min(nx, ny) // 3)))
    tblock = int(rng.choice([8, 12, 8, 12, 0]))
    chunk = int(rng.choice([0, 8, 13, 24, 40, 64, 100, 200]))
    variant = int(rng.choice([0, 0, 0, 11, 12, 13, 3, 1]))
    fast = int(rng.choice([3, 3, 0, 1, 2]))
    edge = int(rng.choice([1, 1, 0]))
    deep = int(rng.choice([1, 1, 2]))
    ns = int(rng.integers(8, 40))
    cfg = dict(prog=prog, nx=nx, ny=ny, npml=npml, tblock=tblock, chunk=chunk, variant=variant, fast=fast, edge=edge, deep=deep, ns=ns)
    try:
        emu.fdtd2d_tune(4, chunk, 0, 0, 0)
        emu.fdtd2d_tune2(0, deep)
        emu.fdtd2d_tune2(2, variant)
        emu.fdtd2d_tune2(3, edge)
        emu.fdtd2d_tune2(4, fast)
        sim = _sim_for(prog, nx, ny, np.float32, npml=npml, device="cpu")
        parts = [ns] if rng.random() < 0.5 else [ns // 3, ns - ns // 3]
        for part in parts:
            if part:
                sim.advance(part, tblock=tblock or None)
        g, src = cases.grid_program(prog, nx, ny, ns, np.float32, npml=npml, dft=False)
        orc.advance_2d(g, src)
        _assert_same(sim, g, prog, exact_zero_sign=(prog != "3_1"))
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", cfg, type(e).__name__, str(e)[:300], flush=True)
    finally:
        emu.fdtd2d_tune(0, 0, 0, 0, 0)
        emu.fdtd2d_tune2(0, 1)
        emu.fdtd2d_tune2(2, 0)
        emu.fdtd2d_tune2(3, 1)
        emu.fdtd2d_tune2(4, 3)
    n += 1
print(f"{n} random configurations, {bad} failures (seed {seed})")
mp.undo()
