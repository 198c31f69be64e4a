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
t_end, n, bad, n_slab = time.time() + budget, 0, 0, 0
slab_share = float(os.environ.get("FUZZ_SLAB_SHARE", "0.35"))
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
    slabbed = prog != "3_1" and rng.random() < slab_share
    cfg = dict(prog=prog, nx=nx, ny=ny, npml=npml, tblock=tblock, chunk=chunk, variant=variant, fast=fast, edge=edge, deep=deep, ns=ns,
               slabbed=slabbed)
    try:
        emu.fdtd2d_tune(4, chunk, 0, 0, 0)
        emu.fdtd2d_tune2(0, deep)
        emu.fdtd2d_tune2(2, variant)
        emu.fdtd2d_tune2(3, edge)
        emu.fdtd2d_tune2(4, fast)
        if slabbed:
            # the halo exchange fused into the pass: every "GPU" a slab in this process, peers wired by raw pointers,
            # the ring careful kernel carrying the handshake around deep interior passes
            from simulation_b200 import fd2d
            T = tblock or 8
            nslab = int(rng.integers(2, 5))
            if nx // nslab < 2 * T + 2:
                continue
            nblocks = int(rng.integers(1, 4))
            cfg.update(T=T, nslab=nslab, nblocks=nblocks)
            cuts = np.linspace(0, nx, nslab + 1).astype(int)
            slabs = [_sim_for(prog, nx, ny, np.float32, npml=npml, rows=(int(lo), int(hi)), ghost=T, tblock=T, device="cpu")
                     for lo, hi in zip(cuts[:-1], cuts[1:])]
            names = [k for k in fd2d.FIELD_NAMES if k not in ("ez", "iz")]
            words = [np.zeros(8, dtype=np.int64) for _ in slabs]
            for r, sl in enumerate(slabs):
                peer = lambda q: None if q is None else {"row_base": slabs[q].row_base, "sync": words[q].ctypes.data,
                                                         "sets": [{k: slabs[q]._sets[i][k].data_ptr() for k in names} for i in range(2)]}
                sl.p2p = {"halo": T, "sync": type("W", (), {"ptr": words[r].ctypes.data})(),
                          "up": peer(r - 1 if r > 0 else None), "dn": peer(r + 1 if r < nslab - 1 else None)}
            for epoch in range(1, nblocks + 1):
                for r in rng.permutation(nslab):
                    slabs[int(r)].advance(T, tblock=T, lazy_ez=epoch < nblocks, epoch=epoch)
            g, src = cases.grid_program(prog, nx, ny, nblocks * T, np.float32, npml=npml, dft=False)
            orc.advance_2d(g, src)
            for k in names + ["ez"]:
                assert np.concatenate([sl.get(k) for sl in slabs]).tobytes() == getattr(g, k).tobytes(), k
            n += 1
            n_slab += 1
            continue
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
print(f"{n} random configurations ({n_slab} of them slab runs with the fused halo exchange), {bad} failures (seed {seed})")
mp.undo()
