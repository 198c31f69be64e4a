"""Second fuzz of the 2D path on the CPU emulator (test infrastructure, tests/emu): the running DFT carried with the rows,
row slabs with ghost rows refreshed between blocks (what slab.py does), lazy Ez, uploaded random state and media --
bit-for-bit against the numpy oracle.   python tools/fuzz_emulated_2d_modes.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fdtd_oracle as orc  # noqa: E402
from simulation_b200 import fd2d, surface  # noqa: E402
from tests import cases  # noqa: E402
from tests.emu import device  # noqa: E402
from tests.test_gpu_fd2d import _sim_for  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
mp = pytest.MonkeyPatch()
emu = device.install(mp)
t_end, n, bad = time.time() + budget, 0, 0
STATE = ("dz", "hx", "hy", "ihx", "ihy")
while time.time() < t_end:
    mode = str(rng.choice(sys.argv[3].split(",") if len(sys.argv) > 3 else ["dft", "slab", "random_state", "p2p", "streamed"]))
    dtype = np.float32 if rng.random() < 0.7 else np.float64
    nx, ny = int(rng.integers(40, 260)), int(rng.integers(40, 420))
    npml = int(rng.integers(2, min(nx, ny) // 2 - 8))
    cfg = dict(mode=mode, nx=nx, ny=ny, npml=npml, dtype=np.dtype(dtype).name)
    try:
        emu.fdtd2d_tune(int(rng.choice([0, 1, 2, 4])), int(rng.choice([0, 5, 16, 40])), 0, 0, 0)
        if mode == "dft":
            ns, tb = int(rng.integers(1, 20)), int(rng.choice([0, 1, 2, 3, 4, 6]))
            radius = float(rng.uniform(0.05, 0.3))
            cfg.update(ns=ns, tblock=tb)
            g, src = cases.grid_program("3_4", nx, ny, ns, dtype, npml=npml, radius=radius, dft=True)
            sim = fd2d.Fdtd2D(nx, ny, npml, dtype, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)),
                              naz=g.naz.copy(), nbz=g.nbz.copy(), freqs=g.freqs, device="cpu")
            if rng.random() < 0.7:                       # signal at every strip / chunk boundary from the first step on
                import torch
                for name in STATE + ("iz",):
                    a = rng.standard_normal((nx, ny)).astype(dtype)
                    if name == "iz":
                        a = np.where(g.nbz != 0, a, 0).astype(dtype)      # iz lives where the medium is lossy
                    sim.set(name, a)
                    getattr(g, name)[...] = a
                for name in ("r_pt", "i_pt", "r_in", "i_in"):
                    a = getattr(g, name)
                    a[...] = rng.standard_normal(a.shape).astype(dtype)
                    getattr(sim.ft, name).copy_(torch.from_numpy(a.reshape(tuple(getattr(sim.ft, name).shape))))
            cut = int(rng.integers(0, ns + 1))
            sim.advance(cut, tblock=tb or None)
            sim.advance(ns - cut, tblock=tb or None)
            orc.advance_2d(g, src)
            names = ("dz", "ez", "iz", "hx", "hy", "ihx", "ihy", "r_pt", "i_pt", "r_in", "i_in")
            for name in names:
                assert sim.get(name).tobytes() == getattr(g, name).tobytes(), name
        elif mode == "slab":
            prog = str(rng.choice(["3_2", "3_3", "3_4"]))
            T = int(rng.choice([1, 2, 3, 4, 6, 8]))
            nslab = int(rng.integers(2, 5))
            if nx // nslab < T + 2:
                continue
            nblocks = int(rng.integers(1, 4))
            cfg.update(prog=prog, T=T, nslab=nslab, nblocks=nblocks)
            cuts = np.linspace(0, nx, nslab + 1).astype(int)
            slabs = [_sim_for(prog, nx, ny, dtype, npml=npml, radius=0.15, rows=(int(lo), int(hi)), ghost=T, tblock=T, device="cpu")
                     for lo, hi in zip(cuts[:-1], cuts[1:])]
            names = STATE + (("iz",) if prog == "3_4" else ())
            init = {}
            if rng.random() < 0.7:                       # non-zero state everywhere: the slab boundaries carry signal at once
                g0, _ = cases.grid_program(prog, nx, ny, 1, dtype, npml=npml, radius=0.15, dft=False)
                for k in names:
                    a = rng.standard_normal((nx, ny)).astype(dtype)
                    init[k] = np.where(g0.nbz != 0, a, 0).astype(dtype) if k == "iz" else a
                    for s in slabs:
                        s.set(k, init[k])                # whole-grid array: owned and ghost rows
            for _ in range(nblocks):
                for s in slabs:
                    s.advance(T, lazy_ez=True)
                whole = {k: np.concatenate([s.get(k) for s in slabs]) for k in names}
                for s in slabs:
                    for k in names:
                        s.set(k, whole[k])
            for s in slabs:
                s.advance(1)
            ns = nblocks * T + 1
            g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml, radius=0.15, dft=False)
            for k, a in init.items():
                getattr(g, k)[...] = a
            orc.advance_2d(g, src)
            for k in names + ("ez",):
                assert np.concatenate([s.get(k) for s in slabs]).tobytes() == getattr(g, k).tobytes(), k
        elif mode == "p2p":
            # the halo exchange fused into the pass: every "GPU" a slab in this process, peers wired by raw pointers
            prog = str(rng.choice(["3_2", "3_3", "3_4"]))
            T = int(rng.choice([1, 2, 3, 4, 6, 8]))
            nslab = int(rng.integers(2, 5))
            if nx // nslab < 2 * T + 2:
                continue
            nblocks = int(rng.integers(1, 5))
            cfg.update(prog=prog, T=T, nslab=nslab, nblocks=nblocks)
            cuts = np.linspace(0, nx, nslab + 1).astype(int)
            slabs = [_sim_for(prog, nx, ny, dtype, npml=npml, radius=0.15, rows=(int(lo), int(hi)), ghost=T, tblock=T, device="cpu")
                     for lo, hi in zip(cuts[:-1], cuts[1:])]
            names = [k for k in fd2d.FIELD_NAMES if k != "ez" and (k != "iz" or prog == "3_4")]
            words = [np.zeros(4, dtype=np.int64) for _ in slabs]
            for r, s in enumerate(slabs):
                peer = lambda q: None if q is None else {"row_base": slabs[q].row_base, "sync": words[q].ctypes.data,
                                                         "sets": [{k: slabs[q]._sets[i][k].data_ptr() for k in names} for i in range(2)]}
                s.p2p = {"halo": T, "sync": type("W", (), {"ptr": words[r].ctypes.data})(),
                         "up": peer(r - 1 if r > 0 else None), "dn": peer(r + 1 if r < nslab - 1 else None)}
            init = {}
            if rng.random() < 0.7:
                g0, _ = cases.grid_program(prog, nx, ny, 1, dtype, npml=npml, radius=0.15, dft=False)
                for k in names:
                    a = rng.standard_normal((nx, ny)).astype(dtype)
                    init[k] = np.where(g0.nbz != 0, a, 0).astype(dtype) if k == "iz" else a
                    for s in slabs:
                        s.set(k, init[k])
            for epoch in range(1, nblocks + 1):
                for r in rng.permutation(nslab):
                    slabs[int(r)].advance(T, tblock=T, lazy_ez=epoch < nblocks, epoch=epoch)
            ns = nblocks * T
            g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml, radius=0.15, dft=False)
            for k, a in init.items():
                getattr(g, k)[...] = a
            orc.advance_2d(g, src)
            for k in names + ["ez"]:
                assert np.concatenate([s.get(k) for s in slabs]).tobytes() == getattr(g, k).tobytes(), k
        elif mode == "streamed":
            import torch
            ns = int(rng.integers(1, 30))
            schedule = str(rng.choice(["skewed", "wavefront"]))
            plan = [int(x) for x in rng.integers(8, 120, size=int(rng.integers(1, 12)))]
            cfg.update(ns=ns, schedule=schedule, plan=plan)
            naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(np.float32)
            src = fd2d.PointSource(nx // 2 - 5, ny // 2 - 5, surface.Sinusoid(1500e6))
            a = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, naz=naz, device="cpu")
            a.advance(ns)
            b = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, device="cpu")
            host_ez = torch.empty((nx, ny), dtype=torch.float32)
            b.run_streamed(ns, torch.from_numpy(naz), host_ez, block_rows=plan, streams=int(rng.integers(1, 9)), schedule=schedule,
                           window=[None, 1, 2, 3][int(rng.integers(0, 4))])
            assert torch.equal(host_ez, a.tensor("ez"))
            for k in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
                assert torch.equal(a.tensor(k), b.tensor(k)), k
        else:
            prog = str(rng.choice(["3_2", "3_3"]))
            ns, tb = int(rng.integers(1, 20)), int(rng.choice([0, 1, 3, 4, 6, 8]))
            cfg.update(prog=prog, ns=ns, tblock=tb)
            naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(dtype)
            sim = _sim_for(prog, nx, ny, dtype, npml=npml, naz=naz, device="cpu")
            g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml, naz=naz.copy())
            for name in STATE:
                a = rng.standard_normal((nx, ny)).astype(dtype)
                sim.set(name, a)
                getattr(g, name)[...] = a
            sim.advance(ns, tblock=tb or None)
            orc.advance_2d(g, src)
            for name in STATE + ("ez",):
                assert sim.get(name).tobytes() == getattr(g, name).tobytes(), name
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", cfg, type(e).__name__, str(e)[:300], flush=True)
    finally:
        emu.fdtd2d_tune(0, 0, 0, 0, 0)
    n += 1
print(f"{n} random 2D mode configurations, {bad} failures (seed {seed})")
mp.undo()
