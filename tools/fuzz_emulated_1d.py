"""Fuzz of the fused 1D path on the CPU emulator (test infrastructure, tests/emu): random programs, line lengths, pass
depths, step splits, random initial state, with and without the running DFT, bit-for-bit against the numpy oracle.
    python tools/fuzz_emulated_1d.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fdtd_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402
from tests.emu import device  # noqa: E402
from tests.test_gpu_fd1d import _sim_for  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
mp = pytest.MonkeyPatch()
emu = device.install(mp)
t_end, n, bad = time.time() + budget, 0, 0
while time.time() < t_end:
    prog = str(rng.choice(["1_1", "1_2", "1_3", "1_4", "1_5", "2_1", "2_2", "2_3"]))
    nx = int(rng.integers(12, 3000 if rng.random() < 0.5 else 200))
    tblock = int(rng.integers(1, 65))
    ns = int(rng.integers(1, 80))
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    p, src = cases.line_program(prog, nx, ns, dtype)
    dft = p.freqs is not None and rng.random() < 0.7
    if not dft:
        p.freqs = None
    elif rng.random() < 0.3:
        p.freqs = p.freqs[:int(rng.integers(1, 3))]
        p.__post_init__()
    cfg = dict(prog=prog, nx=nx, tblock=tblock, ns=ns, dtype=np.dtype(dtype).name, dft=dft)
    try:
        sim = _sim_for(prog, nx, dtype, tblock=tblock, device="cpu", **({"freqs": p.freqs} if dft else {}))
        names = ["ex", "hy"] + (["dx", "ix"] if p.form == "flux" else []) + (["sx"] if sim.debye else []) + (["bc"] if p.abc else [])
        acc = ["r_pt", "i_pt", "r_in", "i_in"] if dft else []
        if rng.random() < 0.7:
            for name in names + acc:
                a = getattr(p, name)
                a[...] = rng.uniform(-1, 1, a.shape).astype(dtype)
                (getattr(sim.ft, name) if name in acc else sim.tensor(name)).copy_(torch.from_numpy(a))
        cut = int(rng.integers(0, ns + 1))
        for part in (cut, ns - cut):
            sim.advance(part)
        orc.advance_1d(p, src)
        for name in names + acc:
            assert sim.get(name).tobytes() == np.ascontiguousarray(getattr(p, name)).tobytes(), name
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", cfg, type(e).__name__, str(e)[:300], flush=True)
    n += 1
print(f"{n} random 1D configurations, {bad} failures (seed {seed})")
mp.undo()
