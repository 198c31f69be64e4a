"""Worker for tests/test_slab_gloo.py: run under torch.distributed.run with the gloo backend (CPU).

Exercises simulation_b200.slab (partition, ghost sizing, grouped send/recv, block scheduling) with a numpy
stand-in for the per-rank CUDA stepper: the ORACLE arithmetic applied to the rank's stored rows (owned +
ghost) as a sub-grid.  The stitched result must equal the monolithic oracle run bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fdtd_oracle as orc          # noqa: E402  (test infrastructure)
from simulation_b200 import slab               # noqa: E402


class NumpyStandIn:
    """Same surface as fd2d.Fdtd2D for what slab.py touches; numpy/oracle arithmetic on the stored rows."""

    def __init__(self, nx, ny, npml, dtype, rows, ghost, tblock=4, source=None, naz=None):
        self.nx, self.ny = nx, ny
        self.row_lo, self.row_hi = rows
        self.row_base = max(self.row_lo - ghost, 0)
        self.rows_alloc = min(self.row_hi + ghost, nx) - self.row_base
        self.lossy, self.t, self.source = False, 0, source
        pml = orc.pml_vectors(nx, ny, npml, dtype)
        sl = slice(self.row_base, self.row_base + self.rows_alloc)
        pml = {k: (v[sl].copy() if k[1] == "x" else v) for k, v in pml.items()}
        point = None
        if source is not None:
            si, sj = source["i"], source["j"]
            if self.row_base <= si < self.row_base + self.rows_alloc:
                point = (si - self.row_base, sj)
        self.g = orc.Grid2D(self.rows_alloc, ny, npml, dtype, point=point, pml=pml,
                            naz=None if naz is None else naz[sl].copy())
        self.table = source["table"] if source is not None else None

    def tensor(self, name, stored=False):
        t = torch.from_numpy(getattr(self.g, name))
        if stored:
            return t
        o = self.row_lo - self.row_base
        return t[o:o + self.row_hi - self.row_lo]

    def advance(self, n, tblock=None, lazy_ez=False):
        for k in range(n):
            t = self.t + 1
            orc.step_2d(self.g, np.int32(t), self.table[t - 1] if self.table is not None else 0.0)
            self.t = t

    def synchronize(self):
        pass


def main():
    nx, ny, npml, ns, ghost, dtype = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                                      int(sys.argv[5]), np.dtype(sys.argv[6]).type)
    engine = sys.argv[7] if len(sys.argv) > 7 else "numpy"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    table = orc.source_table("sine", ns, freq=1500e6)
    rng = np.random.default_rng(3)
    naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(dtype)
    src = {"i": nx // 2 - 5, "j": ny // 2 - 5, "table": table}
    if engine == "emu":
        # the PRODUCT's per-rank stepper (fd2d.Fdtd2D host code + the kernels' own source on the CTA emulator of
        # tests/emu) under slab.py's grouped send/recv exchange, CPU tensors as device memory
        import pytest
        from tests.emu import device
        device.install(pytest.MonkeyPatch())
        from simulation_b200 import fd2d, surface
        s = slab.SlabFdtd2D(nx, ny, npml, dtype, ghost=ghost, tblock=min(ghost, 6), halo="nccl", device="cpu", naz=naz,
                            source=fd2d.PointSource(src["i"], src["j"], surface.Sinusoid(1500e6)))
        assert s.halo_mode == "nccl" and type(s.engine) is fd2d.Fdtd2D
    else:
        s = slab.SlabFdtd2D(nx, ny, npml, dtype, ghost=ghost, engine_factory=NumpyStandIn, source=src, naz=naz)
    lo, hi = slab.partition(nx, world, rank)
    assert (s.row_lo, s.row_hi) == (lo, hi)
    # uneven advance calls: 7 steps, then the rest
    s.advance(min(7, ns))
    s.advance(ns - min(7, ns))
    mono = orc.Grid2D(nx, ny, npml, dtype, point=(src["i"], src["j"]), naz=naz.copy())
    orc.advance_2d(mono, table)
    ok = True
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        whole = s.gather(name)
        if rank == 0 and whole.tobytes() != getattr(mono, name).tobytes():
            print(f"MISMATCH {name}", np.argwhere(whole != getattr(mono, name))[:5].tolist(), flush=True)
            ok = False
    expected_blocks = -(-min(7, ns) // ghost) + -(-(ns - min(7, ns)) // ghost)
    assert s.exchanges == expected_blocks, (s.exchanges, expected_blocks)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
