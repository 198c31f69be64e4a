"""Row-slab decomposition of one 2D TM grid over the GPUs of a node: one process per GPU
(``torch.distributed``), rank r owns a contiguous band of rows and keeps ``ghost`` extra rows of its
neighbours on each side.

The reference has no multi-GPU path; this is the slab protocol of the north star built on
the fused kernel: a time block of ``ghost`` steps needs NO communication (the kernel recomputes the
shrinking ghost band), then each rank refreshes its ghost rows from its neighbours' freshly computed owned
rows -- one grouped NCCL send/recv per neighbour per block, ``5 * ghost`` rows each way (dz, hx, hy, ihx,
ihy; plus iz for a lossy medium; ez is recomputed, naz / nbz are static).  The stencil reaches one row per
step (hy[i-1] in dfield, ez[i+1] in hfield), so ``ghost`` rows stay valid for ``ghost`` steps.

Result: every rank's owned rows are bit-identical to the same rows of a single-device run.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

EXCHANGED = ("dz", "hx", "hy", "ihx", "ihy", "iz")


def partition(nx: int, world: int, rank: int) -> tuple:
    """Rows [lo, hi) of rank ``rank``: contiguous, sizes differ by at most one row."""
    base, extra = divmod(nx, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class SlabFdtd2D:
    """One rank's slab.  Same ``advance`` / ``tensor`` / ``get`` surface as :class:`fd2d.Fdtd2D`.

    ``engine_factory(nx, ny, npml, dtype, rows=(lo, hi), ghost=g, **kw)`` builds the per-rank stepper;
    the default is the CUDA :class:`fd2d.Fdtd2D` (tests inject a CPU stand-in to exercise the exchange
    logic under gloo without a GPU)."""

    def __init__(self, nx: int, ny: int, npml: int = 0, dtype=np.float32, *, ghost: Optional[int] = None,
                 tblock: int = 4, group=None, engine_factory: Optional[Callable] = None, halo: Optional[str] = None,
                 **engine_kw):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nx, self.ny = int(nx), int(ny)
        self.row_lo, self.row_hi = partition(self.nx, self.world, self.rank)
        if ghost is None:
            if tblock:
                ghost = int(tblock)
            elif engine_factory is None:
                # the library chooses the pass depth by slab size: ask it (the thinnest slab decides for everybody)
                from .fd2d import plan_depths
                thin = partition(self.nx, self.world, self.world - 1)
                freqs = engine_kw.get("freqs")
                # (the deepest pass any split may start with: a long run's first pass, or the single pass of a 12-step run)
                ghost = max(plan_depths(self.nx, self.ny, dtype, n, 0, rows=thin, lossy=engine_kw.get("nbz") is not None,
                                        nf=0 if freqs is None else min(len(freqs), 3))[0] for n in (1 << 10, 12))
            else:
                ghost = 4
        self.ghost = int(ghost)
        if self.world > 1 and self.ghost < 1:
            raise ValueError(f"ghost band of {self.ghost} rows: at least one ghost row is needed between exchanges")
        if self.world > 1 and (self.row_hi - self.row_lo) < self.ghost:
            raise ValueError(f"slab of {self.row_hi - self.row_lo} rows is thinner than the ghost band {self.ghost}")
        want = halo or os.environ.get("FDTD_SLAB_HALO", "p2p")          # "p2p" (fused into the pass) or "nccl"
        p2p_wanted = self.world > 1 and want == "p2p" and engine_factory is None and engine_kw.get("freqs") is None
        if engine_factory is None:
            from .fd2d import Fdtd2D
            engine_factory = Fdtd2D
        if p2p_wanted:
            engine_kw = dict(engine_kw, ipc=True)
        self.engine = engine_factory(self.nx, self.ny, npml, dtype, rows=(self.row_lo, self.row_hi),
                                     ghost=self.ghost if self.world > 1 else 0, tblock=tblock, **engine_kw)
        self.row_base = self.engine.row_base
        self.up = self.rank - 1 if self.rank > 0 else None            # neighbour owning the rows above
        self.down = self.rank + 1 if self.rank < self.world - 1 else None
        self.exchanges = 0
        self._ghost_dirty = False
        self._epoch = 0
        self.halo_mode = "nccl"
        if p2p_wanted:
            try:
                self._enable_p2p()
            except Exception as e:                      # no peer access / IPC: the grouped NCCL exchange still works
                self.halo_mode = f"nccl (p2p unavailable: {type(e).__name__}: {e})"
            ok = torch.tensor([1 if self.halo_mode == "p2p" else 0], device=self.engine.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0 and self.halo_mode == "p2p":
                self.halo_mode = "nccl (a peer could not map this rank)"
                self.engine.p2p = None

    # ---- delegation ---------------------------------------------------------------------------------
    @property
    def t(self):
        return self.engine.t

    @t.setter
    def t(self, v):
        self.engine.t = v

    @property
    def naz(self):
        return self.engine.naz

    def tensor(self, name, stored=False):
        return self.engine.tensor(name, stored=stored)

    def get(self, name):
        return self.engine.get(name)

    def set(self, name, host):
        self.engine.set(name, host)
        self._ghost_dirty = True          # ghost rows are refreshed before the next block

    def synchronize(self):
        self.engine.synchronize()
        self.check_halo()

    def check_halo(self) -> None:
        """Raise if a pass of the fused halo exchange gave up waiting for a neighbour (a rank that died or skipped a
        call): the library bounds that wait and records the failure instead of hanging (``fdtd2d_halo_status``)."""
        eng = self.engine
        if eng is None or getattr(eng, "p2p", None) is None:
            return
        from . import _lib
        import ctypes as C
        prob = _lib.Problem2D()
        prob.sync_local = eng.p2p["sync"].ptr
        word = C.c_ulonglong(0)
        with torch.cuda.device(eng.device):
            _lib.check(_lib.lib().fdtd2d_halo_status(C.byref(prob), C.byref(word)), "fdtd2d_halo_status")
        if word.value:
            raise _lib.FdtdError("fused halo exchange failed: " + _lib.lib().fdtd_last_error().decode(errors="replace"))

    # ---- checkpoint / restore ------------------------------------------------------------------------
    def checkpoint(self) -> dict:
        """This rank's part of a checkpoint (owned rows, see :meth:`fd2d.Fdtd2D.checkpoint`); collective only in
        that every rank should take it after the same ``advance`` calls."""
        return self.engine.checkpoint()

    def restore(self, ckpt: dict) -> None:
        """Inverse of :meth:`checkpoint` on every rank.  The engine restores the owned rows only, so the ghost rows
        are marked stale and refreshed from the neighbours before the next block of steps."""
        self.engine.restore(ckpt)
        self._ghost_dirty = self.world > 1

    # ---- fused halo exchange over peer memory ------------------------------------------------------------
    def _enable_p2p(self) -> None:
        """Map the neighbours' state arrays and sync words into this process (CUDA IPC, opened with THIS rank's
        device current) so that the pass itself stores its edge rows into their ghost rows and flags completion."""
        from . import _lib
        eng = self.engine
        names = self._names()
        with torch.cuda.device(eng.device):
            sync = _lib.DeviceBuffer((8,), np.int64)          # {from_up, from_down, counter, counter, error, reserved}
            mine = {"row_base": eng.row_base, "sync": sync.ipc_handle(),
                    "sets": [{n: eng._ipc_handles[s][n] for n in names} for s in range(2)]}
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=self.group)

            def open_peer(r):
                if r is None:
                    return None
                d = everyone[r]
                return {"row_base": d["row_base"], "sync": _lib.ipc_open(d["sync"]),
                        "sets": [{n: _lib.ipc_open(d["sets"][s][n]) for n in names} for s in range(2)]}

            eng.p2p = {"halo": self.ghost, "sync": sync, "up": open_peer(self.up), "dn": open_peer(self.down)}
        self.halo_mode = "p2p"            # the caller's all_reduce doubles as the barrier after the mapping

    def close(self) -> None:
        """Collective teardown: unmap the neighbours' arrays before anybody frees them (CUDA IPC rule), then drop the
        engine.  Needed only when another problem is built afterwards in the same process."""
        from . import _lib
        eng = self.engine
        if eng is None:
            return
        torch.cuda.synchronize(eng.device)
        p2p = getattr(eng, "p2p", None)
        if self.world > 1:
            dist.barrier(group=self.group)
        if p2p is not None:
            with torch.cuda.device(eng.device):
                for side in ("up", "dn"):
                    nb = p2p[side]
                    if nb is None:
                        continue
                    _lib.ipc_close(nb["sync"])
                    for st in nb["sets"]:
                        for ptr in st.values():
                            _lib.ipc_close(ptr)
            eng.p2p = None
            if self.world > 1:
                dist.barrier(group=self.group)
        self.engine = None

    # ---- ghost exchange ------------------------------------------------------------------------------
    def _names(self):
        return [n for n in EXCHANGED if n != "iz" or getattr(self.engine, "lossy", False)]

    def exchange_ghosts(self) -> None:
        """Refresh both ghost bands from the neighbours' owned rows (all fields, one grouped batch)."""
        if self.world == 1:
            return
        g, ops, keep = self.ghost, [], []
        o = self.row_lo - self.row_base                     # array row of the first owned row
        n_own = self.row_hi - self.row_lo
        for name in self._names():
            t = self.engine.tensor(name, stored=True)
            if self.up is not None:
                ops.append(dist.P2POp(dist.isend, t[o:o + g], self.up, self.group))
                ops.append(dist.P2POp(dist.irecv, t[o - g:o], self.up, self.group))
            if self.down is not None:
                ops.append(dist.P2POp(dist.isend, t[o + n_own - g:o + n_own], self.down, self.group))
                ops.append(dist.P2POp(dist.irecv, t[o + n_own:o + n_own + g], self.down, self.group))
            keep.append(t)
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        self.exchanges += 1

    def advance(self, nsteps: int, tblock: Optional[int] = None) -> None:
        """``nsteps`` steps: blocks of at most ``ghost`` steps, a ghost exchange after every block."""
        left = int(nsteps)
        if self._ghost_dirty:
            if self.halo_mode == "p2p":
                # NCCL refresh of uploaded state: every rank must be idle on both sides of it
                torch.cuda.synchronize(self.engine.device)
                dist.barrier(group=self.group)
            self.exchange_ghosts()
            self.exchanges -= 1
            if self.halo_mode == "p2p":
                torch.cuda.synchronize(self.engine.device)
                dist.barrier(group=self.group)
            self._ghost_dirty = False
        while left > 0:
            if self.world == 1:
                self.engine.advance(left, tblock=tblock)
                return
            # The library splits a run into passes of least total cost (e.g. 20 steps = 8 + 6 + 6, not 8 + 8 + 4): the
            # blocks between exchanges follow THAT split of everything still to come, not a split of ghost-sized blocks.
            depths = self._depths_left(left, tblock)
            if self.halo_mode == "p2p":
                # the pass pushes its edge rows into the neighbours' ghost rows and waits on their flags itself.
                # ONE pass per call: the push lands in the set the neighbour's earlier passes of a multi-pass call
                # would still be reading (the per-call handshake only orders whole calls)
                n = depths[0]
                left -= n
                self._epoch += 1
                self.engine.advance(n, tblock=tblock, lazy_ez=left > 0, epoch=self._epoch)
                self.exchanges += 1
            else:
                n = 0
                for d in depths:                   # as many whole passes as the ghost band covers
                    if n + d > self.ghost:
                        break
                    n += d
                left -= n
                self.engine.advance(n, tblock=tblock, lazy_ez=left > 0)    # ez is stored by the last block only
                self.exchange_ghosts()

    def _depths_left(self, left: int, tblock) -> list:
        """Pass depths of the next steps: the engine's split of all ``left`` steps when its first pass fits the ghost
        band, else its split of one ghost-sized block."""
        if not hasattr(self.engine, "pass_depths"):         # a stand-in stepper (tests): ghost-sized blocks
            return [min(left, self.ghost)]
        depths = self.engine.pass_depths(left, tblock)
        if depths[0] > self.ghost:
            depths = self.engine.pass_depths(min(left, self.ghost), tblock)
        return depths

    def run_streamed(self, nsteps: int, naz_host: torch.Tensor, ez_host: torch.Tensor, **kw) -> None:
        """Host medium in, ``nsteps <= ghost`` steps, host Ez out, with the transfers hidden behind the kernels and NO
        exchange at all: every rank runs the block wavefront of :meth:`fd2d.Fdtd2D.run_streamed` on its stored rows
        and lets the ghost band decay (one row per step).  ``naz_host`` covers the stored rows, ``ez_host`` the owned
        rows.  The ghost rows are refreshed (one exchange) before the next ``advance``."""
        self.engine.run_streamed(nsteps, naz_host, ez_host, **kw)
        self._ghost_dirty = self.world > 1

    def gather(self, name: str) -> Optional[np.ndarray]:
        """Whole-grid field on rank 0 (tests / small grids only)."""
        mine = self.engine.tensor(name).contiguous()
        if self.world == 1:
            return mine.cpu().numpy()
        parts = [None] * self.world
        dist.all_gather_object(parts, mine.cpu().numpy(), group=self.group)
        return np.concatenate(parts, axis=0) if self.rank == 0 else None
