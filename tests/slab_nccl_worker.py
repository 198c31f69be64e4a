"""Worker for tests/test_gpu_slab.py: run under torch.distributed.run, one rank per GPU (NCCL).
N-rank slab run of the fused CUDA path vs the single-device run of the same problem: bitwise equal.

    slab_nccl_worker.py prog nx ny npml ns tblock [ghost [mode [v4]]]
mode: plain (default) | streamed (first block of steps through run_streamed, no exchange) | skip (rank 1 skips a call:
the neighbours' bounded wait must report it).  v4: force 4-wide vectors and 128-row chunks -- the kernels and launch plan
of the bench on a small grid."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from simulation_b200 import _lib, fd2d, slab, surface   # noqa: E402


def make_source(prog, nx, ny):
    if prog == "3_3":
        return fd2d.IncidentWave(surface.Gaussian(20, 8.0))
    return fd2d.PointSource(nx // 2 - 5, ny // 2 - 5, surface.Sinusoid(1500e6), hard=True)


def skip_case(s, rank):
    """Rank 1 skips the second call.  Rank 0's third call waits for rank 1's epoch-2 flag, which never comes: the pass
    gives up after the bound, raises the error word, and synchronize() turns it into an exception."""
    _lib.lib().fdtd2d_tune2(_lib.TUNE_HALO_WAIT_MS, 300)
    s.advance(6)
    ok = True
    if rank == 1:
        s._epoch += 1                       # (keeps the epoch counters aligned with the ranks that did run)
    else:
        s.advance(6)
    if rank == 0:
        s.advance(6)
        try:
            s.synchronize()
            print("rank 0: the missing neighbour went unnoticed", flush=True)
            ok = False
        except _lib.FdtdError as e:
            print(f"rank 0 reports: {e}", flush=True)
            ok = "gave up waiting" in str(e)
    torch.cuda.synchronize()
    return ok


def main():
    prog, nx, ny, npml, ns, tblock = sys.argv[1], *[int(x) for x in sys.argv[2:7]]
    ghost = int(sys.argv[7]) if len(sys.argv) > 7 else None          # ghost rows (default: tblock); > tblock: several passes per exchange
    mode = sys.argv[8] if len(sys.argv) > 8 else "plain"
    v4 = len(sys.argv) > 9 and sys.argv[9] == "v4"
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank = dist.get_rank()
    if v4:
        _lib.lib().fdtd2d_tune(4, 128, 0, 0, 0)
    rng = np.random.default_rng(11)
    naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(np.float32)
    streamed = mode == "streamed"
    s = slab.SlabFdtd2D(nx, ny, npml, np.float32, tblock=tblock, ghost=ghost, source=make_source(prog, nx, ny),
                        naz=None if streamed else naz)
    if mode == "skip":
        ok = skip_case(s, rank)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag.item()) == 1 else 1)
    if streamed:
        e = s.engine
        host_naz = torch.from_numpy(naz[e.row_base:e.row_base + e.rows_alloc].copy()).pin_memory()
        host_ez = torch.empty((s.row_hi - s.row_lo, ny), dtype=torch.float32).pin_memory()
        first = min(ns - 7, s.ghost)
        s.run_streamed(first, host_naz, host_ez, blocks=3, streams=4)       # consumes the ghost band, no exchange
        s.synchronize()
        mine = s.tensor("ez").cpu()
        assert torch.equal(host_ez, mine), "streamed Ez on the host differs from the device copy"
        s.advance(ns - first)          # ghost rows are refreshed first
    else:
        s.advance(7)                   # ragged split of the step count across advance() calls
        s.advance(ns - 7)
    s.synchronize()                    # (also checks the fused exchange's error word)
    ok = True
    fields = {name: s.gather(name) for name in ("dz", "ez", "hx", "hy", "ihx", "ihy")}
    if rank == 0:
        one = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=make_source(prog, nx, ny), naz=naz, tblock=tblock)
        one.advance(ns)
        for name, whole in fields.items():
            ref = one.get(name)
            if whole.tobytes() != ref.tobytes():
                bad = np.argwhere(whole != ref)
                print(f"MISMATCH {name}: {len(bad)} cells, first {bad[:4].tolist()}", flush=True)
                ok = False
        assert np.abs(fields["ez"]).max() > 1e-3
        print(f"slab x{dist.get_world_size()} {prog} {nx}x{ny} ns={ns} T={tblock}{' v4' if v4 else ''}: {'OK' if ok else 'FAIL'}, "
              f"{s.exchanges} exchanges, halo mode: {s.halo_mode}", flush=True)
        want = os.environ.get("FDTD_SLAB_HALO", "p2p")
        if want == "p2p" and s.halo_mode != "p2p":
            print("P2P halo exchange was requested but is not active", flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
