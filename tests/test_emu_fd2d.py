"""CPU check of the fused 2D kernels' OWN SOURCE and of the product's HOST code together: the whole library
(simulation_b200/csrc/fd1d.cu, fd2d_steps.cu, fd2d_march.cu) compiled for the host against the fiber-based CTA
emulator of tests/emu/, and simulation_b200.fd2d.Fdtd2D pointed at it with CPU tensors as device memory
(tests/emu/device.py).  Bit-for-bit against the numpy oracle: strips / chunks / halos of the march kernel, the
interior-vs-careful split and the special-strip lists, PML / TFSF / source cells, the incident-line history, the
running DFT carried with the rows, row slabs with ghost rows, lazy Ez, checkpoint / restore.

Test infrastructure only -- the product path is the sm_100a build and has no CPU fallback; tests/test_gpu_fd2d.py
runs these and many more cases on the device.  The emulator completes a cp.async only when a wait_group covering it executes, so the
ring discipline is checked too; what it cannot see is timing and memory-ordering races between warps (CTAs of one launch
interleave only at barriers)."""
import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from tests import cases
from tests.emu import device
from tests.test_gpu_fd2d import _assert_same, _sim_for

torch = pytest.importorskip("torch")


@pytest.fixture
def emu(monkeypatch):
    return device.install(monkeypatch)


def _run(emu, prog, nx, ny, npml, ns, dtype, tblock=None, radius=0.12, tune=(0, 0, 0, 0, 0), parts=None, deep=1, edge=1,
         variant=0, fast=3):
    before = emu.emu_launches()
    emu.fdtd2d_tune(*tune)
    emu.fdtd2d_tune2(0, deep)
    emu.fdtd2d_tune2(3, edge)
    emu.fdtd2d_tune2(2, variant)
    emu.fdtd2d_tune2(4, fast)
    try:
        sim = _sim_for(prog, nx, ny, dtype, npml=npml, radius=radius, device="cpu")
        for part in (parts or (ns,)):
            sim.advance(part, tblock=tblock)
    finally:
        emu.fdtd2d_tune(0, 0, 0, 0, 0)
        emu.fdtd2d_tune2(0, 1)
        emu.fdtd2d_tune2(3, 1)
        emu.fdtd2d_tune2(2, 0)
        emu.fdtd2d_tune2(4, 3)
    assert sim.t == ns and emu.emu_launches() > before
    g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml, radius=radius, dft=False)
    orc.advance_2d(g, src)
    _assert_same(sim, g, prog)
    return sim


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tblock", [1, 3, 4, 6, 8])
@pytest.mark.parametrize("prog,nx,ny,npml,ns", [("3_2", 56, 72, 8, 41), ("3_3", 64, 48, 7, 45), ("3_4", 60, 72, 8, 37),
                                                ("3_1", 40, 56, 0, 30)])
def test_emulated_advance_matches_oracle(emu, prog, nx, ny, npml, ns, tblock, dtype):
    _run(emu, prog, nx, ny, npml, ns, dtype, tblock)


@pytest.mark.parametrize("force_v,chunk_rows,tblock", [(4, 0, 6), (2, 16, 6), (1, 5, 4), (4, 40, 8), (2, 0, 3)])
@pytest.mark.parametrize("prog,nx,ny,npml", [("3_2", 300, 1100, 8), ("3_3", 280, 1000, 12), ("3_4", 260, 1040, 10)])
def test_emulated_interior_and_careful_kernels(emu, prog, nx, ny, npml, force_v, chunk_rows, tblock):
    """Grids with a true interior: the mask-free identity-coefficient kernel (packed arithmetic at fp32) and the
    careful kernel share every pass; every vector width, several row partitions, ragged last strips and chunks."""
    before = emu.emu_launches()
    _run(emu, prog, nx, ny, npml, 2 * tblock + 1, np.float32, tblock, radius=0.3, tune=(force_v, chunk_rows, 0, 0, 0))
    # 3 passes, each = careful kernel + interior kernel (+ the incident-line kernel with a TFSF source); the oracle-side
    # setup launches nothing, the device-side identity check two kernels
    # 3_4 (lossy cylinder in free space): two interior kernels per pass -- lossy for the warps that meet the cylinder's
    # box, lossless for the others -- and one more setup check (the lossless-outside promise)
    per_pass = {"3_2": 2, "3_3": 3, "3_4": 4}[prog]
    # the two depth-8 passes at 4-wide vectors are deep passes: without TFSF / loss the column and the row variant of the
    # warp-chain kernel launch beside the interior one
    extra = 2 * 2 if (prog == "3_2" and force_v == 4 and tblock == 8) else 0
    assert emu.emu_launches() - before == 3 * per_pass + extra + (3 if prog == "3_4" else 2), "an interior kernel did not run in every pass"


def test_emulated_interior_kernel_equals_careful_kernel(emu):
    nx, ny, npml, ns = 330, 900, 10, 19
    a = _run(emu, "3_3", nx, ny, npml, ns, np.float32)
    b = _run(emu, "3_3", nx, ny, npml, ns, np.float32, tune=(0, 0, 0, 0, 1))      # every warp through the careful kernel
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        assert a.get(name).tobytes() == b.get(name).tobytes(), name


@pytest.mark.parametrize("ny", [61, 62, 63, 130])
def test_emulated_odd_widths(emu, ny):
    _run(emu, "3_3", 50, ny, 6, 29, np.float32)


def test_emulated_split_calls_mixed_with_reference_named_steps(emu):
    nx, ny, npml = 70, 90, 8
    sim = _sim_for("3_3", nx, ny, np.float64, npml=npml, device="cpu")
    sim.advance(23)
    sim.step()                       # ezinct, dfield, inctdz, efield, hxinct, hfield, incthx, incthy: one kernel each
    sim.advance(30, tblock=3)
    g, src = cases.grid_program("3_3", nx, ny, 54, np.float64, npml=npml)
    orc.advance_2d(g, src)
    _assert_same(sim, g, "3_3")


def test_emulated_random_state_and_medium(emu):
    rng = np.random.default_rng(7)
    nx, ny, npml, ns = 66, 140, 8, 21
    naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(np.float32)
    sim = _sim_for("3_2", nx, ny, np.float32, npml=npml, naz=naz, device="cpu")
    g, src = cases.grid_program("3_2", nx, ny, ns, np.float32, npml=npml, naz=naz.copy())
    for name in ("dz", "hx", "hy", "ihx", "ihy"):
        a = rng.standard_normal((nx, ny)).astype(np.float32)
        sim.set(name, a)
        getattr(g, name)[...] = a
    sim.advance(ns)
    orc.advance_2d(g, src)
    _assert_same(sim, g, "3_2")


@pytest.mark.parametrize("prog", ["3_1", "3_2", "3_3"])
def test_emulated_advance_matches_reference_goldens(emu, prog):
    ref = cases.golden(f"drive_{prog}_f32")
    nx, ny, ns = int(ref["nx"]), int(ref["ny"]), int(ref["ns"])
    sim = _sim_for(prog, nx, ny, np.float32, npml=int(ref["npml"]) if "npml" in ref else 0, device="cpu")
    sim.advance(ns)
    for name in ("dz", "ez", "hx", "hy"):
        if prog == "3_1":
            assert np.array_equal(sim.get(name), ref[name]), name
        else:
            assert sim.get(name).tobytes() == ref[name].tobytes(), name


@pytest.mark.parametrize("tblock", [1, 4, None])
@pytest.mark.parametrize("nx,ny,npml", [(60, 72, 8), (300, 420, 10)])
def test_emulated_running_dft_carried_with_the_rows(emu, nx, ny, npml, tblock):
    from simulation_b200 import fd2d, surface
    ns = 27
    g, src = cases.grid_program("3_4", nx, ny, ns, np.float32, npml=npml, radius=0.2, dft=True)
    sim = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)),
                      naz=g.naz.copy(), nbz=g.nbz.copy(), freqs=g.freqs, device="cpu")
    sim.advance(10, tblock=tblock)
    sim.advance(ns - 10, tblock=tblock)
    orc.advance_2d(g, src)
    for name in ("dz", "ez", "iz", "hx", "hy", "ihx", "ihy", "r_pt", "i_pt", "r_in", "i_in"):
        got, want = sim.get(name), getattr(g, name)
        assert got.tobytes() == want.tobytes(), (name, np.argwhere(got != want)[:4].tolist())


@pytest.mark.parametrize("prog,nslab", [("3_3", 3), ("3_2", 2)])
def test_emulated_row_slabs_with_ghost_rows(emu, prog, nslab):
    """One Fdtd2D per row slab, T ghost rows per side, ghost rows refreshed from the neighbours between blocks (what
    slab.py does over NVLink): owned rows bit-identical to the monolithic oracle."""
    nx, ny, npml, T, nblocks, dtype = 96, 80, 8, 4, 5, np.float32
    cuts = np.linspace(0, nx, nslab + 1).astype(int)
    slabs = [_sim_for(prog, nx, ny, dtype, npml=npml, rows=(int(lo), int(hi)), ghost=T, tblock=T, device="cpu")
             for lo, hi in zip(cuts[:-1], cuts[1:])]
    names = ("dz", "hx", "hy", "ihx", "ihy")
    for _ in range(nblocks):
        for s in slabs:
            s.advance(T, lazy_ez=True)
        whole = {n: np.concatenate([s.get(n) for s in slabs]) for n in names}
        for s in slabs:
            for n in names:
                s.set(n, whole[n])                   # whole-grid array: fills owned AND ghost rows
    for s in slabs:
        s.advance(1)                                 # a last non-lazy step stores Ez
    ns = nblocks * T + 1
    g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml)
    orc.advance_2d(g, src)
    for n in names + ("ez",):
        got = np.concatenate([s.get(n) for s in slabs])
        assert got.tobytes() == getattr(g, n).tobytes(), n


@pytest.mark.parametrize("schedule", ["skewed", "wavefront"])
@pytest.mark.parametrize("plan,ns", [(4, 40), ([48, 96, 192], 26), ([24, 30, 24] * 30, 19), (2, 7)])
def test_emulated_streamed_run(emu, plan, ns, schedule):
    """run_streamed in ISSUE order (the emulated library runs every launch at once): the row ranges of the skewed
    (parallelogram) and of the wavefront schedule, the upload / download slices and the ping-pong bookkeeping give the
    plain run's bits.  (What concurrent streams may reorder is the GPU tests' business.)"""
    from simulation_b200 import fd2d, surface
    rng = np.random.default_rng(5)
    nx, ny, npml = 420, 132, 12
    naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(np.float32)
    src = fd2d.PointSource(nx // 2 - 5, ny // 2 - 5, surface.Sinusoid(1500e6))
    a = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, naz=naz, device="cpu")
    a.advance(ns)
    b = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, device="cpu")
    host_ez = torch.empty((nx, ny), dtype=torch.float32)
    kw = {"block_rows": plan} if isinstance(plan, list) else {"blocks": plan}
    b.run_streamed(ns, torch.from_numpy(naz), host_ez, streams=5, schedule=schedule, **kw)
    assert torch.equal(host_ez, a.tensor("ez"))
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        assert torch.equal(a.tensor(name), b.tensor(name)), name
    assert b.t == ns and float(host_ez.abs().max()) > 1e-3


@pytest.mark.parametrize("prog,plan,ns,tblock,schedule", [("3_3", 4, 41, None, "skewed"), ("3_4", [40, 90, 130], 33, None, "skewed"),
                                                          ("3_4", [40, 90, 130], 33, None, "wavefront"), ("3_3", 3, 29, 8, "skewed"),
                                                          ("3_3", 3, 29, 8, "wavefront"), ("3_4", 2, 50, 4, "skewed")])
def test_emulated_streamed_run_tfsf_and_lossy(emu, prog, plan, ns, tblock, schedule):
    """run_streamed with a TFSF plane wave (the incident line advanced once per pass level up front, every block's pass
    reading that level's history) and with the lossy cylinder of program 3_4 (nbz streamed beside naz; iz carried):
    the plain run's bits on every array, the incident line included."""
    from simulation_b200 import fd2d, surface
    nx, ny, npml = 400, 132, 12
    a = _sim_for(prog, nx, ny, np.float32, npml=npml, radius=0.3, device="cpu")
    a.advance(ns, tblock=tblock)
    naz, nbz = a.naz.clone(), (a.nbz.clone() if prog == "3_4" else None)
    kw = dict(nbz=torch.zeros_like(nbz)) if nbz is not None else {}
    b = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), device="cpu", **kw)
    host_ez = torch.empty((nx, ny), dtype=torch.float32)
    bk = {"block_rows": plan} if isinstance(plan, list) else {"blocks": plan}
    b.run_streamed(ns, naz, host_ez, streams=5, schedule=schedule, tblock=tblock, nbz_host=nbz, **bk)
    assert torch.equal(host_ez, a.tensor("ez"))
    names = ["dz", "ez", "hx", "hy", "ihx", "ihy"] + (["iz"] if prog == "3_4" else [])
    for name in names:
        assert torch.equal(a.tensor(name), b.tensor(name)), name
    for name in ("ezi", "hxi", "bc"):
        assert a.get(name).tobytes() == b.get(name).tobytes(), name
    assert b.t == ns and float(host_ez.abs().max()) > 1e-3
    b.advance(7)                                   # and the run goes on from there (incident line, lossy box, ping-pong)
    a.advance(7)
    for name in names:
        assert torch.equal(a.tensor(name), b.tensor(name)), name


@pytest.mark.parametrize("schedule", ["skewed", "wavefront"])
def test_emulated_streamed_run_on_a_slab(emu, schedule):
    from simulation_b200 import fd2d, surface
    rng = np.random.default_rng(9)
    nx, ny, npml, rows, ghost, ns = 500, 132, 12, (150, 330), 24, 24
    naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(np.float32)
    src = fd2d.PointSource(nx // 2 - 5, ny // 2 - 5, surface.Sinusoid(1500e6))
    whole = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, naz=naz, device="cpu")
    whole.advance(ns)
    part = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, rows=rows, ghost=ghost, device="cpu")
    lo, hi = part.row_base, part.row_base + part.rows_alloc
    host_ez = torch.empty((rows[1] - rows[0], ny), dtype=torch.float32)
    part.run_streamed(ns, torch.from_numpy(naz[lo:hi].copy()), host_ez, blocks=3, streams=3, schedule=schedule)
    assert torch.equal(host_ez, whole.tensor("ez")[rows[0]:rows[1]])
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        assert torch.equal(part.tensor(name), whole.tensor(name)[rows[0]:rows[1]]), name


class _SyncWords:
    """stands in for the library-allocated sync words of slab.py: {flag from up, flag from down, counter 0, counter 1,
    error word, reserved x 3}"""

    def __init__(self):
        self.words = np.zeros(8, dtype=np.int64)
        self.ptr = self.words.ctypes.data


@pytest.mark.parametrize("prog,nslab,order", [("3_2", 2, "down"), ("3_3", 3, "up"), ("3_2", 4, "random"), ("3_4", 3, "random")])
def test_emulated_fused_halo_exchange(emu, prog, nslab, order):
    """The halo exchange fused into the pass (multi-GPU, slab.py's p2p mode) with every "GPU" a slab in this process:
    the careful warps store their edge rows straight into the neighbours' ghost rows (here: the neighbours' host
    arrays) and publish an epoch flag; each pass first checks its neighbours' flags.  Ranks run one after the other
    within an epoch, in varying order -- any order is legal once the previous epoch is complete everywhere.  Owned
    rows come out bit-identical to the monolithic oracle; no NCCL-style exchange happens anywhere."""
    from simulation_b200 import fd2d
    nx, ny, npml, T, nblocks, dtype = 200, 300, 8, 6, 5, np.float32
    cuts = np.linspace(0, nx, nslab + 1).astype(int)
    slabs = [_sim_for(prog, nx, ny, dtype, npml=npml, radius=0.3, rows=(int(lo), int(hi)), ghost=T, tblock=T, device="cpu")
             for lo, hi in zip(cuts[:-1], cuts[1:])]
    names = [n for n in fd2d.FIELD_NAMES if n != "ez" and (n != "iz" or prog == "3_4")]
    for s in slabs:
        s._sync_words = _SyncWords()
    for r, s in enumerate(slabs):
        def peer(q):
            if q is None:
                return None
            o = slabs[q]
            return {"row_base": o.row_base, "sync": o._sync_words.ptr,
                    "sets": [{n: o._sets[k][n].data_ptr() for n in names} for k in range(2)]}
        s.p2p = {"halo": T, "sync": s._sync_words, "up": peer(r - 1 if r > 0 else None), "dn": peer(r + 1 if r < nslab - 1 else None)}
    rng = np.random.default_rng(3)
    for epoch in range(1, nblocks + 1):
        ranks = list(range(nslab))
        if order == "up":
            ranks.reverse()
        elif order == "random":
            rng.shuffle(ranks)
        for r in ranks:
            slabs[r].advance(T, tblock=T, lazy_ez=epoch < nblocks, epoch=epoch)
    for r, s in enumerate(slabs):                      # every rank announced every epoch to each neighbour it has
        assert s._sync_words.words[0] == (nblocks if r > 0 else 0) and s._sync_words.words[1] == (nblocks if r < nslab - 1 else 0)
    ns = nblocks * T
    g, src = cases.grid_program(prog, nx, ny, ns, dtype, npml=npml, radius=0.3, dft=False)
    orc.advance_2d(g, src)
    for n in names + ["ez"]:
        got = np.concatenate([s.get(n) for s in slabs])
        assert got.tobytes() == getattr(g, n).tobytes(), (n, np.argwhere(got != getattr(g, n))[:4].tolist())


# ---------------------------------------------------------------------------------------- deep passes (fd2d_deep.cu)
@pytest.mark.parametrize("tblock,chunk_rows", [(12, 40), (8, 40), (12, 0), (8, 64)])
@pytest.mark.parametrize("prog,nx,ny,npml", [("3_2", 300, 1100, 8), ("3_3", 280, 1000, 12)])
def test_emulated_deep_passes(emu, prog, nx, ny, npml, tblock, chunk_rows):
    """Depth 8 and 12: the warp-chain interior kernel (TMA boxes, mbarrier hand-off between the warps of a group; its
    column and row variants for the PML strips / chunks) + the shared-memory-ring careful kernel, on grids with a true
    interior; ragged last strips and chunks, a remainder pass of the register-pipeline kernels at the end."""
    before = emu.emu_launches()
    sim = _run(emu, prog, nx, ny, npml, 2 * tblock + 3, np.float32, tblock, radius=0.3, tune=(4, chunk_rows, 0, 0, 0))
    emu.fdtd2d_tune(4, chunk_rows, 0, 0, 0)
    try:
        assert sim.pass_depths(2 * tblock + 3, tblock) == [tblock, tblock, 3]
    finally:
        emu.fdtd2d_tune(0, 0, 0, 0, 0)
    assert sim.max_tblock == 12
    if chunk_rows:       # the two identity checks of the setup + 3 passes x (careful + interior [+ incident line]); the two deep
        # passes of a problem without TFSF also launch the column and the row variant of the warp-chain kernel
        want = 2 + (3 * 3 if prog == "3_3" else 3 * 2 + 2 * 2)
        assert emu.emu_launches() - before == want, "an interior kernel did not run in every pass"


@pytest.mark.parametrize("prog,nx,ny,npml,vec,chunk,tblock", [("3_2", 420, 600, 8, 4, 64, 6), ("3_3", 400, 560, 12, 4, 96, 8),
                                                             ("3_4", 380, 520, 10, 2, 64, 6), ("3_3", 500, 300, 20, 2, 48, 4)])
def test_emulated_short_edge_chunks(emu, prog, nx, ny, npml, vec, chunk, tblock):
    """The first and the last row chunk are cut just tall enough to hold the rows that need the careful kernel (PML,
    grid edge, TFSF box rows); ordinary chunks lie between.  With and without the short edge chunks: the oracle's bits,
    and FEWER careful warps with them (the launch count is the same, so compare the two runs' arrays only)."""
    ns = 2 * tblock + 3
    a = _run(emu, prog, nx, ny, npml, ns, np.float32, tblock, radius=0.3, tune=(vec, chunk, 0, 0, 0), edge=1)
    b = _run(emu, prog, nx, ny, npml, ns, np.float32, tblock, radius=0.3, tune=(vec, chunk, 0, 0, 0), edge=0)
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        assert a.get(name).tobytes() == b.get(name).tobytes(), name


@pytest.mark.parametrize("variant,fast,tblock", [(3, 3, 8), (1, 3, 8), (2, 3, 8), (3, 3, 12),          # shared-memory-accumulator kernels
                                                 (11, 3, 8), (12, 3, 8), (13, 3, 8), (11, 3, 12), (12, 3, 12),   # other warp-chain shapes
                                                 (0, 0, 8), (0, 1, 12), (0, 2, 8)])                      # chain without / with one of its variants
def test_emulated_deep_kernel_variants(emu, variant, fast, tblock):
    """Every interior kernel of the deep passes on the CTA emulator: the shared-memory-accumulator kernels of
    fd2d_deep.cu (variants 1..3: the fallback without a tensor-map encoder), the warp-chain shapes that are not shipped
    (run-time staging ring with 2- and 3-row TMA boxes, 64-bit queue stores spelled in PTX, four warps of two stages, ...)
    and the shipped shape with its column / row variants switched off one by one -- the oracle's bits each time."""
    _run(emu, "3_2", 300, 1100, 8, 2 * tblock + 3, np.float32, tblock, tune=(4, 40, 0, 0, 0), variant=variant, fast=fast)


def test_emulated_deep_equals_register_pipeline(emu):
    """Same problem through the deep passes (8 + 8 + 8), the register-pipeline kernels alone (deep off: 8 + 8 + 8 with
    the naz ring) and depth 6: all arrays byte-identical."""
    nx, ny, npml, ns = 300, 900, 10, 24
    a = _run(emu, "3_3", nx, ny, npml, ns, np.float32, 8, tune=(4, 48, 0, 0, 0), deep=1)
    b = _run(emu, "3_3", nx, ny, npml, ns, np.float32, 8, tune=(4, 48, 0, 0, 0), deep=0)
    c = _run(emu, "3_3", nx, ny, npml, ns, np.float32, 6, tune=(4, 48, 0, 0, 0))
    for name in ("dz", "ez", "hx", "hy", "ihx", "ihy"):
        assert a.get(name).tobytes() == b.get(name).tobytes() == c.get(name).tobytes(), name


@pytest.mark.parametrize("prog,tblock,vec", [("3_2", 6, 4), ("3_3", 4, 4), ("3_4", 6, 4), ("3_4", 3, 4), ("3_1", 8, 4), ("3_3", 12, 4),
                                              ("3_4", 6, 2), ("3_3", 4, 2), ("3_2", 8, 2)])
def test_emulated_ring_careful_kernel_at_every_depth(emu, prog, tblock, vec):
    """fdtd2d_tune2(deep = 2): the shared-memory-ring careful kernel (run-time depth, rolled stages) replaces the
    register-shifting one in every float pass of 4- or 2-wide vectors -- lossless and lossy, TFSF, point source; once
    for the whole grid (force_careful) and once beside the interior kernels."""
    for force in (1, 0):
        _run(emu, prog, 150, 560, 0 if prog == "3_1" else 9, 2 * tblock + 1, np.float32, tblock, radius=0.3,
             tune=(vec, 24, 0, 0, force), deep=2)


def test_emulated_plan_query(emu):
    """fdtd2d_plan is what advance() does: depths by grid size and step count, without launching anything."""
    from simulation_b200 import fd2d
    before = emu.emu_launches()
    assert fd2d.plan_depths(32768, 32768, np.float32, 20) == [12, 8]                  # the driver's bench call: least-cost split
    assert fd2d.plan_depths(16384, 16384, np.float32, 20) == [8, 6, 6]                # (depth 12 is offered from 500 M cells up)
    assert fd2d.plan_depths(32768, 32768, np.float32, 96) == [8] * 12
    assert fd2d.plan_depths(32768, 32768, np.float32, 10) == [6, 4]                   # (a depth-8 pass costs 1.09 depth-6 passes)
    assert fd2d.plan_depths(32768, 32768, np.float32, 96, tblock=12) == [12] * 8
    assert fd2d.plan_depths(32768, 32768, np.float32, 96, tblock=6) == [6] * 16
    assert fd2d.plan_depths(32768, 32768, np.float32, 31, tblock=7) == [6] * 5 + [1]
    assert fd2d.plan_depths(32768, 32768, np.float64, 20) == [6, 6, 6, 2]               # float64: register pipeline only
    assert fd2d.plan_depths(32768, 32766, np.float32, 20) == [6, 6, 6, 2]               # ny not a multiple of 4
    assert fd2d.plan_depths(32768, 32768, np.float32, 20, lossy=True) == [6, 6, 6, 2]   # lossy: 2-wide register pipeline
    assert fd2d.plan_depths(32768, 32768, np.float32, 20, nf=3) == [4] * 5              # fused DFT: depth <= 4
    assert fd2d.plan_depths(1024, 1024, np.float32, 20) == [4] * 5                      # small grids: shallow, many warps
    assert fd2d.plan_depths(32768, 32768, np.float32, 8, rows=(4096, 8192)) == [8]      # a slab of the 8-GPU strong split
    emu.fdtd2d_tune2(0, 0)
    try:
        assert fd2d.plan_depths(32768, 32768, np.float32, 20) == [6, 6, 6, 2]
    finally:
        emu.fdtd2d_tune2(0, 1)
    assert emu.emu_launches() == before


def test_emulated_fused_halo_exchange_deep_and_missing_neighbour(emu):
    """The fused exchange at depth 12 (deep interior + ring careful kernel carrying the handshake), then a rank that
    SKIPS an epoch: its neighbour's wait is bounded -- it gives up, raises the error word and the host reports it
    instead of hanging."""
    import ctypes as C
    from simulation_b200 import _lib, fd2d
    nx, ny, npml, T, nslab, dtype = 360, 640, 8, 12, 3, np.float32
    emu.fdtd2d_tune(4, 24, 0, 0, 0)
    try:
        cuts = np.linspace(0, nx, nslab + 1).astype(int)
        slabs = [_sim_for("3_2", nx, ny, dtype, npml=npml, rows=(int(lo), int(hi)), ghost=T, tblock=T, device="cpu")
                 for lo, hi in zip(cuts[:-1], cuts[1:])]
        names = [n for n in fd2d.FIELD_NAMES if n not in ("ez", "iz")]
        for s in slabs:
            s._sync_words = _SyncWords()
        for r, s in enumerate(slabs):
            def peer(q):
                if q is None:
                    return None
                o = slabs[q]
                return {"row_base": o.row_base, "sync": o._sync_words.ptr,
                        "sets": [{n: o._sets[k][n].data_ptr() for n in names} for k in range(2)]}
            s.p2p = {"halo": T, "sync": s._sync_words, "up": peer(r - 1 if r > 0 else None), "dn": peer(r + 1 if r < nslab - 1 else None)}
        for epoch in (1, 2, 3):
            for r in (1, 0, 2):
                slabs[r].advance(T, tblock=T, lazy_ez=epoch < 3, epoch=epoch)
        g, src = cases.grid_program("3_2", nx, ny, 3 * T, dtype, npml=npml, dft=False)
        orc.advance_2d(g, src)
        for n in names + ["ez"]:
            got = np.concatenate([s.get(n) for s in slabs])
            assert got.tobytes() == getattr(g, n).tobytes(), (n, np.argwhere(got != getattr(g, n))[:4].tolist())
        assert all(int(s._sync_words.words[4]) == 0 for s in slabs)
        # epoch 4: rank 0 runs, rank 1 skips it; epoch 5 on rank 0 then needs rank 1's epoch-4 flag, which never comes
        emu.fdtd2d_tune2(1, 50)                    # bound the wait to 50 ms
        slabs[0].advance(T, tblock=T, epoch=4)
        slabs[0].advance(T, tblock=T, epoch=5)     # returns (the emulated kernel gives up after the bound)
        word = C.c_ulonglong(0)
        prob = _lib.Problem2D()
        prob.sync_local = slabs[0]._sync_words.ptr
        assert emu.fdtd2d_halo_status(C.byref(prob), C.byref(word)) == 0
        assert word.value == (5 << 2) | 2, word.value          # epoch 5, lower neighbour
        assert b"gave up" in emu.fdtd_last_error() and b"lower" in emu.fdtd_last_error()
    finally:
        emu.fdtd2d_tune(0, 0, 0, 0, 0)
        emu.fdtd2d_tune2(1, 0)


def test_emulated_more_than_three_dft_frequencies(emu):
    """advance() with 4 frequencies: the fused kernels carry at most 3, so every step is one single-step pass (WITHOUT
    the accumulators attached) + the fourier kernel; bit-identical to the oracle's per-step accumulation."""
    nx, ny, npml, ns = 60, 72, 8, 23
    freqs = [50e6, 300e6, 700e6, 1100e6]
    from simulation_b200 import fd2d, surface
    g, src = cases.grid_program("3_4", nx, ny, ns, np.float32, npml=npml, radius=0.12, dft=True, freqs=freqs)
    sim = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)),
                      naz=g.naz.copy(), nbz=g.nbz.copy(), freqs=g.freqs, device="cpu")
    sim.advance(9)
    sim.advance(ns - 9, tblock=4)
    orc.advance_2d(g, src)
    for name in ("dz", "ez", "iz", "hx", "hy", "r_pt", "i_pt", "r_in", "i_in"):
        got, want = sim.get(name), getattr(g, name)
        assert got.tobytes() == want.tobytes(), (name, np.argwhere(got != want)[:4].tolist())


def test_emulated_reference_positional_argument_lists(emu):
    """The module-level step functions take the reference's own positional argument lists (fd2d/program/fd2d_3_1.py:44,
    fd2d_3_3.py:68,81; fd2d/python/fd2d_3_4.py:131): dfield with and without pml / ezi, efield(md, dz, iz, ez)."""
    from simulation_b200 import fd2d, surface
    nx, ny, npml, ns = 48, 64, 6, 30
    rgrid = 9
    naz, nbz = surface.dielectric_cylinder(nx, ny, npml, rgrid, surface.DT, 30.0, 0.30, np.float32)
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32)
    ezi, hxi, bc = z(ny), z(ny), z(4)
    dz, ez, iz, hx, hy, ihx, ihy = (z(nx, ny) for _ in range(7))
    pml = fd2d.pmlparam(nx, ny, npml, np.float32, device="cpu")
    md = fd2d.medium(torch.from_numpy(naz), torch.from_numpy(nbz))
    wave = fd2d.IncidentWave(surface.Gaussian(20, 8.0))
    for t in range(1, ns + 1):
        fd2d.ezinct(ny, ezi, hxi, bc)
        fd2d.dfield(t, nx, ny, pml, ezi, dz, hx, hy, source=wave)          # reference 3_3 / 3_4 order
        fd2d.inctdz(nx, ny, npml, hxi, dz)
        fd2d.efield(nx, ny, md, dz, iz, ez)                                # reference 3_4 order
        fd2d.hxinct(ny, ezi, hxi)
        fd2d.hfield(nx, ny, pml, ez, ihx, ihy, hx, hy)
        fd2d.incthx(nx, ny, npml, ezi, hx)
        fd2d.incthy(nx, ny, npml, ezi, hy)
    g = orc.Grid2D(nx, ny, npml, np.float32, tfsf=True, lossy=True, naz=naz.copy(), nbz=nbz.copy())
    orc.advance_2d(g, orc.source_table("gaussian", ns, t0=20, spread=8.0))
    for name, got in (("dz", dz), ("ez", ez), ("iz", iz), ("hx", hx), ("hy", hy), ("ihx", ihx), ("ihy", ihy)):
        assert got.numpy().tobytes() == getattr(g, name).tobytes(), name
    # free space (program 3_1): dfield(t, nx, ny, dz, hx, hy) without a pmlayer, efield(nx, ny, naz, dz, ez)
    dz, ez, hx, hy = (z(nx, ny) for _ in range(4))
    one = torch.ones(nx, ny)
    src = fd2d.PointSource(nx // 2, ny // 2, surface.Gaussian(20, 6.0))
    for t in range(1, 21):
        fd2d.dfield(t, nx, ny, dz, hx, hy, source=src)
        fd2d.efield(nx, ny, one, dz, ez)
        fd2d.hfield(nx, ny, fd2d.pmlparam(nx, ny, 0, np.float32, device="cpu"), ez, z(nx, ny), z(nx, ny), hx, hy)
    g = orc.Grid2D(nx, ny, 0, np.float32, point=(nx // 2, ny // 2))
    orc.advance_2d(g, orc.source_table("gaussian", 20, t0=20, spread=6.0))
    assert np.array_equal(ez.numpy(), g.ez) and np.array_equal(hx.numpy(), g.hx)
    with pytest.raises(TypeError):
        fd2d.efield(nx, ny, md, dz)                                        # ez missing
    with pytest.raises(Exception):
        fd2d.efield(nx, ny, md, dz, ez)                                    # lossy medium without iz


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_emulated_device_pmlparam(emu, dtype):
    """fdtd2d_pmlparam (float64 on the device, the host's cubes expanded) == surface.pmlparam == the reference's Python;
    without the host's cubes (double-double cube on the device) float32 is still bit-identical."""
    from simulation_b200 import fd2d, surface
    for nx, ny, npml in ((60, 60, 8), (100, 131, 0), (97, 64, 32), (401, 300, 150), (1024, 777, 80), (2, 2, 1), (7, 9, 3)):
        host = surface.pmlparam(nx, ny, npml, dtype)
        dev = fd2d.pmlparam(nx, ny, npml, dtype, device="cpu", where="device")
        for name, h, d in zip(host._fields, host, dev):
            assert d.numpy().tobytes() == h.tobytes(), (nx, ny, npml, name)
        own = fd2d.pmlparam(nx, ny, npml, dtype, device="cpu", where="device", host_cubes=False)
        for name, h, d in zip(host._fields, host, own):
            if dtype == np.float32:
                assert d.numpy().tobytes() == h.tobytes(), (nx, ny, npml, name)
            else:
                assert np.all(np.abs(d.numpy() - h) <= 4 * np.spacing(np.maximum(np.abs(h), 1e-3))), (nx, ny, npml, name)


def test_emulated_lossless_outside_split(emu):
    """A lossy object in free space: interior warps outside the object's box run the lossless kernel.  Same bits as the
    lossy kernel everywhere (split disabled), as the oracle, and the promise is re-derived when iz is uploaded."""
    nx, ny, npml, ns = 300, 900, 10, 20
    a = _run(emu, "3_4", nx, ny, npml, ns, np.float32, 6, radius=0.25)
    r0, r1, c0, c1 = a._lossy_box()
    assert 0 < r1 - r0 < 60 and 0 < c1 - c0 < 60 and a.check_lossless_outside() == 0        # rgrid = 24 cells
    b = _run(emu, "3_4", nx, ny, npml, ns, np.float32, 6, radius=0.25, tune=(0, 0, 0, 0, 2))   # lossy kernel everywhere
    for name in ("dz", "ez", "iz", "hx", "hy", "ihx", "ihy"):
        assert a.get(name).tobytes() == b.get(name).tobytes(), name
    # the wave must have reached the object for iz to mean anything: a narrow grid, enough steps
    d = _run(emu, "3_4", 300, 200, npml, 110, np.float32, 6, radius=0.25)
    assert float(np.abs(d.get("iz")).max()) > 0 and 0 < d._lossy_box()[1] - d._lossy_box()[0] < 60
    # iz uploaded from outside, non-zero far from the cylinder: the box must grow to hold it, results stay exact
    rng = np.random.default_rng(2)
    g, src = cases.grid_program("3_4", nx, ny, ns, np.float32, npml=npml, radius=0.25, dft=False)
    c = _sim_for("3_4", nx, ny, np.float32, npml=npml, radius=0.25, device="cpu")
    iz = np.zeros((nx, ny), dtype=np.float32)
    iz[40:44, 700:720] = rng.standard_normal((4, 20)).astype(np.float32)
    iz[250, 100] = -0.0                                     # a negative zero is not +0 either
    c.set("iz", iz)
    g.iz[...] = iz
    box = c._lossy_box()
    assert box[0] <= 40 and box[1] >= 251 and box[2] <= 100 and box[3] >= 720 and c.check_lossless_outside() == 0
    c.advance(ns)
    orc.advance_2d(g, src)
    _assert_same(c, g, "3_4")
    # a stale promise is caught by the device-side check
    c._lossy_box_cache = (r0, r1, c0, c1)
    c.tensor("iz")[5, 5] = 1.0
    assert c.check_lossless_outside() >= 1


def test_emulated_checkpoint_restore(emu):
    nx, ny, npml = 72, 100, 8
    one = _sim_for("3_4", nx, ny, np.float32, npml=npml, device="cpu")
    one.advance(17)
    two = _sim_for("3_4", nx, ny, np.float32, npml=npml, device="cpu")
    two.restore(one.checkpoint())
    two.advance(20)
    g, src = cases.grid_program("3_4", nx, ny, 37, np.float32, npml=npml, dft=False)
    orc.advance_2d(g, src)
    _assert_same(two, g, "3_4")


def test_emulated_identity_promise_is_checked(emu):
    sim = _sim_for("3_2", 128, 160, np.float32, npml=8, device="cpu")
    assert sim.check_identity() == 0
    sim.pml.gy2[40] = 0.5
    assert sim.check_identity() == 1


def test_emulated_device_cylinder_rasteriser(emu):
    """fdtd2d_dielectric_cylinder (k_cylinder) against the host evaluation of the reference's Python statements."""
    from simulation_b200 import fd2d, surface
    nx, ny, npml, rgrid = 90, 120, 8, 17
    md = fd2d.dielectric(nx, ny, npml, rgrid, surface.DT, 30.0, 0.30, np.float32, device="cpu")
    naz, nbz = surface.dielectric_cylinder(nx, ny, npml, rgrid, surface.DT, 30.0, 0.30, np.float32)
    assert md.naz.numpy().tobytes() == naz.tobytes() and md.nbz.numpy().tobytes() == nbz.tobytes()


def test_plain_c_host_runs_against_the_emulated_library(tmp_path):
    """examples/c_host_3_3.c -- a C main() shaped like the reference's fd2d/cuda/test_3_3.cu, calling nothing but the C
    ABI of include/fdtd_b200.h -- linked against the emulated build: the reference-named step functions in a loop and
    one fused fdtd2d_advance give identical bytes (the program's own check), without a GPU."""
    import os
    import subprocess
    from tests.emu import build_emu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = build_emu.build_library()
    exe = tmp_path / "c_host_3_3_emu"
    subprocess.run(["gcc", "-Wall", "-O1", os.path.join(root, "examples", "c_host_3_3.c"), "-I", os.path.join(root, "include"),
                    so, "-lm", "-lstdc++", "-pthread", "-o", str(exe)], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(so) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe), "120", "200", "40", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical bytes" in r.stdout

