"""GPU parity tests of the 1D Ex/Hy path (reference-named functions and the fused time-blocked advance)
against the numpy oracle and the reference-made goldens.  Bar: bit-exact, fp32 and fp64."""
import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from tests import cases

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _sim_for(prog, nx, dtype, **kw):
    from simulation_b200 import fd1d, surface
    S, DT = fd1d.LineSource, surface.DT
    if prog in ("1_1", "1_2"):
        return fd1d.Fdtd1D(nx, dtype, abc=(prog == "1_2"), source=S(nx // 2, surface.Gaussian(40, 12.0), hard=True), **kw)
    if prog in ("1_3", "1_4", "1_5"):
        ca, cb = surface.dielectric_fdtd(nx, DT, 4.0, 0.04 if prog == "1_5" else 0.0, dtype)
        w = surface.Gaussian(40, 12.0) if prog == "1_3" else surface.Sinusoid(700e6)
        return fd1d.Fdtd1D(nx, dtype, source=S(1, w), ca=(ca if prog == "1_5" else None), cb=cb, **kw)
    if prog == "2_1":
        nax, nbx, _, _ = surface.dielectric_flux(nx, DT, 4.0, 0.04, dtype)
        return fd1d.Fdtd1D(nx, dtype, form="flux", source=S(1, surface.Sinusoid(700e6), field="dx"), nax=nax, nbx=nbx, **kw)
    if prog == "2_2":
        nax, nbx, _, _ = surface.dielectric_flux(nx, DT, 4.0, 0.0, dtype)
        return fd1d.Fdtd1D(nx, dtype, form="flux", source=S(1, surface.Gaussian(50, 10.0), field="dx"), nax=nax, nbx=nbx, **kw)
    if prog == "2_3":
        nax, nbx, ncx, ndx = surface.dielectric_flux(nx, DT, 2.0, 0.01, dtype, chi=2.0, tau=0.001e-6)
        return fd1d.Fdtd1D(nx, dtype, form="flux", source=S(1, surface.Gaussian(50, 10.0), field="dx"),
                           nax=nax, nbx=nbx, ncx=ncx, ndx=ndx, **kw)
    raise KeyError(prog)


def _oracle(prog, nx, ns, dtype):
    p, src = cases.line_program(prog, nx, ns, dtype)
    p.freqs = None                       # fields do not depend on the running DFT (tested separately below)
    orc.advance_1d(p, src)
    return p


def _assert_same(sim, p):
    names = ["ex", "hy"] + (["dx", "ix"] if p.form == "flux" else [])
    if p.form == "flux" and sim.debye:
        names.append("sx")
    if p.abc:
        names.append("bc")
    for n in names:
        got, want = sim.get(n), getattr(p, n)
        if got.tobytes() != want.tobytes():
            bad = np.argwhere(got != want)
            raise AssertionError(f"{n}: {len(bad)} cells differ, first {bad[:5].ravel().tolist()}")


ALL = ["1_1", "1_2", "1_3", "1_4", "1_5", "2_1", "2_2", "2_3"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("prog", ALL)
def test_step_functions_match_oracle(prog, dtype):
    nx, ns = 300, 700
    sim = _sim_for(prog, nx, dtype)
    for _ in range(ns):
        sim.step()
    _assert_same(sim, _oracle(prog, nx, ns, dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tblock", [1, 7, 32, 64])
@pytest.mark.parametrize("prog", ALL)
def test_advance_matches_oracle(prog, tblock, dtype):
    nx, ns = 5000, 1203                  # three CTAs, ragged last segment, ns not a multiple of tblock
    sim = _sim_for(prog, nx, dtype)
    sim.advance(ns, tblock=tblock)
    _assert_same(sim, _oracle(prog, nx, ns, dtype))


@pytest.mark.parametrize("prog", ALL)
def test_reference_main_goldens_fp64(prog):
    nx, ns = cases.LINE_MAIN[prog]
    sim = _sim_for(prog, nx, np.float64)
    sim.advance(ns)
    assert sim.get("ex").tobytes() == cases.golden(f"main_fd1d_{prog}")["ex"].tobytes()


@pytest.mark.parametrize("prog", ALL)
def test_reference_twin_goldens_fp32(prog):
    g = cases.golden(f"twin_fd1d_{prog}")
    sim = _sim_for(prog, int(g["nx"]), np.float32)
    sim.advance(int(g["ns"]))
    assert sim.get("ex").tobytes() == g["ex"].tobytes()


def test_baseline_config_1_free_space_pulse():
    """BASELINE config 1: ke=200, 100 steps, hard Gaussian in the middle (fd1d_1_1 rule), fp64."""
    sim = _sim_for("1_1", 200, np.float64)
    sim.advance(100)
    _assert_same(sim, _oracle("1_1", 200, 100, np.float64))


def test_baseline_config_2_long_lossy_line():
    """BASELINE config 2 shape: 1e6 cells, lossy slab, sinusoid, ABC -- 600 steps vs the oracle, and the
    fused path against the unfused reference-named functions."""
    from simulation_b200 import fd1d, surface
    nx, ns = 1_000_000, 600
    ca, cb = surface.dielectric_fdtd(nx, surface.DT, 4.0, 0.04, np.float32, start=nx // 2, stop=nx // 2 + nx // 4)
    mk = lambda: fd1d.Fdtd1D(nx, np.float32, source=fd1d.LineSource(1, surface.Sinusoid(700e6)), ca=ca, cb=cb)
    a, b = mk(), mk()
    a.advance(ns, tblock=64)
    for _ in range(ns):
        b.step()
    oca, ocb = orc.lossy_halfspace_fdtd(nx, cases.DT, 4.0, 0.04, np.float32, start=nx // 2, stop=nx // 2 + nx // 4)
    p = orc.Line1D(nx, np.float32, ca=oca, cb=ocb)
    orc.advance_1d(p, orc.source_table("sine", ns, freq=700e6))
    for n in ("ex", "hy", "bc"):
        assert a.get(n).tobytes() == getattr(p, n).tobytes(), n
        assert b.get(n).tobytes() == getattr(p, n).tobytes(), n


@pytest.mark.parametrize("dtype,tblock", [(np.float32, 32), (np.float64, 16), (np.float32, 4), (np.float64, 10)])
def test_last_warp_owning_one_cell(dtype, tblock):
    """nx-1 a multiple of the segment length: the last warp owns cell nx-1 alone and the right-hand ABC's ex[nx-2] is its
    innermost halo cell (found by tools/fuzz_emulated_1d.py on the CPU emulator; non-zero state up to the line's end)."""
    vec, w = (4, 512) if dtype == np.float32 else (2, 256)
    halo = -(-tblock // vec) * vec
    nx = (w - 2 * halo) * {32: 3, 16: 6, 4: 2, 10: 5}[tblock] + 1
    ns = 2 * tblock + 7
    for prog in ("1_2", "1_5", "2_3"):
        p, src = cases.line_program(prog, nx, ns, dtype)
        p.freqs = None
        sim = _sim_for(prog, nx, dtype, tblock=tblock)
        rng = np.random.default_rng(4)
        names = ["ex", "hy", "bc"] + (["dx", "ix"] if p.form == "flux" else []) + (["sx"] if sim.debye else [])
        for n in names:
            a = getattr(p, n)
            a[...] = rng.uniform(-1, 1, a.shape).astype(dtype)
            sim.set(n, a)
        sim.advance(ns)
        orc.advance_1d(p, src)
        for n in names:
            assert sim.get(n).tobytes() == getattr(p, n).tobytes(), (prog, n)


def test_tiny_lines():
    for nx in (3, 4, 17):
        sim = _sim_for("1_2", nx, np.float64)
        sim.advance(40, tblock=5)
        _assert_same(sim, _oracle("1_2", nx, 40, np.float64))


# ------------------------------------------------------------------ running DFT (fourier), programs 2_2 / 2_3
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("prog", ["2_2", "2_3"])
def test_running_dft_matches_oracle(prog, dtype):
    from simulation_b200 import fd1d
    nx, ns = 400, 500
    p, src = cases.line_program(prog, nx, ns, dtype)
    sim = _sim_for(prog, nx, dtype, freqs=p.freqs)
    sim.advance(ns)
    orc.advance_1d(p, src)
    for n in ("ex", "hy", "r_pt", "i_pt", "r_in", "i_in"):
        assert sim.get(n).tobytes() == getattr(p, n).tobytes(), n


DFT_NAMES = ("ex", "hy", "dx", "ix", "r_pt", "i_pt", "r_in", "i_in")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tblock", [1, 3, 8, 16, 32])
@pytest.mark.parametrize("prog,nx", [("2_2", 397), ("2_3", 4099), ("2_2", 12), ("2_3", 1000)])
def test_running_dft_carried_through_the_passes(prog, nx, tblock, dtype):
    """The fused pass carries the DFT accumulators (k1_advance_dft): every pass depth, ragged and multi-warp lines,
    advance() split at odd places -- bitwise equal to the oracle AND to the per-step fdtd1d_fourier path."""
    ns = 230
    p, src = cases.line_program(prog, nx, ns, dtype)
    orc.advance_1d(p, src)
    fused = _sim_for(prog, nx, dtype, freqs=p.freqs, tblock=tblock)
    for part in (7, 100, 1, 122):
        fused.advance(part)
    assert fused.t == ns
    for n in DFT_NAMES:
        assert fused.get(n).tobytes() == getattr(p, n).tobytes(), n
    if tblock == 8:
        steps = _sim_for(prog, nx, dtype, freqs=p.freqs)
        steps.advance(ns, fused_dft=False)
        for n in DFT_NAMES:
            assert fused.get(n).tobytes() == steps.get(n).tobytes(), n


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tblock", [7, 32])
@pytest.mark.parametrize("prog", ALL)
def test_advance_from_a_random_state(prog, tblock, dtype):
    """Uploaded non-zero state everywhere (fields, ABC delay line, DFT accumulators): every segment boundary of every
    warp carries signal from the first step on, which the programs' own pulses reach only after hundreds of steps."""
    nx, ns = 3001, 75
    p, src = cases.line_program(prog, nx, ns, dtype)
    dft = p.freqs is not None
    sim = _sim_for(prog, nx, dtype, tblock=tblock, **({"freqs": p.freqs} if dft else {}))
    rng = np.random.default_rng(11)
    names = ["ex", "hy"] + (["bc"] if p.abc else []) + (["dx", "ix"] if p.form == "flux" else []) + (["sx"] if sim.debye else [])
    for n in names:
        a = getattr(p, n)
        a[...] = rng.uniform(-1, 1, a.shape).astype(dtype)
        sim.set(n, a)
    if dft:
        for n in ("r_pt", "i_pt", "r_in", "i_in"):
            a = getattr(p, n)
            a[...] = rng.uniform(-1, 1, a.shape).astype(dtype)
            getattr(sim.ft, n).copy_(torch.from_numpy(a))
    sim.advance(ns)
    orc.advance_1d(p, src)
    for n in names + (["r_pt", "i_pt", "r_in", "i_in"] if dft else []):
        assert sim.get(n).tobytes() == getattr(p, n).tobytes(), n


@pytest.mark.parametrize("nf", [1, 2, 4])
def test_running_dft_other_frequency_counts(nf):
    """nf < 3 leaves accumulator slots idle; nf > 3 exceeds the fused variant and takes the per-step path."""
    nx, ns, dtype = 700, 160, np.float32
    freqs = np.array((100e6, 200e6, 500e6, 900e6)[:nf], dtype=dtype)
    p, src = cases.line_program("2_2", nx, ns, dtype)
    p.freqs = freqs
    p.__post_init__()
    orc.advance_1d(p, src)
    sim = _sim_for("2_2", nx, dtype, freqs=freqs)
    sim.advance(ns)
    for n in DFT_NAMES:
        assert sim.get(n).tobytes() == getattr(p, n).tobytes(), n


def test_running_dft_survives_checkpoint_restore():
    nx, a_steps, b_steps, dtype = 520, 90, 75, np.float32
    p, src = cases.line_program("2_3", nx, a_steps + b_steps, dtype)
    orc.advance_1d(p, src)
    one = _sim_for("2_3", nx, dtype, freqs=p.freqs)
    one.advance(a_steps)
    two = _sim_for("2_3", nx, dtype, freqs=p.freqs)
    two.restore(one.checkpoint())
    two.advance(b_steps)
    for n in DFT_NAMES + ("sx",):
        assert two.get(n).tobytes() == getattr(p, n).tobytes(), n


@pytest.mark.parametrize("prog", ["2_2", "2_3"])
def test_dft_amplitude_goldens(prog):
    """amplt[2] as the reference programs plot it: fp64 main() golden and fp32 benchmark-twin golden."""
    for tag, dtype in (("main", np.float64), ("twin", np.float32)):
        g = cases.golden(f"{tag}_fd1d_{prog}")
        nx, ns = (cases.LINE_MAIN[prog] if tag == "main" else (int(g["nx"]), int(g["ns"])))
        p, _ = cases.line_program(prog, nx, ns, dtype)
        sim = _sim_for(prog, nx, dtype, freqs=p.freqs)
        sim.advance(ns)
        amp, _ = orc.dft_amplitude_phase(sim.get("r_pt"), sim.get("i_pt"), sim.get("r_in"), sim.get("i_in"))
        assert amp[2].tobytes() == g["amplt2"].tobytes(), tag


@pytest.mark.parametrize("prog", ["1_2", "1_5", "2_3"])
def test_checkpoint_restore_continues_bit_identically(prog):
    nx, a_steps, b_steps = 300, 130, 170
    one = _sim_for(prog, nx, np.float32)
    one.advance(a_steps)
    two = _sim_for(prog, nx, np.float32)
    two.restore(one.checkpoint())
    assert two.t == a_steps
    two.advance(b_steps)
    _assert_same(two, _oracle(prog, nx, a_steps + b_steps, np.float32))
