"""The product's setup surface (host numpy) against the oracle's and the reference-made goldens: the
coefficient arrays handed to the kernels must be bit-identical to the reference's."""
import numpy as np
import pytest

from oracle import fdtd_oracle as orc
from simulation_b200 import surface
from tests import cases


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nx,ny,npml", [(60, 60, 8), (40, 56, 0), (64, 48, 7), (1024, 768, 80)])
def test_pmlparam(nx, ny, npml, dtype):
    mine, ref = surface.pmlparam(nx, ny, npml, dtype), orc.pml_vectors(nx, ny, npml, dtype)
    for name in orc.PML_NAMES:
        assert getattr(mine, name).tobytes() == ref[name].tobytes(), name


def test_pmlparam_rejects_oversized_layer():
    with pytest.raises(ValueError):
        surface.pmlparam(20, 20, 11)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_waveforms(dtype):
    g = surface.Gaussian(20, 8.0).table(1, 300)
    s = surface.Sinusoid(1500e6).table(1, 300)
    assert g.tobytes() == orc.source_table("gaussian", 300, t0=20, spread=8.0).tobytes()
    assert s.tobytes() == orc.source_table("sine", 300, freq=1500e6).tobytes()
    # a table started mid-run is the tail of the full table (absolute step index)
    assert surface.Sinusoid(700e6).table(101, 50).tobytes() == orc.source_table("sine", 150, freq=700e6)[100:].tobytes()
    assert surface.Samples(g).table(11, 5).tobytes() == g[10:15].tobytes()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_media_1d(dtype):
    nx, dt = 512, surface.DT
    for sigma in (0.0, 0.04):
        a, b = surface.dielectric_fdtd(nx, dt, 4.0, sigma, dtype), orc.lossy_halfspace_fdtd(nx, dt, 4.0, sigma, dtype)
        assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    a = surface.dielectric_flux(nx, dt, 2.0, 0.01, dtype, chi=2.0, tau=0.001e-6)
    b = orc.lossy_halfspace_flux(nx, dt, 2.0, 0.01, dtype, chi=2.0, tau=0.001e-6)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    a = surface.dielectric_flux(nx, dt, 4.0, 0.04, dtype)
    b = orc.lossy_halfspace_flux(nx, dt, 4.0, 0.04, dtype)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))


def test_cylinder_matches_oracle_and_numba_golden():
    ref = cases.golden("drive_3_4_f64")
    nx, ny, npml = (int(ref[k]) for k in ("nx", "ny", "npml"))
    naz, nbz = surface.dielectric_cylinder(nx, ny, npml, int(ref["rgrid"]), surface.DT, 30.0, 0.30, np.float64)
    onaz, onbz = orc.cylinder_medium(nx, ny, npml, int(ref["rgrid"]), surface.DT, 30.0, 0.30, np.float64)
    assert naz.tobytes() == onaz.tobytes() and nbz.tobytes() == onbz.tobytes()
    assert np.abs(naz - ref["naz"]).max() <= 2.3e-16 and np.abs(nbz - ref["nbz"]).max() <= 2.3e-16
    # a row slab of the cylinder equals the same rows of the whole-grid arrays (multi-GPU setup)
    snaz, snbz = surface.dielectric_cylinder(nx, ny, npml, int(ref["rgrid"]), surface.DT, 30.0, 0.30, np.float64,
                                             rows=slice(17, 41))
    assert snaz.tobytes() == naz[17:41].tobytes() and snbz.tobytes() == nbz[17:41].tobytes()
