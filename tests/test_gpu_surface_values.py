"""-m gpu: the reads the path's callers make of its 3D outputs -- IceModelVec3::getSurfaceValues / getHorSlice
(util/iceModelVec3.cc:153-240) on the device, and PISM.sia.computeSIASurfaceVelocities
(site-packages/PISM/sia.py:24-74, exercised by test/miscellaneous.py:321-363) on top of them.  The interpolation
is bit-exact against the oracle on identical 3D data; through a whole update it inherits the 1e-10 of u, v."""
import ctypes as C

import numpy as np
import pytest

import cases
import gpu_util as U
import oracle_lib as O
from pism_b200 import capi, grid as G
from pism_b200.capi import F, lib
from pism_b200.sia import PISMRuntimeError, computeSIASurfaceVelocities

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name", ["Fs", "dome_33_13", "dome_64_41_quadratic", "C4s"])
def test_surface_values_match_oracle(name):
    grid, cfg, inputs, gb = cases.case(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert run.status == 0
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    H, wg = inputs["thickness"], cfg.w_geom
    for f in ("u", "v"):
        got = sia.getSurfaceValues(f)
        assert got.shape == (grid.My, grid.Mx)
        # the same interpolation of the same numbers: bit for bit
        assert np.array_equal(got, O.value_at_height(run.p, sia.download(f), cfg.w_uv, H, wg)), f
        # and through the whole update: the tolerance of u, v
        assert cases.rel_max(got, O.value_at_height(run.p, run.a[f], cfg.w_uv, H, wg)) <= U.TOL, f
        if name == "C4s":  # and against the frozen vectors of tools/make_golden.py
            import os
            d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_reads_C4s.npz"))
            assert cases.rel_max(got, d[f + "_surface"]) <= U.TOL, f


def test_horizontal_slices_and_the_other_3d_fields():
    grid, cfg, inputs, gb = cases.case("dome_33_13")
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    p, z, H, wg = cfg.oracle_params(grid), grid.z, inputs["thickness"], cfg.w_geom
    E = np.ascontiguousarray(inputs["enthalpy"], dtype=np.float64)
    sia.upload("enthalpy", E)  # the host-array update keeps the device copy current only near ice
    # below the base, the end levels, exactly on a level, between two levels, above the top (getValZ's branches)
    for z0 in (-5.0, 0.0, float(z[3]), float(0.25 * z[3] + 0.75 * z[4]), float(z[-1]), float(z[-1]) + 1.0):
        assert np.array_equal(sia.getHorSlice("enthalpy", z0), O.value_at_height(p, E, cfg.w_3d_in, z0=z0)), z0
    assert np.array_equal(sia.getSurfaceValues("enthalpy"), O.value_at_height(p, E, cfg.w_3d_in, H, wg))
    # fields without ghosts (StressBalance.cc:142)
    w = sia.compute_vertical_velocity()
    assert np.array_equal(sia.getSurfaceValues("w"), O.value_at_height(p, w, 0, H, wg))
    sigma = sia.compute_volumetric_strain_heating("gpbld", 3.0, 1.0)
    assert np.array_equal(sia.getSurfaceValues("strain_heating"), O.value_at_height(p, sigma, 0, H, wg))
    assert np.array_equal(sia.getHorSlice("u", 0.0), sia.download("u")[1:-1, 1:-1, 0])


def test_surface_values_errors():
    grid, cfg, inputs, gb = cases.case("dome_33_13")
    sia = U.make_sia(grid, cfg, gb)
    out = torch.zeros((grid.My, grid.Mx), dtype=torch.float64, device="cuda")
    # not computed yet; not a 3D field; no output array
    assert lib.siafd_b200_surface_values(sia.handle, F["u"], out.data_ptr()) == capi.ERR_BAD_ARGUMENT
    assert lib.siafd_b200_surface_values(sia.handle, F["flux"], out.data_ptr()) == capi.ERR_BAD_ARGUMENT
    assert lib.siafd_b200_hor_slice(sia.handle, F["thickness"], 0.0, out.data_ptr()) == capi.ERR_BAD_ARGUMENT
    U.gpu_update(sia, inputs, True)
    assert lib.siafd_b200_surface_values(sia.handle, F["u"], None) == capi.ERR_BAD_ARGUMENT
    with pytest.raises(PISMRuntimeError):
        sia.getSurfaceValues("D")
    assert lib.siafd_b200_surface_values(sia.handle, F["u"], out.data_ptr()) == capi.OK


def test_computeSIASurfaceVelocities_sia_test():
    """test/miscellaneous.py:321-363 (sia_test): a 100 x 100 x 11 slab of constant thickness at 270 K; the
    reference only checks that the call goes through -- a flat surface moreover means no flow at all."""
    grid = G.Grid(100, 100, 11, 1e5, 1e5, 1000.0)
    p = O.default_params()
    w = p.w_geom
    shape2 = (grid.My + 2 * w, grid.Mx + 2 * w)
    thk, bed = np.full(shape2, 1000.0), np.zeros(shape2)
    enthalpy = np.full((grid.My + 2 * p.w_3d_in, grid.Mx + 2 * p.w_3d_in, grid.Mz),
                       O.lib().orc_ec_enthalpy(C.byref(p), 270.0, 0.0, 0.0))
    us, vs = computeSIASurfaceVelocities(grid, thk, bed, enthalpy)
    assert us.shape == vs.shape == (100, 100)
    assert np.all(us == 0.0) and np.all(vs == 0.0)


def test_computeSIASurfaceVelocities_test_F():
    """The same helper on the Test F state: the surface velocities siafd_test.cc:105-151 compares with exactFG."""
    grid, cfg, inputs, _ = cases.case("Fs")
    run = cases.oracle_run(grid, cfg, inputs, full=True)
    us, vs = computeSIASurfaceVelocities(grid, inputs["thickness"], inputs["bed"], inputs["enthalpy"],
                                         **cfg.overrides())
    H, wg = inputs["thickness"], cfg.w_geom
    assert cases.rel_max(us, O.value_at_height(run.p, run.a["u"], cfg.w_uv, H, wg)) <= U.TOL
    assert cases.rel_max(vs, O.value_at_height(run.p, run.a["v"], cfg.w_uv, H, wg)) <= U.TOL
    Uex, r = cases.interior(inputs["exact_surface_speed"], wg), cases.interior(inputs["radius"], wg)
    X, Y = np.meshgrid(grid.x, grid.y)
    sel = (r >= 1.0) & (r <= 750000.0 - 1.0) & (cases.interior(H, wg) > 0)
    rr = np.where(sel, r, 1.0)
    err = np.hypot(us - X / rr * Uex, vs - Y / rr * Uex)[sel]
    assert err.max() < 0.2 * Uex.max()
