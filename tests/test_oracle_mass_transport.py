"""The reference's test/mass_transport.py restated on the oracle: a disc of floating ice spreading radially at a
uniform speed, thickness C / r outside the fixed disc.  Its golden numbers (mass_transport.py:169-173, the average
error at N = 51 and 101) pin the ADVECTIVE part of GeometryEvolution's interface fluxes -- upwinding, the velocity
at ice margins, limit_advective_velocity, the thickness Dirichlet mask in the flux divergence -- which pismv -test C
(zero sliding) never exercises.  The reference runs this test with geometry.part_grid.enabled; the oracle's
part_grid variant (orc_mass_flow_step_part_grid) shares the interface-flux code with the default variant the CUDA
kernels are checked against (orc_mass_flow_step)."""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle_lib as O
from pism_b200 import grid as G


def disc(grid, w, H0, R_inner, R_outer):
    """mass_transport.py:15-40 on a ghosted array (ghosts by periodic wrap, like update_ghosts)."""
    X, Y = np.meshgrid(grid.x, grid.y)
    d2 = X ** 2 + Y ** 2
    with np.errstate(divide="ignore"):
        a = np.where(d2 <= R_inner ** 2, H0, np.where(d2 <= R_outer ** 2, H0 * R_inner / np.sqrt(d2), 0.0))
    return G.wrap_ghosts(np.pad(a, w), w)


def run(N, t_final=1.0, Ccfl=1.0, part_grid=True):
    """mass_transport.py:63-146."""
    grid = G.Grid(N, N, 3, 1.0, 1.0, 1.0)
    cfg = cases.Cfg(smoother_range=0.0)
    p = cfg.oracle_params(grid)                 # GeometryEvolution's own gc: ice_free_thickness_standard
    p0 = cfg.oracle_params(grid)
    p0.ice_free_thickness = 0.0                 # geometry.ensure_consistency(0.0)
    w, ws = p.w_geom, p.w_sliding
    R_inner, speed = 0.25 * min(grid.Lx, grid.Ly), 0.7
    H = disc(grid, w, 1.0, R_inner, R_inner)
    Href = np.zeros_like(H)
    bed = np.full_like(H, -10.0)
    sea = np.zeros_like(H)
    X, Y = np.meshgrid(grid.x, grid.y)
    r = np.maximum(np.sqrt(X * X + Y * Y), 0.001)
    vel = G.wrap_ghosts(np.pad(np.stack([speed * X / r, speed * Y / r], axis=-1), ((ws, ws), (ws, ws), (0, 0))), ws)
    vel = np.ascontiguousarray(vel)
    Q = np.zeros(O.shape(p, p.w_stag, 2))
    v_bc = np.zeros_like(H)
    H_bc = disc(grid, w, 1.0, R_inner, R_inner)
    mask, surf = np.zeros_like(H), np.zeros_like(H)

    def ensure_consistency():
        sel = (H > 0.0) & (Href > 0.0)          # Geometry.cc:131-146
        H[sel] += Href[sel]
        Href[sel] = 0.0
        O.lib().orc_geometry_compute(C.byref(p0), H.size, O.dptr(sea), O.dptr(bed), O.dptr(H), O.dptr(mask),
                                     O.dptr(surf))

    ensure_consistency()
    outs = [np.zeros((N, N)) for _ in range(4)]
    t, steps = 0.0, 0
    out4 = (C.c_double * 4)()
    while t < t_final:
        assert O.lib().orc_cfl_2d(C.byref(p), 60.0 * 3.15e7, O.dptr(mask), O.dptr(vel), out4) == 0
        dt = out4[0] * Ccfl
        if t + dt > t_final:
            dt = t_final - t
        if part_grid:
            st = O.lib().orc_mass_flow_step_part_grid(C.byref(p), dt, O.dptr(sea), O.dptr(bed), O.dptr(H),
                                                      O.dptr(Href), O.dptr(vel), O.dptr(v_bc), O.dptr(H_bc),
                                                      O.dptr(Q), 10, *[O.dptr(o) for o in outs])
        else:
            st = O.lib().orc_mass_flow_step(C.byref(p), dt, O.dptr(sea), O.dptr(bed), O.dptr(H), O.dptr(vel),
                                            O.dptr(v_bc), O.dptr(H_bc), O.dptr(Q), O.dptr(outs[0]), O.dptr(outs[1]),
                                            O.dptr(outs[3]))
            G.wrap_ghosts(H, w)
        assert st == 0
        ensure_consistency()
        t += dt
        steps += 1
    total = cases.interior(H + Href, w)
    exact = cases.interior(disc(grid, w, 1.0, R_inner, R_inner + speed * t_final), w)
    return total, exact, steps


def test_part_grid_convergence_golden_numbers():
    """mass_transport.py:169-173: assert_almost_equal([average_error(51), average_error(101)],
    [0.0338388, 0.0158498]) -- seven decimals."""
    errs = []
    for N in (51, 101):
        total, exact, steps = run(N)
        errs.append(np.abs(exact - total).sum() / (N * N))
        print("N = %d: %d steps, average error %.7f" % (N, steps, errs[-1]))
    np.testing.assert_almost_equal(errs, [0.0338388, 0.0158498])


def test_part_grid_symmetry():
    """mass_transport.py:176-197."""
    H, _, _ = run(51)
    np.testing.assert_almost_equal(H, np.flipud(H))
    np.testing.assert_almost_equal(H, np.fliplr(H))
    np.testing.assert_almost_equal(H, np.flipud(np.fliplr(H)))


def test_default_variant_spreads_the_same_disc():
    """The variant the CUDA kernels implement (part_grid off) on the same setup and the same interface-flux code:
    the same disc to within the first-order error of either scheme, and symmetric."""
    a, exact, _ = run(51, part_grid=True)
    b, _, _ = run(51, part_grid=False)
    assert abs(a.sum() - b.sum()) <= 5e-3 * a.sum()
    assert np.abs(exact - b).sum() / 51 ** 2 < 0.06
    np.testing.assert_almost_equal(b, np.flipud(np.fliplr(b)))
