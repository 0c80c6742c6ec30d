"""-m gpu: the setup of the reference's test/mass_transport.py (spreading disc, uniform radial velocity, thickness
Dirichlet mask) stepped on the device: siafd_b200_mass_flow_step with a nonzero advective velocity and both B.C.
masks against the oracle (default variant, part_grid off), bit for bit.  The oracle's interface fluxes are pinned by
that test's golden numbers in tests/test_oracle_mass_transport.py."""
import ctypes as C

import numpy as np
import pytest

import cases
import gpu_util as U
import oracle_lib as O
from pism_b200 import grid as G
from pism_b200.capi import lib
from test_oracle_mass_transport import disc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [51, 64])
def test_spreading_disc_flow_steps_bit_exact(N):
    grid = G.Grid(N, N, 3, 1.0, 1.0, 1.0)
    cfg = cases.Cfg(smoother_range=0.0)
    cfg.w_sliding = 1
    p = cfg.oracle_params(grid)
    w = p.w_geom
    R_inner, speed = 0.25, 0.7
    H = disc(grid, w, 1.0, R_inner, R_inner)
    bed = np.full_like(H, -10.0)
    X, Y = np.meshgrid(grid.x, grid.y)
    r = np.maximum(np.sqrt(X * X + Y * Y), 0.001)
    vel = np.ascontiguousarray(G.wrap_ghosts(np.pad(np.stack([speed * X / r, speed * Y / r], axis=-1),
                                                    ((1, 1), (1, 1), (0, 0))), 1))
    H_bc = disc(grid, w, 1.0, R_inner, R_inner)
    v_bc = np.zeros_like(H)
    v_bc[w + N // 2 - 3:w + N // 2 + 3, w + N // 2 - 3:w + N // 2 + 3] = 1.0   # both branches of the velocity B.C. rule
    G.wrap_ghosts(v_bc, w)
    Q = np.zeros(O.shape(p, p.w_stag, 2))
    sea = np.zeros_like(H)
    mask, surf = np.zeros_like(H), np.zeros_like(H)

    sia = U.make_sia(grid, cfg)
    for name, a in (("thickness", H), ("bed", bed), ("sliding", vel), ("thk_bc_mask", H_bc), ("vel_bc_mask", v_bc),
                    ("flux", Q)):
        sia.upload(name, a)
    sia._check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
    O.lib().orc_geometry_compute(C.byref(p), H.size, O.dptr(sea), O.dptr(bed), O.dptr(H), O.dptr(mask), O.dptr(surf))
    assert np.array_equal(sia.download("mask"), mask)
    divQ, dH, ce = np.zeros((N, N)), np.zeros((N, N)), np.zeros((N, N))
    out4 = (C.c_double * 4)()
    for step in range(12):
        assert O.lib().orc_cfl_2d(C.byref(p), 1e9, O.dptr(mask), O.dptr(vel), out4) == 0
        dt = out4[0]
        assert O.lib().orc_mass_flow_step(C.byref(p), dt, O.dptr(sea), O.dptr(bed), O.dptr(H), O.dptr(vel),
                                          O.dptr(v_bc), O.dptr(H_bc), O.dptr(Q), O.dptr(divQ), O.dptr(dH),
                                          O.dptr(ce)) == 0
        G.wrap_ghosts(H, w)
        O.lib().orc_geometry_compute(C.byref(p), H.size, O.dptr(sea), O.dptr(bed), O.dptr(H), O.dptr(mask),
                                     O.dptr(surf))
        sia._check(lib.siafd_b200_mass_flow_step(sia.handle, dt))
        sia._check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
        assert np.array_equal(sia.download("flux_div"), divQ), step
        assert np.array_equal(sia.download("thk_change"), dH), step
        assert np.array_equal(sia.download("cons_err"), ce), step
        assert np.array_equal(sia.download("thickness"), H), step
        assert np.array_equal(sia.download("mask"), mask), step
    assert np.abs(cases.interior(H, w) - cases.interior(H_bc, w)).max() > 0.1   # the disc did spread
    assert set(np.unique(mask)) == {3.0, 4.0}                                    # floating ice and open ocean
