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


def spreading_disc_setup(N):
    """Arrays of the spreading-disc state on an N x N grid (mass_transport.py:63-110) plus a small velocity
    Dirichlet region in the middle, so that both branches of the velocity B.C. rule are taken."""
    grid = G.Grid(N, N, 3, 1.0, 1.0, 1.0)
    cfg = cases.Cfg(smoother_range=0.0)
    cfg.w_sliding = 1
    p = cfg.oracle_params(grid)
    w = p.w_geom
    R_inner, speed = 0.25, 0.7
    H = disc(grid, w, 1.0, R_inner, R_inner)
    X, Y = np.meshgrid(grid.x, grid.y)
    r = np.maximum(np.sqrt(X * X + Y * Y), 0.001)
    vel = np.ascontiguousarray(G.wrap_ghosts(np.pad(np.stack([speed * X / r, speed * Y / r], axis=-1),
                                                    ((1, 1), (1, 1), (0, 0))), 1))
    v_bc = np.zeros_like(H)
    v_bc[w + N // 2 - 3:w + N // 2 + 3, w + N // 2 - 3:w + N // 2 + 3] = 1.0
    G.wrap_ghosts(v_bc, w)
    S = dict(N=N, grid=grid, cfg=cfg, p=p, w=w, H=H, bed=np.full_like(H, -10.0), vel=vel,
             H_bc=disc(grid, w, 1.0, R_inner, R_inner), v_bc=v_bc, Q=np.zeros(O.shape(p, p.w_stag, 2)),
             sea=np.zeros_like(H), mask=np.zeros_like(H), surf=np.zeros_like(H), divQ=np.zeros((N, N)),
             dH=np.zeros((N, N)), ce=np.zeros((N, N)), dt=0.0)
    O.lib().orc_geometry_compute(C.byref(p), H.size, O.dptr(S["sea"]), O.dptr(S["bed"]), O.dptr(H), O.dptr(S["mask"]),
                                 O.dptr(S["surf"]))
    return S


def oracle_flow_steps(S, nsteps, after_step=None):
    """nsteps CFL-limited flow steps + ensure_consistency on the oracle (default variant); after_step(S, step) is
    called with the state of each finished step."""
    p, L = S["p"], O.lib()
    out4 = (C.c_double * 4)()
    for step in range(nsteps):
        assert L.orc_cfl_2d(C.byref(p), 1e9, O.dptr(S["mask"]), O.dptr(S["vel"]), out4) == 0
        S["dt"] = out4[0]
        assert L.orc_mass_flow_step(C.byref(p), S["dt"], O.dptr(S["sea"]), O.dptr(S["bed"]), O.dptr(S["H"]),
                                    O.dptr(S["vel"]), O.dptr(S["v_bc"]), O.dptr(S["H_bc"]), O.dptr(S["Q"]),
                                    O.dptr(S["divQ"]), O.dptr(S["dH"]), O.dptr(S["ce"])) == 0
        G.wrap_ghosts(S["H"], S["w"])
        L.orc_geometry_compute(C.byref(p), S["H"].size, O.dptr(S["sea"]), O.dptr(S["bed"]), O.dptr(S["H"]),
                               O.dptr(S["mask"]), O.dptr(S["surf"]))
        if after_step is not None:
            after_step(S, step)


@pytest.mark.parametrize("N", [51, 64])
def test_spreading_disc_flow_steps_bit_exact(N):
    S = spreading_disc_setup(N)
    sia = U.make_sia(S["grid"], S["cfg"])
    for name, a in (("thickness", S["H"]), ("bed", S["bed"]), ("sliding", S["vel"]), ("thk_bc_mask", S["H_bc"]),
                    ("vel_bc_mask", S["v_bc"]), ("flux", S["Q"])):
        sia.upload(name, a)
    sia._check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
    assert np.array_equal(sia.download("mask"), S["mask"])

    def device_step(S, step):
        sia._check(lib.siafd_b200_mass_flow_step(sia.handle, S["dt"]))
        sia._check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
        assert np.array_equal(sia.download("flux_div"), S["divQ"]), step
        assert np.array_equal(sia.download("thk_change"), S["dH"]), step
        assert np.array_equal(sia.download("cons_err"), S["ce"]), step
        assert np.array_equal(sia.download("thickness"), S["H"]), step
        assert np.array_equal(sia.download("mask"), S["mask"]), step

    oracle_flow_steps(S, 12, device_step)
    w = S["w"]
    assert np.abs(cases.interior(S["H"], w) - cases.interior(S["H_bc"], w)).max() > 0.1   # the disc did spread
    assert set(np.unique(S["mask"])) == {3.0, 4.0}                                         # floating ice, open ocean
    if N == 51:  # the frozen vector of tools/make_golden.py
        import os
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_mass_transport_51.npz"))
        assert np.array_equal(sia.download("thickness"), d["thickness"])
        assert np.array_equal(sia.download("mask"), d["mask"])
