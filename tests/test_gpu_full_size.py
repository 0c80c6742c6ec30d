"""BASELINE.json's full-size workload (configs[4]: synthetic dome 4096 x 4096 x 101, gpbld, haseloff) on the GPU,
checked through properties that do not need the oracle to run at that size:
  * windows: the oracle run on 48 x 48 windows cut out of the big domain (with the big domain's own values in
    the windows' ghost cells) must agree with the GPU result in the windows' interior -- the path is a
    width-2 stencil in the map plane, so a window's interior does not know about the rest of the domain;
  * ice-free columns carry exactly the (zero) sliding velocity;
  * D_max equals the maximum of the diffusivity field; the flux is -D * slope bit for bit;
  * the dome is symmetric under x -> -x: u is antisymmetric, v symmetric (to rounding).
Device-resident (torch tensors bound as the handle's field storage); only small slices come to the host."""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle_lib as O
from pism_b200 import capi, grid as G, synthetic as S
from pism_b200.capi import F, lib
from pism_b200.sia import SIAFD

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

M, MZ = 4096, 101


@pytest.fixture(scope="module")
def big():
    L = (M - 1) / 2.0 * 5000.0
    grid = G.Grid(M, M, MZ, L, L, 4000.0)
    cfg = capi.default_config()
    cfg.smoother_range = 0.0
    sia = SIAFD(grid, config=cfg, device=0)
    dev = torch.device("cuda", 0)
    inp = S.dome(grid, grid.whole(), sia.config, device=dev)
    fields = dict(inp)
    for name in ("h_x", "h_y", "D", "flux", "u", "v"):
        fields[name] = torch.zeros(sia.field_shape(name), dtype=torch.float64, device=dev)
    for name, t in fields.items():
        assert lib.siafd_b200_bind(sia.handle, F[name], t.data_ptr()) == 0
    assert lib.siafd_b200_compute_gradient(sia.handle) == 0
    assert lib.siafd_b200_wrap_ghosts_many(sia.handle, 2, (C.c_int * 2)(F["h_x"], F["h_y"])) == 0
    assert lib.siafd_b200_compute_flux_velocity(sia.handle, 1, 0.0) == 0
    assert lib.siafd_b200_wrap_ghosts_many(sia.handle, 2, (C.c_int * 2)(F["u"], F["v"])) == 0
    assert lib.siafd_b200_finish(sia.handle) == 0, lib.siafd_b200_last_error(sia.handle)
    torch.cuda.synchronize()
    yield grid, sia, fields
    del fields


@pytest.mark.parametrize("ci,cj", [(2048, 2048), (2048 + 1500, 2048), (2048 + 1000, 2048 - 1100), (500, 3000), (30, 30)])
def test_windows_match_the_oracle(big, ci, cj):
    grid, sia, f = big
    n, wg = 48, 2
    i0, j0 = ci - n // 2, cj - n // 2
    wgrid = G.Grid(n, n, MZ, (n - 1) / 2.0 * 5000.0, (n - 1) / 2.0 * 5000.0, 4000.0)
    cfg = cases.Cfg(flow_law="gpbld", smoother_range=0.0, D_limit=1e9)

    def cut(name, w_big):  # window [j0-2, j0+n+2) x [i0-2, i0+n+2) of a big local array with ghost width w_big
        t = f[name]
        return np.ascontiguousarray(t[j0 - wg + w_big:j0 + n + wg + w_big, i0 - wg + w_big:i0 + n + wg + w_big].cpu().numpy())

    inputs = {k: cut(k, 2) for k in ("surface", "thickness", "mask", "bed", "enthalpy")}
    run = cases.oracle_run(wgrid, cfg, inputs, None, full=True)
    assert run.status == 0
    # interior of the window: 3 cells away from its edge (the oracle treats the window's edge as the domain's)
    m = 3
    for name, key, w_out in (("D", "D", 1), ("flux", "Q", 1), ("u", "u", 1), ("v", "v", 1)):
        want = cases.interior(run.a[key], w_out)[m:-m, m:-m]
        got = f[name][j0 + m + 1:j0 + n - m + 1, i0 + m + 1:i0 + n - m + 1].cpu().numpy()
        scale = np.max(np.abs(want))
        assert np.max(np.abs(got - want)) <= 1e-10 * max(scale, 1e-300), (name, ci, cj)
    hx = f["h_x"][j0 + m + 1:j0 + n - m + 1, i0 + m + 1:i0 + n - m + 1].cpu().numpy()
    assert np.array_equal(hx, cases.interior(run.a["h_x"], 1)[m:-m, m:-m])


def test_global_properties(big):
    grid, sia, f = big
    D, Q, hx, hy = f["D"], f["flux"], f["h_x"], f["h_y"]
    # D_max = max over owned + ghost staggered points (SIAFD.cc:729; ghosts are wrapped copies)
    assert lib.siafd_b200_max_diffusivity(sia.handle) == float(D.max()) > 0.0
    # flux = -D * slope, bit for bit (SIAFD.cc:783-790)
    assert torch.equal(Q[1:-1, 1:-1, 0], -D[1:-1, 1:-1, 0] * hx[1:-1, 1:-1, 0])
    assert torch.equal(Q[1:-1, 1:-1, 1], -D[1:-1, 1:-1, 1] * hy[1:-1, 1:-1, 1])
    # columns with no ice at any of the four staggered neighbours: u = v = sliding velocity = 0 exactly
    H = f["thickness"][2:-2, 2:-2]
    free = H == 0
    free = free & torch.roll(free, 1, 0) & torch.roll(free, -1, 0) & torch.roll(free, 1, 1) & torch.roll(free, -1, 1)
    u, v = f["u"][1:-1, 1:-1], f["v"][1:-1, 1:-1]
    assert int(free.sum()) > 0.4 * M * M
    assert float(u[free].abs().max()) == 0.0 and float(v[free].abs().max()) == 0.0
    assert bool(torch.isfinite(u).all()) and bool(torch.isfinite(v).all())
    # mirror symmetry of the dome under x -> -x (columns i and Mx-1-i)
    band = slice(1024, 3072, 97)
    ub, vb = u[band], v[band]
    scale = float(ub.abs().max())
    assert scale > 0
    assert float((ub + torch.flip(ub, dims=[1])).abs().max()) <= 1e-9 * scale
    assert float((vb - torch.flip(vb, dims=[1])).abs().max()) <= 1e-9 * max(float(vb.abs().max()), 1e-300)
