"""Random geometries through the host-array call's transfer plan (siafd_b200_host_plan_emulate, no GPU): several ice
bodies per row, ice across the periodic edges of the domain, floating tongues and open ocean, bedrock bumps, thin and
thick ice next to each other, random sliding -- the shapes on which a bounding-interval-per-band plan could go wrong.
For every seed and several band / chunk settings the host's u, v (starting as NaN) must equal the oracle's bit for bit,
and the oracle fed with the enthalpy the device would hold (NaN wherever nothing is uploaded) must reproduce the update."""
import numpy as np
import pytest
import torch

import cases
from pism_b200 import grid as G
from pism_b200 import synthetic as S
from test_host_plan import capi_config, emulate

SETTINGS = [(8, 1, 1, 1, 8), (16, 2, 1, 3, 24), (5, 1, 1, 1, 1000), (12, 3, 1, 2, 16), (8, 1, 1, 0, 64), (64, 1, 1, 1, 8)]


def random_case(seed):
    rng = np.random.default_rng(seed)
    Mx, My, Mz = int(rng.integers(20, 49)), int(rng.integers(18, 45)), int(rng.choice([11, 17, 26]))
    spacing = "equal" if seed % 3 else "quadratic"
    grid = G.Grid(Mx, My, Mz, 0.5 * (Mx - 1) * 5e3, 0.5 * (My - 1) * 5e3, 4000.0, spacing=spacing)
    law = ["gpbld", "pb", "isothermal_glen", "hooke"][seed % 4]
    kw = dict(flow_law=law, smoother_range=0.0, D_limit=1.0e12, dry_simulation=int(seed % 5 == 0))
    if law == "isothermal_glen":
        kw["iso_softness_A"] = 1.0e-16 / cases.SECPERA_UDUNITS
    cfg = cases.Cfg(**kw)
    # owned fields first (periodic wrap makes the ghosts): a few blobs of ice, some of them across the edges
    jj, ii = np.meshgrid(np.arange(My), np.arange(Mx), indexing="ij")
    H = np.zeros((My, Mx))
    for _ in range(int(rng.integers(1, 5))):
        ci, cj = rng.uniform(0, Mx), rng.uniform(0, My)
        ri, rj = rng.uniform(2, 0.45 * Mx), rng.uniform(2, 0.45 * My)
        di = np.minimum(np.abs(ii - ci), Mx - np.abs(ii - ci)) / ri  # periodic distance
        dj = np.minimum(np.abs(jj - cj), My - np.abs(jj - cj)) / rj
        rho = np.sqrt(di * di + dj * dj)
        H = np.maximum(H, np.where(rho < 1.0, rng.uniform(300.0, 3900.0) * np.clip(1.0 - rho ** 1.5, 0.0, None) ** 0.4, 0.0))
    H[rng.random(H.shape) < 0.02] = 0.0                       # holes
    H = np.where(rng.random(H.shape) < 0.01, 25.0, H)          # lone thin columns
    bed = -600.0 + 900.0 * np.sin(2 * np.pi * ii / Mx + rng.uniform(0, 6)) * np.cos(2 * np.pi * jj / My + rng.uniform(0, 6)) \
        + 150.0 * rng.standard_normal(H.shape)
    bed = np.minimum(bed, 3950.0 - H)                          # the surface stays inside the grid
    sea = np.zeros_like(H)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    mask, surface = S.geometry_calculator(cfg, t(sea), t(bed), t(H))
    mask, surface = mask.numpy(), surface.numpy()
    # enthalpy: cold ice with a warm base here and there
    ec = S.ec_constants(cfg)
    z = t(grid.z)[None, None, :]
    depth = t(H)[..., None] - z
    P = S.pressure(ec, depth)
    T = t(245.0 + 20.0 * rng.random(H.shape))[..., None] + 0.004 * torch.clamp(depth, min=0.0)
    omega = torch.where((depth > 0.9 * t(H)[..., None]) & (t(rng.random(H.shape))[..., None] < 0.3), 0.004, 0.0)
    T = torch.where(omega > 0, ec["T_melting"] - ec["beta"] * P + 1.0, torch.minimum(T, ec["T_melting"] - ec["beta"] * P - 0.5))
    E = S.enthalpy_permissive(ec, T, omega.to(torch.float64), P).numpy()
    sliding = 1e-7 * rng.standard_normal((My, Mx, 2)) * (rng.random((My, Mx, 1)) < 0.7)
    wrap = lambda a, w: np.pad(a, ((w, w), (w, w)) + ((0, 0),) * (a.ndim - 2), mode="wrap")
    inputs = dict(thickness=wrap(H, cfg.w_geom), bed=wrap(bed, cfg.w_geom), mask=wrap(mask, cfg.w_geom),
                  surface=wrap(surface, cfg.w_geom), enthalpy=wrap(E, cfg.w_3d_in), sliding=wrap(sliding, cfg.w_sliding))
    return grid, cfg, {k: np.ascontiguousarray(v) for k, v in inputs.items()}


@pytest.mark.parametrize("seed", range(24))
def test_random_geometry(seed):
    grid, cfg, inputs = random_case(seed)
    one = cases.oracle_run(grid, cfg, inputs)
    assert one.status == 0, one.status
    c = capi_config(grid, cfg)
    cut_bytes, full_bytes = None, None
    for rows, band, sparse, cut, cols in SETTINGS:
        E_dev, u, v, up, dn = emulate(c, inputs, one.a["u"], one.a["v"], rows, band, sparse, cut, cols, False)
        assert np.array_equal(u, one.a["u"]) and np.array_equal(v, one.a["v"]), (seed, rows, band, cut, cols)
        again = cases.oracle_run(grid, cfg, dict(inputs, enthalpy=E_dev))
        assert again.status == 0 and again.D_max == one.D_max, (seed, rows, band, cut, cols)
        for k in ("D", "Q", "u", "v"):
            assert np.array_equal(again.a[k], one.a[k]), (seed, k, rows, band, cut, cols)
        if (rows, band) == (8, 1):
            if cut:
                cut_bytes = (up, dn)
            else:
                full_bytes = (up, dn)
    assert cut_bytes[0] <= full_bytes[0] and cut_bytes[1] <= full_bytes[1]


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_geometry_on_patches(seed):
    """The same on the patches of a decomposed domain (what every rank of a multi-GPU run plans for its own patch):
    rectangles clipped at the patch edges, ghost columns from the neighbours, unequal ownership ranges."""
    import oracle_lib as O
    grid, cfg, inputs = random_case(seed)
    rng = np.random.default_rng(seed)
    one = cases.oracle_run(grid, cfg, inputs)
    assert one.status == 0
    Nx, Ny = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    if Nx * Ny == 1:
        Nx = 2

    def ranges(M, n):  # unequal ownership ranges, every one at least 4 wide
        cuts = np.sort(rng.choice(np.arange(1, M // 4), size=n - 1, replace=False)) * 4 if n > 1 else np.array([], dtype=int)
        edges = np.concatenate(([0], cuts, [M]))
        return [int(b - a) for a, b in zip(edges[:-1], edges[1:])]

    patches = G.decompose(grid.Mx, grid.My, Nx * Ny, Nx=Nx, Ny=Ny, procs_x=ranges(grid.Mx, Nx), procs_y=ranges(grid.My, Ny))
    glob = {k: np.ascontiguousarray(cases.interior(np.asarray(v), (v.shape[0] - grid.My) // 2)) for k, v in inputs.items()}
    widths = dict(enthalpy=cfg.w_3d_in, sliding=cfg.w_sliding)
    for rows, band, cut, cols in ((8, 1, 1, 8), (16, 2, 2, 16), (8, 1, 0, 64)):
        runs = []
        for pt in patches:
            c = capi_config(grid, cfg, pt)
            loc = {k: G.global_to_local(glob[k], pt, widths.get(k, cfg.w_geom)) for k in glob}
            want = {k: G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1) for k in ("u", "v")}
            E_dev, u, v, up, dn = emulate(c, loc, want["u"], want["v"], rows, band, 1, cut, cols, True)
            assert np.array_equal(u, want["u"]) and np.array_equal(v, want["v"]), (seed, pt, rows, band, cut, cols)
            runs.append(O.Run(cfg.oracle_params(grid, pt), dict(loc, enthalpy=E_dev)))
        size = len(patches)
        P = (O.Params * size)(*[r.p for r in runs])
        Fa = (O.Fields * size)(*[r.f for r in runs])
        assert O.lib().orc_siafd_update_decomposed(size, P, Fa, 1, 4) == 0
        for q, (r, pt) in enumerate(zip(runs, patches)):
            assert Fa[q].D_max == one.D_max
            for k in ("u", "v"):
                assert np.array_equal(r.a[k], G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1)), (seed, k, q)
            for k in ("D", "Q"):
                want = G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1)
                assert np.array_equal(cases.interior(r.a[k], 1), cases.interior(want, 1)), (seed, k, q)
