"""The host side of siafd_b200_update with host arrays, without a GPU: siafd_b200_host_plan_emulate executes the
call's transfer plan (pism_b200/csrc/siafd_hostplan.hh: row bands, rectangles of columns near ice, level cut, host
fills, replication above the cut, ghost rows / columns) with memcpy, the oracle standing in for the device.

* the host's u, v must come out bit-identical to the full result -- from arrays that start as NaN, so every cell has
  to be written by a copy, a fill or a replication;
* the enthalpy the device would hold (NaN wherever nothing is uploaded) must give the same update as the full one."""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle_lib as O
from pism_b200 import capi, grid as G


def capi_config(grid, cfg, patch=None):
    patch = patch or grid.whole()
    c = capi.default_config()
    for k, v in cfg.overrides().items():
        setattr(c, k, v)
    c.Mx, c.My, c.Mz = grid.Mx, grid.My, grid.Mz
    c.xs, c.xm, c.ys, c.ym = patch.xs, patch.xm, patch.ys, patch.ym
    c.dx, c.dy = grid.dx, grid.dy
    c._z_keep = np.ascontiguousarray(grid.z, dtype=np.float64)
    c.z = c._z_keep.ctypes.data_as(C.POINTER(C.c_double))
    return c


def emulate(c, inputs, u_dev, v_dev, rows, band, sparse, cut, cut_cols, patch):
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    a = {k: f64(inputs[k]) for k in ("thickness", "surface", "bed", "mask", "sliding", "enthalpy")}
    E_dev = np.full_like(a["enthalpy"], np.nan)
    u, v = np.full_like(u_dev, np.nan), np.full_like(v_dev, np.nan)
    up, dn = C.c_int64(0), C.c_int64(0)
    u_dev, v_dev = f64(u_dev), f64(v_dev)
    st = capi.lib.siafd_b200_host_plan_emulate(
        C.byref(c), rows, band, int(sparse), int(cut), cut_cols, int(patch), a["thickness"].ctypes.data,
        a["surface"].ctypes.data, a["bed"].ctypes.data, a["mask"].ctypes.data, a["sliding"].ctypes.data,
        a["enthalpy"].ctypes.data, E_dev.ctypes.data, u_dev.ctypes.data, v_dev.ctypes.data, u.ctypes.data, v.ctypes.data,
        C.byref(up), C.byref(dn))
    assert st == capi.OK
    return E_dev, u, v, up.value, dn.value


SETTINGS = [  # rows per segment, segments per band, sparse, level cut (1 + k: k rows per chunk), columns per chunk
    (16, 1, 1, 1, 128), (16, 1, 1, 1, 8), (8, 2, 1, 1, 16), (32, 1, 1, 1, 24), (64, 100, 1, 1, 9), (5, 3, 1, 1, 40),
    (16, 1, 1, 0, 128), (8, 3, 0, 0, 128), (88, 1, 1, 1, 32),
    (16, 1, 1, 2, 16), (32, 2, 1, 5, 64),  # level cut 1 + k: chunks of k rows (k = 1: plain 2D copies)
]


@pytest.mark.parametrize("name", ["Fs", "dome_96_31_rough", "dome_64_31_quadratic", "C4s_nosmooth", "C1_31", "dome_40_21_big",
                                  "dome_160_11_rough"])  # (160 rows: the scans of the plan run on two threads)
def test_whole_domain_plan_leaves_the_full_result_on_the_host(name):
    big = name.endswith("_big")  # ice up to the edge of the domain: whole rows, ghost columns included
    grid, cfg, inputs, gb = cases.case(name[:-4] if big else name)
    if big:
        from pism_b200 import synthetic as S
        inputs = cases.to_numpy(S.dome(grid, grid.whole(), cfg, variant="rough", Rfrac=1.6))
    rng = np.random.default_rng(11)
    inputs["sliding"] = 1e-7 * rng.standard_normal(inputs["sliding"].shape)
    w = cfg.w_sliding  # periodic ghosts, like every other input
    core = inputs["sliding"][w:-w, w:-w]
    inputs["sliding"] = np.pad(core, ((w, w), (w, w), (0, 0)), mode="wrap")
    one = cases.oracle_run(grid, cfg, inputs, gb)
    assert one.status == 0
    c = capi_config(grid, cfg)
    sizes = {}
    for rows, band, sparse, cut, cols in SETTINGS:
        E_dev, u, v, up, dn = emulate(c, inputs, one.a["u"], one.a["v"], rows, band, sparse, cut, cols, False)
        assert np.array_equal(u, one.a["u"]) and np.array_equal(v, one.a["v"]), (rows, band, sparse, cut, cols)
        sizes[(rows, band, sparse, cut, cols)] = (up, dn)
        if not sparse:
            assert np.array_equal(E_dev, inputs["enthalpy"])
            assert up == inputs["enthalpy"].nbytes and dn == 2 * u.nbytes
            continue
        again = cases.oracle_run(grid, cfg, dict(inputs, enthalpy=E_dev), gb)
        assert again.status == 0 and again.D_max == one.D_max
        for k in ("D", "Q", "u", "v"):
            assert np.array_equal(again.a[k], one.a[k]), (k, rows, band, cut, cols)
    dense, sparse = sizes[(8, 3, 0, 0, 128)], sizes[(16, 1, 1, 0, 128)]
    wide, narrow = sizes[(16, 1, 1, 1, 128)], sizes[(16, 1, 1, 1, 8)]
    assert sparse[0] <= dense[0] and sparse[1] <= dense[1]
    assert narrow[0] <= wide[0] <= sparse[0] and narrow[1] <= wide[1] <= sparse[1]
    best = [min(v[q] for k, v in sizes.items() if k[3]) for q in (0, 1)]
    assert best[0] < sparse[0] and best[1] < sparse[1]  # the cut does cut


@pytest.mark.parametrize("name,decomp", [
    ("dome_64_21", dict(size=8)),
    ("C4s_nosmooth", dict(size=6, Nx=2, Ny=3, procs_x=[40, 21], procs_y=[50, 13, 50])),
    ("dome_96_31_rough", dict(size=4)),
])
def test_patch_plans_leave_the_full_result_on_every_rank(name, decomp):
    grid, cfg, inputs, gb = cases.case(name)
    one = cases.oracle_run(grid, cfg, inputs, gb)
    assert one.status == 0
    size = decomp.pop("size")
    patches = G.decompose(grid.Mx, grid.My, size, **decomp)
    glob = {k: np.ascontiguousarray(cases.interior(np.asarray(v), (v.shape[0] - grid.My) // 2)) for k, v in inputs.items()
            if k in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding")}
    widths = dict(enthalpy=cfg.w_3d_in, sliding=cfg.w_sliding)
    for rows, band, cut, cols in ((16, 1, 1, 16), (8, 2, 1, 128), (16, 1, 0, 128)):
        runs = []
        for pt in patches:
            c = capi_config(grid, cfg, pt)
            loc = {k: G.global_to_local(glob[k], pt, widths.get(k, cfg.w_geom)) for k in glob}
            want = {k: G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1) for k in ("u", "v")}
            E_dev, u, v, up, dn = emulate(c, loc, want["u"], want["v"], rows, band, 1, cut, cols, True)
            assert np.array_equal(u, want["u"]) and np.array_equal(v, want["v"]), (pt, rows, band, cut, cols)
            runs.append(O.Run(cfg.oracle_params(grid, pt), dict(loc, enthalpy=E_dev)))
        # the decomposed oracle on what the devices would hold
        P = (O.Params * size)(*[r.p for r in runs])
        Fa = (O.Fields * size)(*[r.f for r in runs])
        assert O.lib().orc_siafd_update_decomposed(size, P, Fa, 1, 4) == 0
        for q, (r, pt) in enumerate(zip(runs, patches)):
            assert Fa[q].D_max == one.D_max
            for k in ("u", "v"):
                assert np.array_equal(r.a[k], G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1)), (k, q)
            for k in ("D", "Q"):
                want = G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1)
                assert np.array_equal(cases.interior(r.a[k], 1), cases.interior(want, 1)), (k, q)
