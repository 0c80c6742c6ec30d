"""The two facts the level cut of the sparse host path rests on (include/siafd_b200.h,
siafd_b200_host_levels_needed; pism_b200/csrc/siafd_capi.cu), checked on the CPU restatement of the reference:

* u and v of a column are constant from level n - 1 up (SIAFD.cc:857-859, :935-942), and
* the update never reads the enthalpy of a column on the levels >= n (SIAFD.cc:613-627, :676-689),

where n = siafd_b200_host_levels_needed(z, Mz, T) and T is the largest thk_smooth bound (0 without ice, max(H, usurf -
topg) where grounded, H where floating: BedSmoother.cc:306-320, smoother off) over the column and the eight around it.  The library cuts chunks of columns at the largest n of the chunk, which is implied."""
import ctypes as C

import numpy as np
import pytest

import cases
from pism_b200 import capi

CASES = ["Fs", "dome_96_31_rough", "dome_64_31_quadratic", "C4s_nosmooth", "C1_31"]


def levels_needed(z, T):
    z = np.ascontiguousarray(z, dtype=np.float64)
    return capi.lib.siafd_b200_host_levels_needed(z.ctypes.data_as(C.POINTER(C.c_double)), len(z), float(T))


def neighbourhood_max(a):
    """max over the 3 x 3 neighbourhood, clamped at the edges of the (ghosted) array"""
    p = np.pad(a, 1, mode="edge")
    out = a.copy()
    for dj in range(3):
        for di in range(3):
            out = np.maximum(out, p[dj:dj + a.shape[0], di:di + a.shape[1]])
    return out


def level_counts(grid, inputs, w_from, w_to):
    """n per column of an array with ghost width w_to, from geometry arrays with ghost width w_from >= w_to"""
    H, grounded = inputs["thickness"], np.floor(inputs["mask"] + 0.5) < 3  # Mask.hh:37-66
    T = neighbourhood_max(np.where((H != 0.0) & grounded, np.maximum(H, inputs["surface"] - inputs["bed"]), H))
    d = w_from - w_to
    T = T[d:T.shape[0] - d, d:T.shape[1] - d] if d else T
    return np.array([[levels_needed(grid.z, t) for t in row] for row in T])


def test_levels_needed_follows_kBelowHeight():
    z = np.linspace(0.0, 4000.0, 101)
    Mz = len(z)
    for thk, ks in ((0.0, 0), (39.9, 0), (40.0, 1), (1000.0, 25), (1019.0, 25), (3500.0, 87)):
        # kBelowHeight(thk) = ks (IceGrid.cc:427-440): levels 0 .. ks are read, one level of slack on top
        assert levels_needed(z, thk) == max(ks + 2, 2), (thk, ks)
    assert levels_needed(z, 3880.0) == Mz  # a cut of two or three levels is not taken
    assert levels_needed(z, 4000.0) == Mz and levels_needed(z, 1e9) == Mz
    assert levels_needed(z, float("nan")) == Mz and levels_needed(z, float("inf")) == Mz
    assert levels_needed(z, -5.0) == 2
    zq = 4000.0 * np.linspace(0.0, 1.0, 31) ** 2  # unequal spacing
    for thk in (0.0, 1.0, 17.0, 444.4, 444.5, 2000.0, 3300.0):
        ks = int(np.searchsorted(zq, thk, side="right")) - 1
        n = levels_needed(zq, thk)
        assert n == (len(zq) if ks + 2 >= len(zq) - 2 else ks + 2), (thk, ks, n)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("sliding", [False, True])
def test_u_v_are_constant_above_the_cut_and_enthalpy_is_not_read_there(name, sliding):
    grid, cfg, inputs, gb = cases.case(name)
    assert cfg.smoother_range == 0.0  # the cut is only taken with the bed smoother off
    if sliding:
        rng = np.random.default_rng(5)
        inputs["sliding"] = 1e-7 * rng.standard_normal(inputs["sliding"].shape)
    run = cases.oracle_run(grid, cfg, inputs, gb)
    assert run.status == 0
    u, v = run.a["u"], run.a["v"]
    n_uv = level_counts(grid, inputs, cfg.w_geom, cfg.w_uv)
    assert n_uv.shape == u.shape[:2]
    assert (n_uv < grid.Mz).mean() > 0.3, "the case does not exercise the cut"
    for a in (u, v):
        top = np.take_along_axis(a, (n_uv - 1)[..., None], axis=2)
        above = np.arange(grid.Mz)[None, None, :] >= (n_uv - 1)[..., None]
        assert np.array_equal(np.where(above, a, top), np.broadcast_to(top, a.shape))
    # poison the enthalpy above the cut: nothing of the result may change
    n_E = level_counts(grid, inputs, cfg.w_geom, cfg.w_3d_in)
    E = inputs["enthalpy"].copy()
    E[np.arange(grid.Mz)[None, None, :] >= n_E[..., None]] = np.nan
    poisoned = dict(inputs, enthalpy=E)
    run2 = cases.oracle_run(grid, cfg, poisoned, gb)
    assert run2.status == 0
    for k in ("h_x", "h_y", "D", "Q", "u", "v"):
        assert not np.isnan(run2.a[k]).any(), k
        assert np.array_equal(run2.a[k], run.a[k]), k
    assert run2.D_max == run.D_max
