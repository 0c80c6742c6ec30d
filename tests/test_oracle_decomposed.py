"""The oracle's SIAFD::update on the patches of one domain (orc_siafd_update_decomposed: the reference's passes with its
two ghost updates as copies from the owning patch, one OpenMP thread per rank) must reproduce the single-patch oracle
bit for bit -- owned points and ghosts -- for PISM's default decomposition and for unequal ranges
(test/regression/test_02.sh).  bench.py's CPU arm times exactly this call."""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle_lib as O
from pism_b200 import grid as G


@pytest.mark.parametrize("name,decomp", [
    ("dome_64_21", dict(size=8)),
    ("C4s_nosmooth", dict(size=6, Nx=2, Ny=3, procs_x=[40, 21], procs_y=[50, 13, 50])),
    ("C1_mahaffy", dict(size=4)),
    ("dome_35_101", dict(size=3, Nx=3, Ny=1, procs_x=[17, 3, 15])),
])
def test_patches_equal_single_patch(name, decomp):
    grid, cfg, inputs, gb = cases.case(name)
    one = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert one.status == 0
    size = decomp.pop("size")
    patches = G.decompose(grid.Mx, grid.My, size, **decomp)
    glob = {k: np.ascontiguousarray(cases.interior(np.asarray(v), (v.shape[0] - grid.My) // 2)) for k, v in inputs.items()
            if k in ("surface", "thickness", "mask", "bed", "enthalpy")}
    runs = []
    for pt in patches:
        p = cfg.oracle_params(grid, pt)
        loc = {k: G.global_to_local(glob[k], pt, p.w_3d_in if k == "enthalpy" else p.w_geom) for k in glob}
        runs.append(O.Run(p, loc))
    P = (O.Params * size)(*[r.p for r in runs])
    Fa = (O.Fields * size)(*[r.f for r in runs])
    assert O.lib().orc_siafd_update_decomposed(size, P, Fa, 1, 4) == 0
    for q, (r, pt) in enumerate(zip(runs, patches)):
        assert Fa[q].D_max == one.D_max
        for k, w in (("h_x", 1), ("h_y", 1), ("u", 1), ("v", 1)):  # ghosts came from the neighbours
            want = G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], w)), pt, w)
            assert np.array_equal(r.a[k], want), (k, q)
        for k in ("D", "Q"):  # computed locally on owned + 1: the owned points must agree
            want = G.global_to_local(np.ascontiguousarray(cases.interior(one.a[k], 1)), pt, 1)
            assert np.array_equal(cases.interior(r.a[k], 1), cases.interior(want, 1)), (k, q)
