"""PISM's processor grid and ownership ranges (src/util/IceGrid.cc:443-586): the default rule, the -Nx / -Ny /
-procs_x / -procs_y overrides with the reference's validation errors, and the load-balanced ranges bench.py passes
as -procs_x / -procs_y."""
import numpy as np
import pytest

from pism_b200 import grid as G


def test_default_rule_and_uniform_ranges():
    assert G.compute_nprocs(4096, 4096, 8) == (2, 4)
    assert G.ownership_ranges(10, 3) == [4, 3, 3]
    ps = G.decompose(4096, 4096, 8)
    assert [(p.xs, p.xm, p.ys, p.ym) for p in ps[:3]] == [(0, 2048, 0, 1024), (2048, 2048, 0, 1024), (0, 2048, 1024, 1024)]
    assert ps[5].rank == 5 and ps[5].neighbor(1, 1) == 6 and ps[7].neighbor(0, 1) == 1


def test_explicit_ownership_ranges_and_their_errors():
    ps = G.decompose(13, 11, 4, Nx=2, Ny=2, procs_x=[9, 4], procs_y=[3, 8])
    assert [(p.xs, p.xm, p.ys, p.ym) for p in ps] == [(0, 9, 0, 3), (9, 4, 0, 3), (0, 9, 3, 8), (9, 4, 3, 8)]
    with pytest.raises(ValueError, match="procs_x don't sum up to Mx"):
        G.decompose(13, 11, 4, Nx=2, Ny=2, procs_x=[9, 5])
    with pytest.raises(ValueError, match="-Ny has to be equal to the -procs_y size"):
        G.decompose(13, 11, 4, Nx=2, Ny=2, procs_y=[3, 4, 4])
    with pytest.raises(ValueError, match="Nx \\* Ny has to be equal to 4"):
        G.decompose(13, 11, 4, Nx=3, Ny=2)


def test_balanced_ranges_equalise_the_cost_of_a_dome():
    M, R = 512, 0.75
    y = (np.arange(M) - (M - 1) / 2) / ((M - 1) / 2)
    X, Y = np.meshgrid(y, y)
    cost = np.where(X * X + Y * Y < R * R, 2.5, 1.0)
    for n in (2, 4, 8, 6):
        Nx, Ny = G.compute_nprocs(M, M, n)
        lx, ly = G.balanced_ownership_ranges(cost, Nx, Ny)
        assert sum(lx) == M and sum(ly) == M and min(lx) >= 2 and min(ly) >= 2
        ps = G.decompose(M, M, n, procs_x=lx, procs_y=ly)
        c = [cost[p.ys:p.ys + p.ym, p.xs:p.xs + p.xm].sum() for p in ps]
        cu = [cost[p.ys:p.ys + p.ym, p.xs:p.xs + p.xm].sum() for p in G.decompose(M, M, n)]
        assert max(c) <= max(cu) + 1e-9
        assert max(c) / np.mean(c) < 1.03
    # 2 x 4 on the dome: the default ranges leave the four central ranks 22 % above the mean
    cu = [cost[p.ys:p.ys + p.ym, p.xs:p.xs + p.xm].sum() for p in G.decompose(M, M, 8)]
    assert max(cu) / np.mean(cu) > 1.2
