"""pismv tests F and G on the B200, one update from the exact state: the COLUMNS u(z) (radial speed), w(z) and
Sigma(z) of the CUDA path -- SIAFD::update, StressBalance::compute_vertical_velocity, ::compute_volumetric_strain_heating
-- against the reference's own exact solution exactFG (src/verification/tests/exactTestsFG.cc compiled UNMODIFIED into
oracle/_ref/libpism_exact.so), at several radii, with the discretisation error bounded and its ORDER checked under
refinement 31 -> 61 -> 121 (second order in dx and dz expected; the horizontal velocity of the SIA is a z-integral of a
gradient, Bueler et al. 2007).  This is what the path alone can say about test G: the golden rows of
test/regression/test_17.sh need 1000 model years of the ENERGY step too (not on this path).
VERDICT r1, item 7 (i)."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import gpu_util as U
import oracle_lib as O
from pism_b200 import grid as G
from pism_b200 import synthetic as S

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(O.REF_EXACT), reason="oracle/_ref/libpism_exact.so not built")]

SECPERA = 31556926.0


def column_errors(M, t_years, Cp, rlo=100e3):
    """Relative max errors of the radial speed, w and Sigma columns over the annulus rlo < r < 600 km (away from the
    dome, where everything vanishes, and from the margin, where the one-sided stencils take over)."""
    grid = G.Grid(M, M, M, 900e3, 900e3, 4000.0)
    cfg = cases.Cfg(flow_law="arr", smoother_range=0.0, fl_e=1.0, dry_simulation=0, **cases.cold_converter())
    inputs = cases.to_numpy(S.test_FG_state(grid, grid.whole(), cfg, t_years=t_years, Cp=Cp))
    sia = U.make_sia(grid, cfg, None)
    U.gpu_update(sia, inputs, True)
    u, v = cases.interior(sia.velocity_u(), 1), cases.interior(sia.velocity_v(), 1)
    w = sia.compute_vertical_velocity()
    sig = sia.compute_volumetric_strain_heating("arr", 3.0, 1.0)
    ref = C.CDLL(O.REF_EXACT)
    ref.ref_exactFG.argtypes = [C.c_double, C.c_double, C.c_int] + [C.POINTER(C.c_double)] + [C.c_double] + \
        [C.POINTER(C.c_double)] * 7
    z = np.ascontiguousarray(grid.z)
    eu = ew = es = 0.0
    su = sw = ss = 0.0
    n = 0
    for j in range(grid.My):
        for i in range(grid.Mx):
            r = float(np.hypot(grid.x[i], grid.y[j]))
            if not (rlo < r < 600e3):
                continue
            H, Mb = C.c_double(), C.c_double()
            outs = [np.zeros(M) for _ in range(5)]
            assert ref.ref_exactFG(t_years * SECPERA, r, M, O.dptr(z), Cp, C.byref(H), C.byref(Mb),
                                   *[O.dptr(o) for o in outs]) == 0
            ks = grid.k_below_height(H.value)
            ur = (u[j, i, :ks + 1] * grid.x[i] + v[j, i, :ks + 1] * grid.y[j]) / r
            eu, su = max(eu, np.abs(ur - outs[1][:ks + 1]).max()), max(su, np.abs(outs[1][:ks + 1]).max())
            ew, sw = max(ew, np.abs(w[j, i, :ks + 1] - outs[2][:ks + 1]).max()), max(sw, np.abs(outs[2][:ks + 1]).max())
            k = slice(1, max(ks - 1, 2))  # (the base level takes a one-sided u_z)
            es, ss = max(es, np.abs(sig[j, i, k] - outs[3][k] * 910.0 * 2009.0).max()), max(ss, np.abs(outs[3][k]).max() * 910.0 * 2009.0)
            n += 1
    assert n > 100
    return eu / su, ew / sw, es / ss, su * SECPERA


@pytest.mark.parametrize("test,t_years,Cp", [("F", 0.0, 0.0), ("G", 500.0, 200.0)])
def test_columns_against_exactFG_with_refinement(test, t_years, Cp):
    # test G's thickness perturbation f(r) g(t) switches on at r = 0.3 L = 225 km with a jump in its SECOND derivative
    # (exactTestsFG.cc:96-102, :165-170): there the max-norm error of any difference scheme drops to first order for u
    # and does not converge for w (which takes H_rr), so G is sampled from 270 km on
    e = {M: column_errors(M, t_years, Cp, 100e3 if Cp == 0.0 else 270e3) for M in (31, 61, 121)}
    for M in (31, 61, 121):
        print("test %s %3d^3: relative max error of u(z) %.3e, w(z) %.3e, Sigma(z) %.3e (max speed %.2f m/a)" %
              ((test, M) + e[M]))
    # stated bounds, relative to the largest exact value in the annulus.  Measured on the B200 at 31^3 / 61^3 / 121^3:
    #   F: u 3.1e-2 / 1.1e-2 / 3.1e-3,  w 1.1e-2 / 2.6e-3 / 1.2e-3,  Sigma 7.7e-2 / 2.0e-2 / 4.9e-3
    #   G: u 8.4e-2 / 2.4e-2 / 6.0e-3,  w 3.3e-1 / 1.1e-1 / 3.3e-2,  Sigma 6.4e-2 / 1.9e-2 / 4.9e-3
    # (G's w carries the second derivative of the bump and starts from a large error on the 60 km grid.  The reference
    # reports errors of this size for its own time-stepped runs: maxUvec 0.95 m/a of ~10 m/a at 31^2 in test_17.sh.)
    bound = {"F": ((0.05, 0.016, 0.0045), (0.02, 0.004, 0.002), (0.1, 0.03, 0.007)),
             "G": ((0.12, 0.035, 0.009), (0.45, 0.16, 0.05), (0.1, 0.03, 0.007))}[test]
    for q in range(3):
        for M, b in zip((31, 61, 121), bound[q]):
            assert e[M][q] < b, (test, q, M, e[M][q], b)
    # order of convergence per halving of dx = dz (max norm over an annulus whose sample points move with the grid, so
    # single steps scatter): every halving at least halves the error, and over the two halvings the mean order is at
    # least 1.5 for u and w and 1.8 for Sigma (measured for F: 1.66, 1.65, 1.98; for G: 1.90, 1.66, 1.86)
    for q, order in ((0, 1.5), (1, 1.5), (2, 1.8)):
        assert e[61][q] < 0.5 * e[31][q] and e[121][q] < 0.5 * e[61][q], (q, e)
        assert 0.5 * np.log2(e[31][q] / e[121][q]) > order, (q, e)
