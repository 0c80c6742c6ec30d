"""Pins the oracle against the reference's OWN golden output: `pismv -test C` (test/regression/test_15.sh), `pismv -test L`
(test/regression/test_16.sh: non-flat bed) and the mass-conservation criterion of `pismv -test B`
(test/regression/test_12.sh).

The run goes through every function of the hot path -- haseloff gradient, compute_diffusivity on pismv's quadratic
levels, diffusive flux, D_max, 3D velocities (through the 3D CFL restriction) -- plus the mass-continuity consumer
(SURVEY.md 8(f) N1), time-stepped with the reference's step logic (pism_b200/icemodel.py).  The four numbers the
reference prints with 6 decimals (9 significant digits) are reproduced digit for digit on both grids."""
import math

import numpy as np

import pismv_oracle as P


def test_pismv_test_C_golden_rows_of_test_15():
    for M, golden in P.TEST_15_GOLDEN.items():
        m = P.pismv_model("C", M)
        m.run()
        assert m.steps == 84  # 83 steps of 60 years ("max") and the remainder ("end of the run")
        assert m.report() == golden, (M, m.report(), golden)


def test_pismv_test_L_golden_rows_of_test_16():
    """test/regression/test_16.sh, "isothermal SIA with non-flat bed": 1000 years from the steady state of exactL on
    the bed b(r) = -500 cos(1.2 pi r / L), Mx = My = 21 and 31.  The bed enters the path through the surface elevation
    (haseloff gradient), through thk_smooth = max(usurf - topgsmooth, 0) (BedSmoother.cc:306-320, smoother off) and
    through the mask; the steps are limited by the diffusivity (timestepping.cc:52-68), so D_max is pinned too.  The
    four printed numbers come out digit for digit.  (exactL needs an ODE solve; the reference's uses GSL, this one
    scipy: pism_b200/verification.py::exactL.)"""
    for M, golden in P.TEST_16_GOLDEN.items():
        m = P.pismv_model("L", M, run_length_years=1000.0)
        m.run()
        assert m.report() == golden, (M, m.report(), golden)
        assert "diffusivity" in m.m_adaptive_timestep_reason or "end of the run" in m.m_adaptive_timestep_reason
        assert m.steps > 1000.0 / 60.0 + 1  # i.e. not the 60-year steps of test C


def test_pismv_test_C_restrictions():
    """The first step has D_max = 0 (no ice): `max time step`; later ones are limited by 60 years (timestepping.cc:52-68,
    :154-164); the list of dt sums to the run length exactly (Time::step snaps to the end, Time.cc:206-215)."""
    m = P.pismv_model("C", 31, run_length_years=500.0)
    m.run()
    assert m.time.current() == m.time.end()
    assert abs(sum(m.dt_history) - m.time.end()) < 1e-3
    assert max(m.dt_history) == m.max_dt


def test_pismv_test_B_conserves_volume_like_test_12():
    """test_12.sh: `pismv -test B -Mx 31 -My 31 -Mz 31 -ys 1000 -y 5000 -max_dt 25`; the ice volume sampled every 25
    years may not grow by more than 1e-14 of its magnitude (no surface mass balance, mass-conserving flow step)."""
    m = P.pismv_model("B", 31, start_year=1000.0, run_length_years=5000.0, max_dt_years=25.0)
    m.backend.ensure_consistency()
    area = m.grid.dx * m.grid.dy
    vol = [math.fsum(m.backend.thickness().ravel()) * area]
    while m.time.current() < m.time.end():
        m.step()
        vol.append(math.fsum(m.backend.thickness().ravel()) * area)
    vol = np.array(vol)
    threshold = 10 ** (np.floor(np.log10(vol.max())) - 14)  # "14 digits of accuracy", test_12.sh:36
    assert np.diff(vol).max() < threshold
    assert m.steps >= 200
    # and the solution stays close to Halfar's: the reference's own error norms for this grid are O(100 m) at the margin
    prcntVOL, maxH, avH, relmaxETA = m.geometry_errors()
    assert prcntVOL < 1.0 and avH < 30.0
