"""Oracle-only checks (CPU): exact solutions, decomposition independence, semantics gotchas."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import oracle_lib as O
from pism_b200 import grid as G
from pism_b200 import verification as V


@pytest.mark.skipif(not os.path.exists(O.REF_EXACT), reason="oracle/_ref not built (needs /root/reference)")
def test_numpy_exact_solutions_match_reference_build():
    """pism_b200.verification restates exactC / exactFG; check against the reference's own sources
    compiled unmodified (oracle/Makefile target `ref`)."""
    ref = C.CDLL(O.REF_EXACT)
    pd = C.POINTER(C.c_double)
    ref.ref_exactC.argtypes = [C.c_double, C.c_double, pd, pd]
    ref.ref_exactFG.argtypes = [C.c_double, C.c_double, C.c_int, pd, C.c_double] + [pd] * 7
    H, M = C.c_double(), C.c_double()
    for t in (15208.0 * V.SperA, 20000.0 * V.SperA, 0.05 * V.SperA):
        for r in np.linspace(0.0, 900e3, 61):
            ref.ref_exactC(t, r, C.byref(H), C.byref(M))
            h, m = V.exactC(t, np.array([r]))
            assert abs(h[0] - H.value) <= 1e-12 * max(1.0, H.value)
            assert abs(m[0] - M.value) <= 1e-12 * max(1e-12, abs(M.value))
    z = np.linspace(0.0, 4000.0, 31)
    for t, Cp in ((0.0, 0.0), (500 * V.SperA, 200.0)):
        for r in np.linspace(1.0, 749.9e3, 41):
            out = [np.zeros(31) for _ in range(5)]
            assert ref.ref_exactFG(t, r, 31, z.ctypes.data_as(pd), Cp, C.byref(H), C.byref(M),
                                   *[a.ctypes.data_as(pd) for a in out]) == 0
            e = V.exactFG(t, r, z, Cp)
            assert abs(e["H"] - H.value) <= 1e-13 * H.value
            assert np.max(np.abs(e["T"] - out[0])) <= 1e-12
            assert np.max(np.abs(e["U"] - out[1])) <= 1e-13 * np.max(np.abs(out[1])) + 1e-30
    # spot values recorded in SURVEY.md section 8(c)
    assert abs(V.exactC(15208.0 * V.SperA, np.array([300e3]))[0][0] - 3099.659127) < 1e-6
    assert abs(V.exactFG(0.0, 300e3, z, 0.0)["H"] - 2503.108496) < 1e-6


def test_test_C_diffusivity_against_exact_solution():
    """Isothermal SIA: D = Gamma H^(n+2) |grad h|^(n-1), Gamma = 2 A (rho g)^n / (n+2)
    (Bueler et al. 2005; exactTestsABCD.c constants).  On Test C's exact thickness the computed
    staggered D must converge to it away from the margin (second order in dx and dz)."""
    errs = []
    for size, Mz in ((31, 31), (61, 61)):
        grid = G.Grid(size, size, Mz, 1000e3, 1000e3, 4000.0)
        cfg = cases.Cfg(flow_law="isothermal_glen", iso_softness_A=1.0e-16 / cases.SECPERA_UDUNITS,
                        smoother_range=0.0, dry_simulation=1, gradient_method="mahaffy", **cases.cold_converter())
        from pism_b200 import synthetic as S
        inputs = cases.to_numpy(S.test_C_state(grid, grid.whole(), cfg))
        run = cases.oracle_run(grid, cfg, inputs, full=False)
        assert run.status == 0
        D = cases.interior(run.a["D"], 1)[:, :, 0]  # i-offset points
        t0 = 15208.0 * V.SperA
        xs = grid.x + 0.5 * grid.dx
        X, Y = np.meshgrid(xs, grid.y)
        r = np.sqrt(X * X + Y * Y)
        H, _ = V.exactC(t0, r)
        eps = 1.0
        Hp, _ = V.exactC(t0, r + eps)
        Hm, _ = V.exactC(t0, np.maximum(r - eps, 0.0))
        slope = np.abs(Hp - Hm) / (2 * eps)
        Gamma = 2.0 * cfg.iso_softness_A * (910.0 * 9.81) ** 3 / 5.0
        Dex = Gamma * H ** 5 * slope ** 2
        sel = (r > 100e3) & (r < 550e3) & (np.arange(size)[None, :] < size - 1)
        errs.append(np.max(np.abs(D[sel] - Dex[sel]) / Dex[sel]))
    assert errs[0] < 0.08 and errs[1] < 0.04 and errs[1] < 0.6 * errs[0], errs


def test_test_F_surface_velocity_against_exact_solution():
    """Shape of siafd_test.cc:105-151: one full update on the Test F state, surface speed vs exactFG."""
    grid, cfg, inputs, _ = cases.case("Fs")
    run = cases.oracle_run(grid, cfg, inputs, full=True)
    assert run.status == 0
    u, v = cases.interior(run.a["u"], 1), cases.interior(run.a["v"], 1)
    w = cfg.w_geom
    H = cases.interior(inputs["thickness"], w)
    Uex = cases.interior(inputs["exact_surface_speed"], w)
    r = cases.interior(inputs["radius"], w)
    maxerr, n = 0.0, 0
    for j in range(grid.My):
        for i in range(grid.Mx):
            if 1.0 <= r[j, i] <= 750000.0 - 1.0 and H[j, i] > 0:
                # IceModelVec3::getValZ: linear interpolation in z (util/iceModelVec3.cc)
                k = grid.k_below_height(H[j, i])
                lam = (H[j, i] - grid.z[k]) / (grid.z[k + 1] - grid.z[k])
                us = u[j, i, k] + lam * (u[j, i, k + 1] - u[j, i, k])
                vs = v[j, i, k] + lam * (v[j, i, k + 1] - v[j, i, k])
                uex, vex = grid.x[i] / r[j, i] * Uex[j, i], grid.y[j] / r[j, i] * Uex[j, i]
                maxerr = max(maxerr, np.hypot(us - uex, vs - vex))
                n += 1
    secpera = V.SperA
    # the 31^2 Test G golden row of test_17.sh has maxUvec 0.945 m/a after 1000 model years; a single
    # update from the exact state must be in that range, and tiny next to the ~10 m/a speeds
    print("Test F 31x31x31: max |U_surface error| = %.4f m/a, max exact speed = %.4f m/a" %
          (maxerr * secpera, np.max(Uex) * secpera))
    assert n > 300 and maxerr * secpera < 1.0, maxerr * secpera
    assert np.max(Uex) * secpera > 2.0 and maxerr < 0.2 * np.max(Uex)


@pytest.mark.parametrize("nranks", [2, 3, 4, 6])
@pytest.mark.parametrize("name", ["C2t", "C4s"])
def test_decomposition_independence(name, nranks):
    """test/regression/test_02.sh: results must not depend on the number of ranks, bit for bit.
    Patches get their ghosts by copying from neighbours exactly where the reference communicates
    (SIAFD.cc:498-499 and :946-947)."""
    grid, cfg, inputs, gb = cases.case(name)
    whole = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert whole.status == 0
    patches = G.decompose(grid.Mx, grid.My, nranks)
    p0 = cfg.oracle_params(grid)
    sm = O.preprocess_bed(p0, gb) if cfg.smoother_range > 0 else None
    runs = []
    for pt in patches:
        loc = {k: G.global_to_local(cases.interior(inputs[k], w), pt, w)
               for k, w in (("surface", cfg.w_geom), ("thickness", cfg.w_geom), ("mask", cfg.w_geom),
                            ("bed", cfg.w_geom), ("enthalpy", cfg.w_3d_in), ("sliding", cfg.w_sliding))}
        smoothed = None
        if sm is not None:
            smoothed = {k: G.global_to_local(sm[k], pt, cfg.w_geom) for k in ("topgsmooth", "maxtl", "C2", "C3", "C4")}
            smoothed["active"] = sm["active"]
        runs.append(O.Run(cfg.oracle_params(grid, pt), loc, smoothed))

    def exchange(field, w):
        """Ghost update: assemble the global owned values, then re-cut every patch with ghosts."""
        shape = runs[0].a[field].shape[2:]
        g = np.zeros((grid.My, grid.Mx) + shape)
        for pt, r in zip(patches, runs):
            g[pt.ys:pt.ys + pt.ym, pt.xs:pt.xs + pt.xm] = cases.interior(r.a[field], w)
        for pt, r in zip(patches, runs):
            r.a[field][...] = G.global_to_local(g, pt, w)
        return g

    for r in runs:
        assert r.gradient() == 0
    if cfg.gradient_method == O.GRADIENTS["haseloff"]:
        exchange("h_x", 1)
        exchange("h_y", 1)
    for r in runs:
        assert r.flux_velocity(True) == 0
    for field, w in (("u", 1), ("v", 1), ("D", 1), ("Q", 1), ("h_x", 1), ("h_y", 1)):
        g = np.zeros((grid.My, grid.Mx) + runs[0].a[field].shape[2:])
        for pt, r in zip(patches, runs):
            g[pt.ys:pt.ys + pt.ym, pt.xs:pt.xs + pt.xm] = cases.interior(r.a[field], w)
        assert np.array_equal(g, cases.interior(whole.a[field], w)), field
    assert max(r.D_max for r in runs) == whole.D_max


def test_semantics_gotchas():
    """SURVEY.md 8(a'): G2 edge override for both offsets, G3 delta unaffected, G9/G10."""
    grid, cfg, inputs, _ = cases.case("dome_48_21")
    # make ice reach the domain edge so the override matters
    w = cfg.w_geom
    tilt = 300.0 * (np.arange(grid.Mx)[None, :] / grid.Mx + np.arange(grid.My)[:, None] / grid.My)
    for k in ("thickness", "surface"):
        inputs[k][w:-w, w:-w] = np.maximum(inputs[k][w:-w, w:-w], 500.0) + tilt
        G.wrap_ghosts(inputs[k], w)
    inputs["mask"][...] = 2.0
    cfg.D_limit = 1.0e9  # this small, steep dome is far above the default 100 m2/s
    run = cases.oracle_run(grid, cfg, inputs, full=True)
    assert run.status == 0
    D = cases.interior(run.a["D"], 1)
    assert np.all(D[:, grid.Mx - 1, :] == 0.0) and np.all(D[grid.My - 1, :, :] == 0.0)   # G2
    assert np.any(cases.interior(run.a["delta_0"], 1)[:, grid.Mx - 1, :] != 0.0)         # G3
    assert np.all(D[:-1, :-1, :] >= 0.0) and run.D_max == D.max()
    # G10: full_update = false leaves u, v untouched
    run2 = cases.oracle_run(grid, cfg, inputs, full=False)
    assert np.all(run2.a["u"] == 0.0) and np.array_equal(run2.a["D"], run.a["D"])
    # limit_diffusivity: D >= D_limit is capped and counted (G13)
    cfg.limit_diffusivity, cfg.D_limit = 1, 0.5 * run.D_max
    run3 = cases.oracle_run(grid, cfg, inputs, full=False)
    assert run3.status == 0 and run3.D_max == cfg.D_limit and run3.f.high_diffusivity_counter > 0
    # not limiting and D_max > D_limit -> the reference throws (SIAFD.cc:752-760)
    cfg.limit_diffusivity = 0
    run4 = cases.oracle_run(grid, cfg, inputs, full=False)
    assert run4.status == 5


@pytest.mark.skipif(not os.path.exists(O.REF_EXACT), reason="oracle/_ref not built (needs /root/reference)")
def test_strain_heating_against_exact_solution_F():
    """the oracle's volumetric strain heating on the test-F state against the reference's exact
    Sigma (exactTestsFG.cc, compiled unmodified): the reference reports maxSig / avSig errors of this size for
    `pismv -test F` on a 61^3 grid (a second-order scheme on a 30 km grid: a few per cent away from the margin)."""
    import cases
    grid, cfg, inputs, _ = cases.case("F")
    run = cases.oracle_run(grid, cfg, inputs, None, full=True)
    assert run.status == 0
    p = cfg.oracle_params(grid)
    p.flow_law, p.fl_n, p.fl_e = O.FLOW_LAWS["arr"], 3.0, 1.0
    sig = np.zeros((grid.My, grid.Mx, grid.Mz))
    a = run.a
    assert O.lib().orc_strain_heating(C.byref(p), O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["enthalpy"]),
                                      O.dptr(a["u"]), O.dptr(a["v"]), O.dptr(sig)) == 0
    ref = C.CDLL(O.REF_EXACT)
    ref.ref_exactFG.argtypes = [C.c_double, C.c_double, C.c_int] + [C.POINTER(C.c_double)] + [C.c_double] + \
        [C.POINTER(C.c_double)] * 7
    Mz = grid.Mz
    z = np.ascontiguousarray(grid.z)
    worst, n, errs = 0.0, 0, []
    for j in range(grid.My):
        for i in range(grid.Mx):
            r = float(np.hypot(grid.x[i], grid.y[j]))
            if not (100e3 < r < 600e3):  # away from the dome (Sigma -> 0) and from the margin (one-sided stencils)
                continue
            H, M = C.c_double(), C.c_double()
            outs = [np.zeros(Mz) for _ in range(5)]
            assert ref.ref_exactFG(0.0, r, Mz, O.dptr(z), 0.0, C.byref(H), C.byref(M), *[O.dptr(o) for o in outs]) == 0
            exact = outs[3] * (910.0 * 2009.0)  # K s-1 -> W m-3, as compute_strain_heating_errors does (iCMthermo.cc:412)
            ks = grid.k_below_height(H.value)
            k = slice(1, max(ks - 1, 2))  # the base level uses a one-sided (first-order) u_z: ~15 % low there
            scale = np.abs(exact[k]).max()
            errs.append(np.abs(sig[j, i, k] - exact[k]).max() / scale)
            worst = max(worst, errs[-1])
            n += 1
    assert n > 500
    print("strain heating vs exact F: worst %.4f mean %.4f" % (worst, np.mean(errs)))
    assert worst < 0.05 and np.mean(errs) < 0.03, (worst, np.mean(errs))


@pytest.mark.skipif(not os.path.exists(O.REF_EXACT), reason="oracle/_ref not built (needs /root/reference)")
def test_vertical_velocity_against_exact_solution_F():
    """the oracle's w from incompressibility on the test-F state against the reference's exact w(z)
    (exactTestsFG.cc, compiled unmodified).  Test F is steady: w balances the accumulation; the reference's own golden
    surface error for Test G after 1000 a on a coarser grid is 0.028 m/a (maxW, test_17.sh)."""
    import cases
    grid, cfg, inputs, _ = cases.case("F")
    run = cases.oracle_run(grid, cfg, inputs, None, full=True)
    assert run.status == 0
    p, a = run.p, run.a
    w = np.zeros((grid.My, grid.Mx, grid.Mz))
    assert O.lib().orc_vertical_velocity(C.byref(p), O.dptr(a["mask"]), O.dptr(a["u"]), O.dptr(a["v"]), None, 0,
                                         O.dptr(w)) == 0
    ref = C.CDLL(O.REF_EXACT)
    ref.ref_exactFG.argtypes = [C.c_double, C.c_double, C.c_int] + [C.POINTER(C.c_double)] + [C.c_double] + \
        [C.POINTER(C.c_double)] * 7
    Mz, z = grid.Mz, np.ascontiguousarray(grid.z)
    secpera = 31556926.0
    errs = []
    for j in range(grid.My):
        for i in range(grid.Mx):
            r = float(np.hypot(grid.x[i], grid.y[j]))
            if not (100e3 < r < 600e3):
                continue
            H, M = C.c_double(), C.c_double()
            outs = [np.zeros(Mz) for _ in range(5)]
            assert ref.ref_exactFG(0.0, r, Mz, O.dptr(z), 0.0, C.byref(H), C.byref(M), *[O.dptr(o) for o in outs]) == 0
            ks = grid.k_below_height(H.value)
            errs.append(np.abs(w[j, i, :ks + 1] - outs[2][:ks + 1]).max() * secpera)
    errs = np.array(errs)
    print("w vs exact F: worst %.4f m/a mean %.4f m/a (|w| up to %.2f m/a)" % (errs.max(), errs.mean(),
                                                                                 np.abs(w).max() * secpera))
    assert len(errs) > 500 and errs.max() < 0.002 and errs.mean() < 0.001


def test_surface_values_and_horizontal_slices():
    """IceModelVec3::getSurfaceValues / getHorSlice (util/iceModelVec3.cc:153-240), the reads
    PISM.sia.computeSIASurfaceVelocities (site-packages/PISM/sia.py:63-72) and siafd_test.cc:105-151 make of u, v:
    the restatement against getValZ written out literally, its end-level rules, and -- the pin -- the surface speeds
    of one update on the Test F state against exactFG, the check siafd_test itself makes."""
    import oracle_lib as O
    grid, cfg, inputs, _ = cases.case("Fs")
    run = cases.oracle_run(grid, cfg, inputs, full=True)
    assert run.status == 0
    p, w, wuv, z = run.p, cfg.w_geom, cfg.w_uv, grid.z
    us = O.value_at_height(p, run.a["u"], wuv, inputs["thickness"], w)
    vs = O.value_at_height(p, run.a["v"], wuv, inputs["thickness"], w)
    u, H = cases.interior(run.a["u"], wuv), cases.interior(inputs["thickness"], w)
    for j in range(grid.My):
        for i in range(grid.Mx):
            h = H[j, i]
            if h >= z[-1]:
                lit = u[j, i, -1]
            elif h <= z[0]:
                lit = u[j, i, 0]
            else:
                k = grid.k_below_height(h)
                incr = (h - z[k]) / (z[k + 1] - z[k])
                lit = u[j, i, k] + incr * (u[j, i, k + 1] - u[j, i, k])
            assert us[j, i] == lit, (i, j)
    # end levels, a height exactly on a level, and one between two levels
    E = inputs["enthalpy"]
    Ei = cases.interior(E, cfg.w_3d_in)
    assert np.array_equal(O.value_at_height(p, E, cfg.w_3d_in, z0=-5.0), Ei[:, :, 0])
    assert np.array_equal(O.value_at_height(p, E, cfg.w_3d_in, z0=0.0), Ei[:, :, 0])
    assert np.array_equal(O.value_at_height(p, E, cfg.w_3d_in, z0=z[-1]), Ei[:, :, -1])
    assert np.array_equal(O.value_at_height(p, E, cfg.w_3d_in, z0=z[-1] + 1.0), Ei[:, :, -1])
    assert np.array_equal(O.value_at_height(p, E, cfg.w_3d_in, z0=z[7]), Ei[:, :, 7])
    zm = 0.25 * z[7] + 0.75 * z[8]
    mid = O.value_at_height(p, E, cfg.w_3d_in, z0=zm)
    assert np.allclose(mid, 0.25 * Ei[:, :, 7] + 0.75 * Ei[:, :, 8], rtol=1e-14, atol=0.0)
    # siafd_test.cc:105-151: max error of the surface velocity vector against exactFG on 1 m <= r <= L - 1 m
    Uex, r = cases.interior(inputs["exact_surface_speed"], w), cases.interior(inputs["radius"], w)
    X, Y = np.meshgrid(grid.x, grid.y)
    sel = (r >= 1.0) & (r <= 750000.0 - 1.0) & (H > 0)
    rr = np.where(sel, r, 1.0)
    err = np.hypot(us - X / rr * Uex, vs - Y / rr * Uex)[sel]
    assert sel.sum() > 300 and err.max() * V.SperA < 1.0 and err.max() < 0.2 * Uex.max(), err.max() * V.SperA

