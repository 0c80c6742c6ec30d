#!/usr/bin/env python3
"""Generate tests/golden/oracle_fixture_<case>.npz: inputs and oracle outputs of one SIAFD::update for small
cases, frozen so that (a) the GPU path is checked against committed vectors without running the oracle, and
(b) a change in the oracle itself shows up as a diff.  Run from the repo root:  python tools/make_golden.py
The reference itself cannot be built in this image, so these are outputs of the oracle
(the CPU restatement pinned by tests/test_oracle_known_answers.py), not of PISM."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

CASES = ["dome_33_13", "C2t", "C4s"]  # gpbld dome (ragged sizes); Test G state, arr, cold converter; Greenland-like + smoother


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name in CASES:
        grid, cfg, inputs, gb = cases.case(name)
        run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
        assert run.status == 0, (name, run.status)
        data = {"in_" + k: v for k, v in inputs.items()}
        for k in ("h_x", "h_y", "D", "Q", "u", "v", "work2d_0", "work2d_1"):
            data["out_" + k] = run.a[k]
        data["out_D_max"] = np.array(run.D_max)
        if gb is not None:
            data["global_bed"] = gb
        path = os.path.join(out, "oracle_fixture_%s.npz" % name)
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")


def consumers(name="C4s", dt_years=0.5, dt2_years=200.0):
    """oracle_consumers_<case>.npz: the SURVEY 8(f) rows on the frozen update of oracle_fixture_<case>.npz -- vertical
    velocity, 3D / 2D CFL scalars, strain heating (gpbld, n = 3, e = 1), one flow step and one source step."""
    import ctypes as C
    import oracle_lib as O
    sec = 365.242198781 * 86400.0
    grid, cfg, inputs, gb = cases.case(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    p, a, L = run.p, run.a, O.lib()
    n = (grid.My, grid.Mx)
    d = {}
    d["w"] = np.zeros(n + (grid.Mz,))
    assert L.orc_vertical_velocity(C.byref(p), O.dptr(a["mask"]), O.dptr(a["u"]), O.dptr(a["v"]), None, 0, O.dptr(d["w"])) == 0
    c3, c2 = (C.c_double * 4)(), (C.c_double * 4)()
    max_dt = 60.0 * sec
    assert L.orc_cfl_3d(C.byref(p), max_dt, O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["u"]), O.dptr(a["v"]),
                        O.dptr(d["w"]), c3) == 0
    assert L.orc_cfl_2d(C.byref(p), max_dt, O.dptr(a["mask"]), O.dptr(a["sliding"]), c2) == 0
    d["cfl3d"], d["cfl2d"] = np.array(list(c3)), np.array(list(c2))
    ph = cfg.oracle_params(grid)
    ph.flow_law, ph.fl_n, ph.fl_e = O.FLOW_LAWS["gpbld"], 3.0, 1.0
    d["sigma"] = np.zeros(n + (grid.Mz,))
    assert L.orc_strain_heating(C.byref(ph), O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["enthalpy"]),
                                O.dptr(a["u"]), O.dptr(a["v"]), O.dptr(d["sigma"])) == 0
    H = a["thickness"].copy()
    d["flux_div"], d["thk_change"], ce = np.zeros(n), np.zeros(n), np.zeros(n)
    assert L.orc_mass_flow_step(C.byref(p), dt_years * sec, None, O.dptr(a["bed"]), O.dptr(H), None, None, None,
                                O.dptr(a["Q"]), O.dptr(d["flux_div"]), O.dptr(d["thk_change"]), O.dptr(ce)) == 0
    w = cfg.w_geom
    d["H_after_flow"] = H[w:-w, w:-w].copy()
    from pism_b200 import grid as G
    G.wrap_ghosts(H, w)
    mask, surf = np.zeros_like(H), np.zeros_like(H)
    L.orc_geometry_compute(C.byref(p), H.size, O.dptr(np.zeros_like(H)), O.dptr(a["bed"]), O.dptr(H), O.dptr(mask), O.dptr(surf))
    d["mask_after_flow"] = mask
    d["smb"] = (np.random.default_rng(0).random(n) - 0.6) * 3e-3
    es, eb = np.zeros(n), np.zeros(n)
    assert L.orc_mass_source_step(C.byref(p), dt2_years * sec, 910.0, 0, O.dptr(H), O.dptr(mask), None, O.dptr(d["smb"]), None,
                                  O.dptr(es), O.dptr(eb)) == 0
    d["H_after_source"] = H[w:-w, w:-w].copy()
    d["dt"], d["dt2"], d["max_dt"] = np.array(dt_years * sec), np.array(dt2_years * sec), np.array(max_dt)
    path = os.path.join(ROOT, "tests", "golden", "oracle_consumers_%s.npz" % name)
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path) // 1024, "KiB")


def reads_and_transport(name="C4s"):
    """oracle_reads_<case>.npz: IceModelVec3::getSurfaceValues of u, v and a horizontal slice of the enthalpy on the
    frozen update (util/iceModelVec3.cc:153-240); oracle_mass_transport_51.npz: the thickness after 12 CFL-limited
    flow steps of the spreading disc of test/mass_transport.py (default variant, part_grid off; advective velocity,
    both B.C. masks) -- the set-up of tests/test_gpu_mass_transport.py."""
    import ctypes as C
    import oracle_lib as O
    from pism_b200 import grid as G
    grid, cfg, inputs, gb = cases.case(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    H, wg = inputs["thickness"], cfg.w_geom
    d = {"u_surface": O.value_at_height(run.p, run.a["u"], cfg.w_uv, H, wg),
         "v_surface": O.value_at_height(run.p, run.a["v"], cfg.w_uv, H, wg),
         "z_slice": np.array(0.3 * grid.Lz + 1.0),
         "enthalpy_slice": O.value_at_height(run.p, inputs["enthalpy"], cfg.w_3d_in, z0=0.3 * grid.Lz + 1.0)}
    path = os.path.join(ROOT, "tests", "golden", "oracle_reads_%s.npz" % name)
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path) // 1024, "KiB")

    from test_gpu_mass_transport import spreading_disc_setup, oracle_flow_steps
    S = spreading_disc_setup(51)
    oracle_flow_steps(S, 12)
    path = os.path.join(ROOT, "tests", "golden", "oracle_mass_transport_51.npz")
    np.savez_compressed(path, thickness=S["H"], mask=S["mask"], dt_last=np.array(S["dt"]))
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
    consumers()
    reads_and_transport()
