"""Shared test cases: (grid, capi-style config values, inputs) for the five BASELINE.json configs
and their scaled-down versions.  Used by the CPU oracle tests and by the GPU parity tests, so both
sides see bit-identical inputs.  No reference tree access at run time."""
import ctypes as C

import numpy as np

import oracle_lib as O
from pism_b200 import grid as G
from pism_b200 import synthetic as S

SECPERA_UDUNITS = 365.242198781 * 86400.0  # UDUNITS-2 "year", what convert(1, "year", "seconds") returns


class Cfg:
    """Plain attribute bag with the members of siafd_b200_config / orc_params (defaults from the oracle)."""

    def __init__(self, **kw):
        p = O.default_params()
        for name, _ in O.PARAM_FIELDS:
            if name != "z":
                setattr(self, name, getattr(p, name))
        for k, v in kw.items():
            if k == "flow_law" and isinstance(v, str):
                v = O.FLOW_LAWS[v]
            if k == "gradient_method" and isinstance(v, str):
                v = O.GRADIENTS[v]
            assert hasattr(self, k), k
            setattr(self, k, v)

    def overrides(self):
        return {name: getattr(self, name) for name, _ in O.PARAM_FIELDS
                if name not in ("z", "Mx", "My", "Mz", "xs", "xm", "ys", "ym", "dx", "dy", "pad0", "pad1", "pad2")}

    def oracle_params(self, grid, patch=None):
        patch = patch or grid.whole()
        p = O.Params()
        for name, _ in O.PARAM_FIELDS:
            if name != "z":
                setattr(p, name, getattr(self, name, 0))
        p.Mx, p.My, p.Mz = grid.Mx, grid.My, grid.Mz
        p.xs, p.xm, p.ys, p.ym = patch.xs, patch.xm, patch.ys, patch.ym
        p.dx, p.dy = grid.dx, grid.dy
        p._z_keep = np.ascontiguousarray(grid.z, dtype=np.float64)
        p.z = O.dptr(p._z_keep)
        return p


def cold_converter():
    """ColdEnthalpyConverter, util/EnthalpyConverter.cc:287-296 (pismv.cc:61)."""
    return dict(ec_T_melting=1e6, ec_beta=0.0)


def to_numpy(d):
    return {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in d.items()}


def case(name, patch=None):
    """Returns (grid, cfg, inputs dict of numpy arrays for `patch` or the whole domain, global_bed or None)."""
    gb = None
    if name.startswith("C1"):  # pismv -test C: pismv.cc:96-102, iceCompModel.cc:65-124
        size = 61 if "_" not in name or not name.split("_")[-1].isdigit() else int(name.split("_")[-1])
        # pismv's default levels are QUADRATIC: pismv_grid_defaults sets EQUAL (pismv.cc:82) but
        # vertical_grid_from_options (pismv.cc:155, IceGrid.cc:1289-1296) recomputes them with the configuration's
        # grid.ice_vertical_spacing = "quadratic", lambda = 4 (the golden rows of test_15.sh only reproduce this way)
        grid = G.Grid(size, size, 31, 1000e3, 1000e3, 4000.0, spacing="equal" if "equal" in name else "quadratic")
        grad = "haseloff"
        for g in ("mahaffy", "eta", "haseloff"):
            if g in name:
                grad = g
        cfg = Cfg(flow_law="isothermal_glen", iso_softness_A=1.0e-16 / SECPERA_UDUNITS, smoother_range=0.0, fl_e=1.0,
                  dry_simulation=1, gradient_method=grad, **cold_converter())
        inputs = S.test_C_state(grid, patch or grid.whole(), cfg)
    elif name.startswith("C2"):  # pismv -test G: pismv.cc:103-110; arr; cold converter
        M = {"C2": (121, 61), "C2s": (41, 31), "C2t": (31, 21)}[name.split("_")[0]]
        # full-size C2 uses pismv's quadratic levels (see C1); the scaled-down C2s / C2t keep equal ones (frozen fixture)
        grid = G.Grid(M[0], M[0], M[1], 900e3, 900e3, 4000.0, spacing="quadratic" if name.split("_")[0] == "C2" else "equal")
        cfg = Cfg(flow_law="arr", smoother_range=0.0, fl_e=1.0, dry_simulation=1, **cold_converter())
        inputs = S.test_FG_state(grid, patch or grid.whole(), cfg, t_years=500.0, Cp=200.0)
    elif name.startswith("F"):  # siafd_test.cc: Test F, Lx = Ly = 900 km, Lz = 4000, arr, cold converter
        M = {"F": (61, 61), "Fs": (31, 31)}[name]
        grid = G.Grid(M[0], M[0], M[1], 900e3, 900e3, 4000.0)
        cfg = Cfg(flow_law="arr", smoother_range=0.0, fl_e=1.0, dry_simulation=0, **cold_converter())
        inputs = S.test_FG_state(grid, patch or grid.whole(), cfg, t_years=0.0, Cp=0.0)
    elif name.startswith("C3"):  # EISMINT II F-shaped: pisms.cc:51-55, runexp.sh:43-44; pb; standard converter
        M = {"C3": (151, 101), "C3s": (41, 41)}[name]
        grid = G.Grid(M[0], M[0], M[1], 750e3, 750e3, 6000.0)
        cfg = Cfg(flow_law="pb", smoother_range=0.0, dry_simulation=1)
        inputs = S.dome(grid, patch or grid.whole(), cfg)
    elif name.startswith("C4"):  # Greenland-shaped, gpbld, haseloff, smoother 5 km
        M = {"C4": (301, 561, 101), "C4s": (61, 113, 21)}[name.split("_")[0]]
        grid = G.Grid(M[0], M[1], M[2], 750e3, 1400e3, 4000.0)
        kw = dict(flow_law="gpbld", smoother_range=5.0e3)
        if "nosmooth" in name:
            kw["smoother_range"] = 0.0
        if "limit" in name:
            kw.update(limit_diffusivity=1, D_limit=5.0)
        cfg = Cfg(**kw)
        inputs = S.greenland_like(grid, patch or grid.whole(), cfg)
        gb = S.global_bed(grid, cfg, "greenland_like")
    elif name.startswith("dome"):  # C5 and scaled-down versions: dx = dy = 5 km, Lz = 4000, Mz = 101, gpbld
        parts = name.split("_")
        M = int(parts[1])
        Mz = int(parts[2]) if len(parts) > 2 and parts[2].isdigit() else 101
        L = (M - 1) / 2.0 * 5000.0
        grid = G.Grid(M, M, Mz, L, L, 4000.0)
        kw = dict(flow_law="gpbld", smoother_range=0.0)
        if M < 256:
            kw["D_limit"] = 1.0e9  # small steep domes exceed the default 100 m2/s cap (SIAFD.cc:752-760)
        for fl in ("pb", "hooke", "gk", "arrwarm"):
            if fl in parts:
                kw["flow_law"] = fl
        for g in ("mahaffy", "eta"):
            if g in parts:
                kw["gradient_method"] = g
        if "quadratic" in parts:
            grid = G.Grid(M, M, Mz, L, L, 4000.0, spacing="quadratic")
        if "n4" in parts:
            kw["fl_n"] = 4.0
        cfg = Cfg(**kw)
        inputs = S.dome(grid, patch or grid.whole(), cfg, variant="rough" if "rough" in parts else "flat")
    else:
        raise KeyError(name)
    return grid, cfg, to_numpy(inputs), gb


def oracle_run(grid, cfg, inputs, gb=None, full=True, current_time=0.0, patch=None):
    """Run the CPU oracle on a whole-domain patch.  Returns the oracle_lib.Run (arrays in .a)."""
    p = cfg.oracle_params(grid, patch)
    smoothed = None
    if cfg.smoother_range > 0.0:
        sm = O.preprocess_bed(p, gb)
        pt = patch or grid.whole()
        smoothed = {k: G.global_to_local(sm[k], pt, cfg.w_geom) for k in ("topgsmooth", "maxtl", "C2", "C3", "C4")}
        smoothed["active"] = sm["active"]
    run = O.Run(p, inputs, smoothed, current_time)
    run.status = run.update_single(full) if patch is None else None
    return run


def rel_max(a, b):
    """Max-norm relative difference |a - b|_inf / |b|_inf (the north_star's tolerance definition)."""
    den = np.max(np.abs(b))
    num = np.max(np.abs(np.asarray(a) - np.asarray(b)))
    return 0.0 if num == 0.0 else num / den


def interior(a, w):
    return a if w == 0 else a[w:-w, w:-w]
