"""Deterministic synthetic SIAFD inputs (closed-form, no RNG), SURVEY.md section 8(d).

Every generator returns the LOCAL ghosted arrays of one patch in PISM layout ([j][i][dof]);
ghost values are the periodic images (the DMDA is always periodic, IceGrid.cc:870-872), evaluated
directly from the closed forms, so no exchange is needed to build any rank's inputs.
Written with torch so the same code fills host arrays (tests, CPU baseline) or device arrays
(bench.py, 4096 x 4096 x 101 does not need a 14 GB host staging copy).
"""
import math

import numpy as np
import torch

from . import verification

F64 = torch.float64


def ec_constants(cfg):
    return dict(p_air=cfg.ec_p_air, g=cfg.ec_g, beta=cfg.ec_beta, rho_i=cfg.ec_rho_i, c_i=cfg.ec_c_i,
                c_w=cfg.ec_c_w, L=cfg.ec_L, T_melting=cfg.ec_T_melting, T_0=cfg.ec_T_0)


def enthalpy_permissive(ec, T, omega, P):
    """EnthalpyConverter::enthalpy_permissive (src/util/EnthalpyConverter.cc:277-285) on tensors."""
    T_m = ec["T_melting"] - ec["beta"] * P
    cold = T < T_m
    E_cold = ec["c_i"] * (T - ec["T_0"])
    E_cts = ec["c_i"] * (T_m - ec["T_0"])
    Lm = ec["L"] + (ec["c_w"] - ec["c_i"]) * (T_m - 273.15)
    E_temp = E_cts + torch.clamp(omega, 0.0, 1.0) * Lm
    return torch.where(cold, E_cold, E_temp)


def pressure(ec, depth):
    """EnthalpyConverter::pressure(depth), scalar form with the depth < 0 clamp (:137-143)."""
    return torch.where(depth >= 0.0, ec["p_air"] + ec["rho_i"] * ec["g"] * depth, torch.full_like(depth, ec["p_air"]))


def geometry_calculator(cfg, sea_level, bed, thickness):
    """GeometryCalculator::compute (src/util/Mask.hh:96-133) -> (mask as float64, surface)."""
    alpha = 1 - cfg.ec_rho_i / cfg.sea_water_density
    hgrounded = bed + thickness
    hfloating = sea_level + alpha * thickness
    is_floating = (hfloating > hgrounded) & (not bool(cfg.dry_simulation))
    ice_free = thickness <= cfg.ice_free_thickness
    surface = torch.where(is_floating, hfloating, hgrounded)
    mask = torch.where(is_floating, torch.where(ice_free, 4.0, 3.0), torch.where(ice_free, 0.0, 2.0)).to(F64)
    return mask, surface


def _coords(grid, patch, w, device):
    """x[i], y[j] of the local ghosted index ranges, through the periodic wrap."""
    ii = torch.arange(patch.xs - w, patch.xs + patch.xm + w, device=device) % grid.Mx
    jj = torch.arange(patch.ys - w, patch.ys + patch.ym + w, device=device) % grid.My
    x = torch.as_tensor(grid.x, dtype=F64, device=device)[ii]
    y = torch.as_tensor(grid.y, dtype=F64, device=device)[jj]
    return x, y


def _alloc(shape, device, pin=False):
    if pin and device == "cpu":
        return torch.empty(shape, dtype=F64, pin_memory=True)
    return torch.empty(shape, dtype=F64, device=device)


def dome_2d(grid, cfg, x, y, variant="flat", H0=3600.0, Rfrac=0.75):
    """2D fields of the Halfar-shaped dome (shape of exactTestsABCD.c:77-78 at t = t0, n = 3)."""
    X, Y = torch.meshgrid(x, y, indexing="xy")  # [len(y), len(x)]
    r = torch.sqrt(X * X + Y * Y)
    R = Rfrac * grid.Lx
    s = torch.clamp(1.0 - (r / R) ** (4.0 / 3.0), min=0.0)
    H = torch.where(r < R, H0 * s ** (3.0 / 7.0), torch.zeros_like(r))
    if variant == "flat":
        bed = torch.zeros_like(r)
    else:  # rough bed of test/bed_smoother.py:63-64, scaled to this box
        bed = 400.0 * torch.sin(2.0 * math.pi * X / (0.5 * grid.Lx)) + \
            100.0 * torch.sin(2.0 * math.pi * (X + 1.5 * Y) / (grid.Lx / 30.0))
    sea = torch.zeros_like(r)
    mask, surface = geometry_calculator(cfg, sea, bed, H)
    return dict(r=r, R=R, thickness=H, bed=bed, mask=mask, surface=surface)


def _enthalpy_rows(ec, z, H, r, R, rows, out):
    """Enthalpy of the rows `rows` (slice over j) of a dome-like geometry, into out[rows]."""
    Hs = H[rows][..., None]                     # [nj, ni, 1]
    rr = (r[rows] / R)[..., None]
    zz = z[None, None, :]
    depth = Hs - zz
    P = pressure(ec, depth)
    Ts = 238.15 + 15.0 * torch.clamp(rr, max=1.0)
    P_base = pressure(ec, Hs)
    Tm_base = ec["T_melting"] - ec["beta"] * P_base
    # basal temperature: at the pressure-melting point inside 0.3 R, cooling to 263.15 K at the margin
    wgt = torch.clamp((rr - 0.3) / 0.7, 0.0, 1.0)
    Tb = Tm_base + (263.15 - Tm_base) * wgt
    zeta = torch.where(Hs > 0, torch.clamp(zz / torch.clamp(Hs, min=1e-30), max=1.0), torch.ones_like(depth))
    T = Ts + (Tb - Ts) * (1.0 - zeta) ** 2
    omega = torch.where((zeta < 0.05) & (rr < 0.3) & (Hs > 0), 0.005, 0.0).to(F64).expand_as(T)
    T = torch.where(omega > 0, ec["T_melting"] - ec["beta"] * P + 1.0, T)  # forces the temperate branch
    out[rows] = enthalpy_permissive(ec, T, omega, P)


def dome(grid, patch, cfg, device="cpu", variant="flat", pin=False, row_chunk=64, with_3d=True, Rfrac=0.75):
    """Config C5 (and its scaled-down versions): synthetic dome, gpbld-ready enthalpy.  Rfrac = margin radius / Lx
    (0.75: 44 % of the columns carry ice; >= sqrt(2): every column does, bench.py --regime allice)."""
    wg, we, ws = cfg.w_geom, cfg.w_3d_in, cfg.w_sliding
    ec = ec_constants(cfg)
    xg, yg = _coords(grid, patch, wg, device)
    g2 = dome_2d(grid, cfg, xg, yg, variant, Rfrac=Rfrac)
    out = {k: g2[k].contiguous() for k in ("surface", "thickness", "mask", "bed")}
    out["sliding"] = torch.zeros((patch.ym + 2 * ws, patch.xm + 2 * ws, 2), dtype=F64, device=device)
    if with_3d:
        xe, ye = _coords(grid, patch, we, device)
        g3 = dome_2d(grid, cfg, xe, ye, variant, Rfrac=Rfrac)
        z = torch.as_tensor(grid.z, dtype=F64, device=device)
        E = _alloc((patch.ym + 2 * we, patch.xm + 2 * we, grid.Mz), device, pin)
        nj = E.shape[0]
        for j0 in range(0, nj, row_chunk):
            _enthalpy_rows(ec, z, g3["thickness"], g3["r"], g3["R"], slice(j0, min(j0 + row_chunk, nj)), E)
        out["enthalpy"] = E
    return out


def greenland_like(grid, patch, cfg, device="cpu", row_chunk=64):
    """Config C4: elliptical dome on a bowl-and-ridge bed with surrounding ocean (all four mask values)."""
    wg, we, ws = cfg.w_geom, cfg.w_3d_in, cfg.w_sliding
    ec = ec_constants(cfg)

    def two_d(x, y):
        X, Y = torch.meshgrid(x, y, indexing="xy")
        xc, yc = 0.08 * grid.Lx, -0.05 * grid.Ly
        a, b = 0.70 * grid.Lx, 0.84 * grid.Ly
        rho = torch.sqrt(((X - xc) / a) ** 2 + ((Y - yc) / b) ** 2)
        s = torch.clamp(1.0 - rho ** (4.0 / 3.0), min=0.0)
        H = torch.where(rho < 1.0, 3200.0 * s ** (3.0 / 7.0), torch.zeros_like(rho))
        # bowl rising towards the ice divide, falling below sea level offshore, plus ridges
        bed = (-300.0 + 500.0 * torch.clamp(1.2 - rho, min=-1.0)) + \
            200.0 * torch.sin(2.0 * math.pi * X / 150.0e3) * torch.cos(2.0 * math.pi * Y / 210.0e3)
        # a thin floating fringe just outside the grounded margin, then open ocean
        fringe = (rho >= 1.0) & (rho < 1.06) & (bed < -150.0)
        H = torch.where(fringe, torch.full_like(H, 120.0), H)
        sea = torch.zeros_like(rho)
        mask, surface = geometry_calculator(cfg, sea, bed, H)
        return dict(r=rho, R=1.0, thickness=H, bed=bed, mask=mask, surface=surface)

    xg, yg = _coords(grid, patch, wg, device)
    g2 = two_d(xg, yg)
    out = {k: g2[k].contiguous() for k in ("surface", "thickness", "mask", "bed")}
    out["sliding"] = torch.zeros((patch.ym + 2 * ws, patch.xm + 2 * ws, 2), dtype=F64, device=device)
    xe, ye = _coords(grid, patch, we, device)
    g3 = two_d(xe, ye)
    z = torch.as_tensor(grid.z, dtype=F64, device=device)
    E = torch.empty((patch.ym + 2 * we, patch.xm + 2 * we, grid.Mz), dtype=F64, device=device)
    for j0 in range(0, E.shape[0], row_chunk):
        _enthalpy_rows(ec, z, g3["thickness"], g3["r"], 1.0, slice(j0, min(j0 + row_chunk, E.shape[0])), E)
    out["enthalpy"] = E
    return out


def global_bed(grid, cfg, kind="greenland_like"):
    """Global (Mx*My, no ghosts) bed for BedSmoother::preprocess_bed, as numpy [My, Mx]."""
    if kind == "greenland_like":
        w = cfg.w_geom
        return greenland_like(grid, grid.whole(), cfg)["bed"][w:-w, w:-w].contiguous().numpy()
    x = torch.as_tensor(grid.x, dtype=F64)
    y = torch.as_tensor(grid.y, dtype=F64)
    return dome_2d(grid, cfg, x, y, kind)["bed"].contiguous().numpy()


def test_C_state(grid, patch, cfg, t_years=None):
    """pismv -test C initial state (iceCompModel.cc:301-356): exactC thickness, flat bed, grounded,
    isothermal.  Enthalpy is irrelevant to isothermal_glen; a constant cold value is used."""
    t = 15208.0 * verification.SperA if t_years is None else t_years * verification.SperA
    wg, we, ws = cfg.w_geom, cfg.w_3d_in, cfg.w_sliding
    out = {}
    xg, yg = _coords(grid, patch, wg, "cpu")
    X, Y = np.meshgrid(xg.numpy(), yg.numpy())
    H, _ = verification.exactC(t, np.sqrt(X * X + Y * Y))
    H = torch.as_tensor(H, dtype=F64)
    bed = torch.zeros_like(H)
    mask, surface = geometry_calculator(cfg, torch.zeros_like(H), bed, H)
    out.update(surface=surface, thickness=H, mask=mask, bed=bed)
    out["sliding"] = torch.zeros((patch.ym + 2 * ws, patch.xm + 2 * ws, 2), dtype=F64)
    E0 = cfg.ec_c_i * (263.15 - cfg.ec_T_0)
    out["enthalpy"] = torch.full((patch.ym + 2 * we, patch.xm + 2 * we, grid.Mz), E0, dtype=F64)
    return out


def test_FG_state(grid, patch, cfg, t_years=0.0, Cp=0.0):
    """Test F (t = 0, Cp = 0) / G state: thickness and temperature from exactFG, cold enthalpy
    (iCMthermo.cc:101-136; siafd_test.cc:154-224).  Also returns the exact surface speed."""
    LforFG = 750000.0
    ST, Tmin = 1.67e-5, 223.15
    wg, we, ws = cfg.w_geom, cfg.w_3d_in, cfg.w_sliding
    ec = ec_constants(cfg)
    z = np.asarray(grid.z)
    cache = {}

    def column(r):
        key = float(r)
        if key not in cache:
            cache[key] = verification.exactFG(t_years * verification.SperA, key, z, Cp)
        return cache[key]

    def fill(w, want3d):
        x, y = _coords(grid, patch, w, "cpu")
        X, Y = np.meshgrid(x.numpy(), y.numpy())
        r = np.maximum(np.sqrt(X * X + Y * Y), 1.0)  # avoid singularity at origin
        H = np.zeros_like(r)
        T = np.empty(r.shape + (grid.Mz,)) if want3d else None
        Us = np.zeros_like(r)
        for j in range(r.shape[0]):
            for i in range(r.shape[1]):
                if r[j, i] > LforFG - 1.0:
                    H[j, i] = 0.0
                    if want3d:
                        T[j, i, :] = Tmin + ST * r[j, i]
                else:
                    e = column(r[j, i])
                    H[j, i] = e["H"]
                    if want3d:
                        T[j, i, :] = e["T"]
                    Us[j, i] = verification.exactFG(t_years * verification.SperA, float(r[j, i]),
                                                    np.array([e["H"]]), Cp)["U"][0]
        return r, H, T, Us

    r2, H2, _, Us = fill(wg, False)
    H = torch.as_tensor(H2, dtype=F64)
    bed = torch.zeros_like(H)
    mask, surface = geometry_calculator(cfg, torch.zeros_like(H), bed, H)
    out = dict(surface=surface, thickness=H, mask=mask, bed=bed)
    out["sliding"] = torch.zeros((patch.ym + 2 * ws, patch.xm + 2 * ws, 2), dtype=F64)
    r3, H3, T3, _ = fill(we, True)
    depth = torch.as_tensor(H3, dtype=F64)[..., None] - torch.as_tensor(z, dtype=F64)[None, None, :]
    out["enthalpy"] = enthalpy_permissive(ec, torch.as_tensor(T3, dtype=F64), torch.zeros_like(depth), pressure(ec, depth))
    out["exact_surface_speed"] = torch.as_tensor(Us, dtype=F64)
    out["radius"] = torch.as_tensor(r2, dtype=F64)
    return out
