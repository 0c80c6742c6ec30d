"""ctypes access to the CPU oracle (oracle/liboracle_siafd.so) -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from pism_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle_siafd.so")
REF_EXACT = os.path.join(ORACLE_DIR, "_ref", "libpism_exact.so")

_i32, _f64, _pd = C.c_int32, C.c_double, C.POINTER(C.c_double)

FLOW_LAWS = {"isothermal_glen": 0, "pb": 1, "gpbld": 2, "hooke": 3, "arr": 4, "arrwarm": 5, "gk": 6}
GRADIENTS = {"haseloff": 0, "mahaffy": 1, "eta": 2}

# struct orc_params, oracle/siafd_oracle.h (same member order as siafd_b200_config by construction)
PARAM_FIELDS = [
    ("Mx", _i32), ("My", _i32), ("Mz", _i32), ("xs", _i32), ("xm", _i32), ("ys", _i32), ("ym", _i32),
    ("dx", _f64), ("dy", _f64), ("z", _pd),
    ("w_geom", _i32), ("w_3d_in", _i32), ("w_stag", _i32), ("w_uv", _i32), ("w_sliding", _i32), ("pad0", _i32),
    ("ec_p_air", _f64), ("ec_g", _f64), ("ec_beta", _f64), ("ec_rho_i", _f64), ("ec_c_i", _f64), ("ec_c_w", _f64),
    ("ec_L", _f64), ("ec_T_melting", _f64), ("ec_T_0", _f64),
    ("flow_law", _i32), ("pad1", _i32),
    ("fl_n", _f64), ("fl_e", _f64), ("fl_e_interglacial", _f64),
    ("fl_A_cold", _f64), ("fl_A_warm", _f64), ("fl_Q_cold", _f64), ("fl_Q_warm", _f64), ("fl_T_crit", _f64),
    ("fl_R", _f64), ("fl_rho", _f64), ("fl_g", _f64), ("fl_beta", _f64), ("fl_T_melting", _f64),
    ("gpbld_T_0", _f64), ("gpbld_water_frac_coeff", _f64), ("gpbld_water_frac_limit", _f64),
    ("iso_softness_A", _f64),
    ("hooke_Q", _f64), ("hooke_A", _f64), ("hooke_C", _f64), ("hooke_K", _f64), ("hooke_Tr", _f64),
    ("grain_size", _f64),
    ("gradient_method", _i32), ("limit_diffusivity", _i32), ("grain_size_age_coupling", _i32),
    ("e_age_coupling", _i32),
    ("D_limit", _f64), ("eemian_start", _f64), ("eemian_end", _f64), ("holocene_start", _f64),
    ("years_per_second", _f64),
    ("smoother_range", _f64), ("theta_min", _f64),
    ("sea_water_density", _f64), ("ice_free_thickness", _f64),
    ("dry_simulation", _i32), ("pad2", _i32),
]


class Params(C.Structure):
    _fields_ = PARAM_FIELDS


class Fields(C.Structure):
    _fields_ = [("surface", _pd), ("thickness", _pd), ("mask", _pd), ("bed", _pd), ("enthalpy", _pd), ("age", _pd),
                ("sliding", _pd), ("topgsmooth", _pd), ("maxtl", _pd), ("C2", _pd), ("C3", _pd), ("C4", _pd),
                ("smoother_active", _i32), ("pad", _i32), ("current_time", _f64),
                ("h_x", _pd), ("h_y", _pd), ("D", _pd), ("Q", _pd), ("u", _pd), ("v", _pd),
                ("work2d_0", _pd), ("work2d_1", _pd), ("delta_0", _pd), ("delta_1", _pd), ("I_0", _pd), ("I_1", _pd),
                ("D_max", _f64), ("high_diffusivity_counter", _i32), ("pad3", _i32)]


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("siafd_oracle.cc", "mass_oracle.cc", "siafd_oracle.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        PP, FP = C.POINTER(Params), C.POINTER(Fields)
        L.orc_default_params.argtypes = [PP]
        L.orc_flow.restype = _f64
        L.orc_flow.argtypes = [PP, _f64, _f64, _f64, _f64]
        for n in ("orc_ec_pressure", "orc_ec_melting_temperature", "orc_ec_enthalpy_cts"):
            getattr(L, n).restype = _f64
            getattr(L, n).argtypes = [PP, _f64]
        for n in ("orc_ec_temperature", "orc_ec_pressure_adjusted_temperature", "orc_ec_water_fraction"):
            getattr(L, n).restype = _f64
            getattr(L, n).argtypes = [PP, _f64, _f64]
        for n in ("orc_ec_enthalpy", "orc_ec_enthalpy_permissive"):
            getattr(L, n).restype = _f64
            getattr(L, n).argtypes = [PP, _f64, _f64, _f64]
        L.orc_grain_size_vostok.restype = _f64
        L.orc_grain_size_vostok.argtypes = [_f64]
        L.orc_vertical_levels.argtypes = [_f64, C.c_int, C.c_int, _f64, _pd]
        L.orc_k_below_height.restype = C.c_int
        L.orc_k_below_height.argtypes = [_pd, C.c_int, _f64, C.POINTER(C.c_int)]
        L.orc_compute_nprocs.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_ownership_ranges.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_geometry_compute.argtypes = [PP, C.c_int, _pd, _pd, _pd, _pd, _pd]
        L.orc_preprocess_bed.argtypes = [PP, _pd, _pd, _pd, _pd, _pd, _pd, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_theta.argtypes = [PP, FP, _pd]
        L.orc_smoothed_thk.argtypes = [PP, FP, _pd]
        L.orc_wrap_ghosts.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _pd]
        L.orc_siafd_gradient.argtypes = [PP, FP]
        L.orc_siafd_flux_velocity.argtypes = [PP, FP, C.c_int]
        L.orc_siafd_update_single.argtypes = [PP, FP, C.c_int]
        L.orc_siafd_update_many.argtypes = [C.c_int, PP, FP, C.c_int, C.c_int]
        L.orc_siafd_update_decomposed.argtypes = [C.c_int, PP, FP, C.c_int, C.c_int]
        L.orc_vertical_velocity.argtypes = [PP, _pd, _pd, _pd, _pd, C.c_int, _pd]
        L.orc_value_at_height.restype = None
        L.orc_value_at_height.argtypes = [PP, _pd, C.c_int, _pd, C.c_int, _f64, _pd]
        L.orc_strain_heating.argtypes = [PP, _pd, _pd, _pd, _pd, _pd, _pd]
        L.orc_mass_flow_step.argtypes = [PP, _f64, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd]
        L.orc_mass_flow_step_part_grid.argtypes = [PP, _f64, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd, C.c_int, _pd, _pd,
                                                   _pd, _pd]
        L.orc_mass_source_step.argtypes = [PP, _f64, _f64, C.c_int, _pd, _pd, _pd, _pd, _pd, _pd, _pd]
        L.orc_cfl_3d.argtypes = [PP, _f64, _pd, _pd, _pd, _pd, _pd, _pd]
        L.orc_cfl_2d.argtypes = [PP, _f64, _pd, _pd, _pd]
        L.orc_regional_gradient_override.argtypes = [PP, _pd, _pd, _pd, _pd, _pd]
        _lib = L
    return _lib


def default_params():
    p = Params()
    lib().orc_default_params(C.byref(p))
    return p


def params_from_config(cfg):
    """Copy a pism_b200.capi.Config (same member list) into oracle Params."""
    p = Params()
    for name, _ in PARAM_FIELDS:
        setattr(p, name, getattr(cfg, name))
    return p


def dptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pd)


def shape(p, w, dof=1):
    s = (p.ym + 2 * w, p.xm + 2 * w)
    return s if dof == 1 else s + (dof,)


class Run:
    """Arrays and orc_fields of one oracle patch; keeps every buffer alive."""

    def __init__(self, p, inputs, smoothed=None, current_time=0.0):
        self.p = p
        self.a = {}
        f = Fields()
        for name in ("surface", "thickness", "mask", "bed", "enthalpy", "age", "sliding"):
            v = inputs.get(name)
            if v is not None:
                v = np.ascontiguousarray(np.asarray(v), dtype=np.float64)
                self.a[name] = v
                setattr(f, name, dptr(v))
        if "sliding" not in self.a:
            self.a["sliding"] = np.zeros(shape(p, p.w_sliding, 2))
            f.sliding = dptr(self.a["sliding"])
        if smoothed is None:
            # smoother off: topgsmooth = ghosted copy of the bed, rest zero (BedSmoother.cc:101-109)
            smoothed = dict(topgsmooth=self.a["bed"].copy(), maxtl=np.zeros(shape(p, p.w_geom)),
                            C2=np.zeros(shape(p, p.w_geom)), C3=np.zeros(shape(p, p.w_geom)),
                            C4=np.zeros(shape(p, p.w_geom)), active=0)
        for name in ("topgsmooth", "maxtl", "C2", "C3", "C4"):
            self.a[name] = np.ascontiguousarray(smoothed[name], dtype=np.float64)
            setattr(f, name, dptr(self.a[name]))
        f.smoother_active = int(smoothed["active"])
        f.current_time = current_time
        out = dict(h_x=(p.w_stag, 2), h_y=(p.w_stag, 2), D=(p.w_stag, 2), Q=(p.w_stag, 2), u=(p.w_uv, p.Mz),
                   v=(p.w_uv, p.Mz), work2d_0=(p.w_geom, 1), work2d_1=(p.w_geom, 1), delta_0=(p.w_stag, p.Mz),
                   delta_1=(p.w_stag, p.Mz), I_0=(p.w_stag, p.Mz), I_1=(p.w_stag, p.Mz))
        for name, (w, dof) in out.items():
            self.a[name] = np.zeros(shape(p, w, dof))
            setattr(f, name, dptr(self.a[name]))
        self.f = f

    def update_single(self, full=True):
        return lib().orc_siafd_update_single(C.byref(self.p), C.byref(self.f), 1 if full else 0)

    def gradient(self):
        return lib().orc_siafd_gradient(C.byref(self.p), C.byref(self.f))

    def flux_velocity(self, full=True):
        return lib().orc_siafd_flux_velocity(C.byref(self.p), C.byref(self.f), 1 if full else 0)

    @property
    def D_max(self):
        return self.f.D_max


def value_at_height(p, a, wa, heights=None, wh=0, z0=0.0):
    """orc_value_at_height: IceModelVec3::getSurfaceValues (heights = thickness) / getHorSlice (heights None)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.shape == shape(p, wa, p.Mz), (a.shape, shape(p, wa, p.Mz))
    out = np.zeros((p.ym, p.xm))
    hp = None
    if heights is not None:
        heights = np.ascontiguousarray(heights, dtype=np.float64)
        assert heights.shape == shape(p, wh)
        hp = dptr(heights)
    lib().orc_value_at_height(C.byref(p), dptr(a), wa, hp, wh, z0, dptr(out))
    return out


def preprocess_bed(p, topg_global):
    """orc_preprocess_bed on a global [My, Mx] bed -> dict of global arrays + (Nx, Ny)."""
    g = np.ascontiguousarray(topg_global, dtype=np.float64)
    outs = [np.zeros_like(g) for _ in range(5)]
    Nx, Ny = C.c_int(), C.c_int()
    st = lib().orc_preprocess_bed(C.byref(p), dptr(g), *[dptr(o) for o in outs], C.byref(Nx), C.byref(Ny))
    assert st == 0, st
    return dict(topgsmooth=outs[0], maxtl=outs[1], C2=outs[2], C3=outs[3], C4=outs[4], Nx=Nx.value, Ny=Ny.value,
                active=int(Nx.value >= 0))
