"""ctypes binding of the C ABI declared in include/siafd_b200.h (one-to-one, no logic)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsiafd_b200.so")

# flow laws / gradient methods / status codes / field ids: keep in sync with siafd_b200.h
FLOW_LAWS = {"isothermal_glen": 0, "pb": 1, "gpbld": 2, "hooke": 3, "arr": 4, "arrwarm": 5, "gk": 6}
GRADIENTS = {"haseloff": 0, "mahaffy": 1, "eta": 2}
OK, ERR_NEGATIVE_THICKNESS, ERR_OMEGA_NEGATIVE, ERR_HEIGHT_BELOW_BASE, ERR_HEIGHT_ABOVE_TOP = 0, 1, 2, 3, 4
ERR_DIFFUSIVITY, ERR_BAD_CONFIG, ERR_CUDA, ERR_BAD_ARGUMENT, ERR_COMM = 5, 6, 7, 8, 9

FIELDS = ["surface", "thickness", "mask", "bed", "enthalpy", "age", "sliding", "topgsmooth", "maxtl", "C2", "C3",
          "C4", "h_x", "h_y", "D", "flux", "u", "v", "thk_smooth", "theta", "w_i", "w_j", "w", "basal_melt",
          "sea_level", "smb", "thk_change", "flux_div", "cons_err", "eff_smb", "eff_bmb", "vel_bc_mask", "thk_bc_mask", "strain_heating",
          "no_model_mask", "no_model_surface", "h_x_no_model", "h_y_no_model"]
F = {name: i for i, name in enumerate(FIELDS)}

_i32, _f64, _pd = C.c_int32, C.c_double, C.POINTER(C.c_double)

# (name, ctype) in the exact order of struct siafd_b200_config
CONFIG_FIELDS = [
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


class Config(C.Structure):
    _fields_ = CONFIG_FIELDS


class Inputs(C.Structure):
    _fields_ = [("surface", _pd), ("thickness", _pd), ("mask", _pd), ("bed", _pd), ("enthalpy", _pd), ("age", _pd),
                ("sliding", _pd), ("current_time", _f64), ("memory_space", _i32), ("ghosts_valid", _i32)]


class Outputs(C.Structure):
    _fields_ = [("h_x", _pd), ("h_y", _pd), ("D", _pd), ("flux", _pd), ("u", _pd), ("v", _pd),
                ("memory_space", _i32), ("pad", _i32)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pism_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C pism_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    sig = {
        "siafd_b200_abi_version": (C.c_int, []),
        "siafd_b200_default_config": (None, [C.POINTER(Config)]),
        "siafd_b200_status_string": (C.c_char_p, [C.c_int]),
        "siafd_b200_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(vp)]),
        "siafd_b200_destroy": (None, [vp]),
        "siafd_b200_last_error": (C.c_char_p, [vp]),
        "siafd_b200_field_size": (i64, [vp, C.c_int]),
        "siafd_b200_field_width": (C.c_int, [vp, C.c_int]),
        "siafd_b200_field_dof": (C.c_int, [vp, C.c_int]),
        "siafd_b200_bind": (C.c_int, [vp, C.c_int, vp]),
        "siafd_b200_device_ptr": (vp, [vp, C.c_int]),
        "siafd_b200_set_stream": (C.c_int, [vp, vp]),
        "siafd_b200_upload": (C.c_int, [vp, C.c_int, vp]),
        "siafd_b200_download": (C.c_int, [vp, C.c_int, vp]),
        "siafd_b200_wrap_ghosts": (C.c_int, [vp, C.c_int]),
        "siafd_b200_wrap_ghosts_dir": (C.c_int, [vp, C.c_int, C.c_int]),
        "siafd_b200_wrap_ghosts_many": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int)]),
        "siafd_b200_halo_count": (i64, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "siafd_b200_halo_pack": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "siafd_b200_halo_unpack": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "siafd_b200_ipc_export": (C.c_int, [vp, C.c_int, vp]),
        "siafd_b200_ipc_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "siafd_b200_halo_attach": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]),
        "siafd_b200_halo_push": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
        "siafd_b200_halo_wait": (C.c_int, [vp, C.c_int]),
        "siafd_b200_comm_init": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p, C.c_double]),
        "siafd_b200_comm_init_local": (C.c_int, [C.POINTER(vp), C.c_int]),
        "siafd_b200_comm_rank": (C.c_int, [vp]),
        "siafd_b200_comm_size": (C.c_int, [vp]),
        "siafd_b200_comm_exchange": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "siafd_b200_comm_allreduce": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]),
        "siafd_b200_update_decomposed": (C.c_int, [vp, C.c_int, C.c_double, C.c_int]),
        "siafd_b200_preprocess_bed": (C.c_int, [vp, vp]),
        "siafd_b200_set_smoothed_bed": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int]),
        "siafd_b200_compute_gradient": (C.c_int, [vp]),
        "siafd_b200_compute_flux_velocity": (C.c_int, [vp, C.c_int, C.c_double]),
        "siafd_b200_compute_vertical_velocity": (C.c_int, [vp, C.c_int, C.c_int]),
        "siafd_b200_compute_gradient_no_model": (C.c_int, [vp]),
        "siafd_b200_apply_no_model_gradient": (C.c_int, [vp]),
        "siafd_b200_compute_strain_heating": (C.c_int, [vp, C.c_int, C.c_double, C.c_double]),
        "siafd_b200_surface_values": (C.c_int, [vp, C.c_int, vp]),
        "siafd_b200_hor_slice": (C.c_int, [vp, C.c_int, C.c_double, vp]),
        "siafd_b200_mass_flow_step": (C.c_int, [vp, C.c_double]),
        "siafd_b200_mass_source_step": (C.c_int, [vp, C.c_double, C.c_double, C.c_int]),
        "siafd_b200_ensure_consistency": (C.c_int, [vp, C.c_int]),
        "siafd_b200_cfl": (C.c_int, [vp, C.c_double, C.c_int, C.POINTER(C.c_double)]),
        "siafd_b200_finish": (C.c_int, [vp]),
        "siafd_b200_max_diffusivity": (C.c_double, [vp]),
        "siafd_b200_high_diffusivity_count": (C.c_int, [vp]),
        "siafd_b200_update": (C.c_int, [vp, C.POINTER(Inputs), C.POINTER(Outputs), C.c_int]),
        "siafd_b200_geometry_compute": (C.c_int, [vp, i64, vp, vp, vp, vp, vp]),
        "siafd_b200_flow_n": (C.c_int, [vp, i64, vp, vp, vp, vp, vp]),
        "siafd_b200_flow_host": (C.c_int, [vp, i64, vp, vp, vp, vp, vp]),
        "siafd_b200_set_tuning": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "siafd_b200_step_breakdown_ms": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
        "siafd_b200_launch_count": (i64, [vp]),
        "siafd_b200_transfer_bytes": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "siafd_b200_host_levels_needed": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_double]),
        "siafd_b200_host_plan_emulate": (C.c_int, [C.POINTER(Config)] + [C.c_int] * 6 + [vp] * 11 + [C.POINTER(i64)] * 2),
        "siafd_b200_kernel_timing": (C.c_int, [vp, C.c_int]),
        "siafd_b200_kernel_time_ms": (C.c_double, [vp, C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTS = _load()


def default_config():
    cfg = Config()
    lib.siafd_b200_default_config(C.byref(cfg))
    return cfg
