"""CPU-side checks of the C ABI boundary: the library loads, exports every symbol the header declares,
struct layouts agree, config validation mirrors the reference's constructor errors, and there is no
CPU fallback (no device -> SIAFD_B200_ERR_CUDA)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from pism_b200 import capi, grid as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "siafd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(siafd_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 30
    raw = C.CDLL(capi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "libsiafd_b200.so does not export %s" % s
    assert sorted(capi.EXPORTS) == syms, "pism_b200/capi.py and include/siafd_b200.h disagree"
    assert capi.lib.siafd_b200_abi_version() == 2


def test_field_ids_of_the_python_binding_match_the_header():
    text = open(os.path.join(ROOT, "include", "siafd_b200.h")).read()
    ids = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"SIAFD_B200_F_([A-Z0-9_]+)\s*=\s*(\d+)", text)}
    count = ids.pop("count")
    assert count == len(capi.FIELDS) == len(ids)
    assert ids == {k.lower(): v for k, v in capi.F.items()}  # ("C2".."C4", "D" keep their reference spelling)


def test_config_struct_matches_oracle_params_and_defaults():
    """Same member list on both sides, and the defaults are pism_config.cdl's."""
    assert [n for n, _ in capi.CONFIG_FIELDS] == [n for n, _ in O.PARAM_FIELDS]
    assert C.sizeof(capi.Config) == C.sizeof(O.Params)
    c, p = capi.default_config(), O.default_params()
    for name, _ in capi.CONFIG_FIELDS:
        if name != "z":
            assert getattr(c, name) == getattr(p, name), name
    assert (c.fl_n, c.D_limit, c.smoother_range, c.ec_rho_i, c.fl_R) == (3.0, 100.0, 5.0e3, 910.0, 8.31441)
    assert c.flow_law == capi.FLOW_LAWS["gpbld"] and c.gradient_method == capi.GRADIENTS["haseloff"]
    assert abs(c.eemian_start / (365.242198781 * 86400.0) + 132000.0) < 1e-6
    for code in range(9):
        assert capi.lib.siafd_b200_status_string(code)


def _cfg(**kw):
    g = G.Grid(16, 16, 11, 1e5, 1e5, 1000.0)
    c = capi.default_config()
    c.Mx, c.My, c.Mz, c.xs, c.xm, c.ys, c.ym = 16, 16, 11, 0, 16, 0, 16
    c.dx, c.dy = g.dx, g.dy
    z = np.ascontiguousarray(g.z)
    c.z = z.ctypes.data_as(C.POINTER(C.c_double))
    c._keep = z
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("kw", [dict(gradient_method=7), dict(flow_law=9), dict(w_geom=1), dict(w_3d_in=1),
                                dict(grain_size_age_coupling=1), dict(Mz=1), dict(xm=17)])
def test_bad_configuration_is_rejected_before_touching_the_gpu(kw):
    """SIAFD.cc:69-86,216-219 and the stencil-width asserts :587-602 -> SIAFD_B200_ERR_BAD_CONFIG."""
    h = C.c_void_p()
    st = capi.lib.siafd_b200_create(C.byref(_cfg(**kw)), -1, C.byref(h))
    assert st == capi.ERR_BAD_CONFIG and not h.value
    assert len(capi.lib.siafd_b200_last_error(None)) > 0


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="a CUDA device is present")
def test_no_cpu_fallback():
    """Without a device the product path must fail loudly, never compute on the CPU."""
    h = C.c_void_p()
    st = capi.lib.siafd_b200_create(C.byref(_cfg()), -1, C.byref(h))
    assert st == capi.ERR_CUDA and not h.value
    assert b"no CPU path" in capi.lib.siafd_b200_last_error(None)
    from pism_b200.sia import SIAFD, PISMRuntimeError
    with pytest.raises(PISMRuntimeError):
        SIAFD(G.Grid(16, 16, 11, 1e5, 1e5, 1000.0))


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under pism_b200/ or include/ may name it."""
    bad = []
    for base in ("pism_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc", ".cpp", "Makefile")):
                    text = open(os.path.join(dirpath, fn), errors="ignore").read()
                    if re.search(r"liboracle|siafd_oracle|oracle_lib|orc_siafd|/oracle/", text):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_null_handle_is_an_error_not_a_crash():
    """Every entry point tolerates the NULL handle a failed siafd_b200_create leaves behind: status calls return
    ERR_BAD_ARGUMENT, getters -1 / NaN / NULL; nothing touches the GPU."""
    L = capi.lib
    out8, i64 = (C.c_double * 8)(), C.c_int64()
    assert L.siafd_b200_update(None, None, None, 1) == capi.ERR_BAD_ARGUMENT
    assert b"NULL handle" in L.siafd_b200_last_error(None)
    for st in (L.siafd_b200_upload(None, 0, None), L.siafd_b200_download(None, 0, None), L.siafd_b200_finish(None),
               L.siafd_b200_compute_gradient(None), L.siafd_b200_compute_flux_velocity(None, 1, 0.0),
               L.siafd_b200_compute_vertical_velocity(None, 0, 0), L.siafd_b200_cfl(None, 1.0, 1, out8),
               L.siafd_b200_mass_flow_step(None, 1.0), L.siafd_b200_ensure_consistency(None, 1),
               L.siafd_b200_surface_values(None, capi.F["u"], None), L.siafd_b200_set_stream(None, None),
               L.siafd_b200_transfer_bytes(None, C.byref(i64), C.byref(i64)), L.siafd_b200_bind(None, 0, None)):
        assert st == capi.ERR_BAD_ARGUMENT
    assert L.siafd_b200_field_size(None, 0) == -1 and L.siafd_b200_launch_count(None) == -1
    assert L.siafd_b200_device_ptr(None, 0) is None
    assert np.isnan(L.siafd_b200_max_diffusivity(None))
    L.siafd_b200_destroy(None)
