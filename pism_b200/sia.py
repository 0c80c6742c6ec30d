"""Host-side mirror of the reference's SSB_Modifier / SIAFD interface over the C ABI.

Same method names, argument meaning and error behaviour as
  * stressbalance::SSB_Modifier   src/stressbalance/SSB_Modifier.hh:39-72
  * stressbalance::SIAFD          src/stressbalance/sia/SIAFD.hh:50-131, SIAFD.cc:42-155
  * stressbalance::Inputs         src/stressbalance/StressBalance.hh:41-65
  * Geometry                      src/geometry/Geometry.cc:30-42, :121-187
so that tests read like the reference's own (test/miscellaneous.py:321-363, siafd_test.cc).
Fields are numpy arrays (host path: every update copies in and out, like a drop-in under
PISM would) or torch CUDA tensors (device-resident path), always in PISM's local ghosted
layout.  All arithmetic happens in libsiafd_b200.so; nothing here computes physics.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import F, lib


class PISMRuntimeError(RuntimeError):
    """pism::RuntimeError (src/util/error_handling.hh:47-68); .status is the C ABI code."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    """Raw address of a numpy array or torch tensor (must be float64, C-contiguous)."""
    if a is None:
        return None
    if _is_torch(a):
        import torch
        assert a.dtype == torch.float64 and a.is_contiguous()
        return a.data_ptr()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _as_pd(a):
    p = _ptr(a)
    return C.cast(C.c_void_p(p), C.POINTER(C.c_double)) if p else None


class Geometry:
    """The four 2D fields SIAFD reads from pism::Geometry, ghost width 2 (Geometry.cc:36-42)."""

    def __init__(self, bed_elevation, ice_thickness, ice_surface_elevation=None, cell_type=None,
                 sea_level_elevation=None):
        self.bed_elevation = bed_elevation
        self.ice_thickness = ice_thickness
        self.ice_surface_elevation = ice_surface_elevation
        self.cell_type = cell_type
        self.sea_level_elevation = sea_level_elevation


class Inputs:
    """stressbalance::Inputs (StressBalance.hh:41-65): only the members SIAFD reads."""

    def __init__(self, geometry=None, enthalpy=None, age=None, new_bed_elevation=True, no_model_mask=None,
                 no_model_surface_elevation=None):
        self.geometry = geometry
        self.new_bed_elevation = new_bed_elevation  # default true, StressBalance.cc:40
        self.enthalpy = enthalpy
        self.age = age
        # regional models only (StressBalance.hh:59-62)
        self.no_model_mask = no_model_mask
        self.no_model_surface_elevation = no_model_surface_elevation


class SSB_Modifier:
    """Owned outputs and getters of SSB_Modifier (SSB_Modifier.cc:30-91)."""

    def __init__(self, grid, patch):
        self.grid, self.patch = grid, patch
        self.m_D_max = 0.0
        self.m_diffusive_flux = None
        self.m_u = None
        self.m_v = None

    def init(self):
        pass

    def diffusive_flux(self):
        return self.m_diffusive_flux

    def max_diffusivity(self):
        return self.m_D_max

    def velocity_u(self):
        return self.m_u

    def velocity_v(self):
        return self.m_v


class SIAFD(SSB_Modifier):
    """stressbalance::SIAFD on a B200 (SIAFD.cc:42-155)."""

    def __init__(self, grid, config=None, patch=None, device=-1, current_time=0.0, global_bed=None, **overrides):
        patch = patch or grid.whole()
        super().__init__(grid, patch)
        cfg = config if config is not None else capi.default_config()
        for k, v in overrides.items():
            if k == "flow_law" and isinstance(v, str):
                v = capi.FLOW_LAWS[v]
            if k == "gradient_method" and isinstance(v, str):
                v = capi.GRADIENTS[v]
            if not hasattr(cfg, k):
                raise AttributeError("unknown configuration parameter %s" % k)
            setattr(cfg, k, v)
        cfg.Mx, cfg.My, cfg.Mz = grid.Mx, grid.My, grid.Mz
        cfg.xs, cfg.xm, cfg.ys, cfg.ym = patch.xs, patch.xm, patch.ys, patch.ym
        cfg.dx, cfg.dy = grid.dx, grid.dy
        self._z = np.ascontiguousarray(grid.z, dtype=np.float64)
        cfg.z = self._z.ctypes.data_as(C.POINTER(C.c_double))
        self.config = cfg
        self.current_time = current_time
        self._global_bed = global_bed
        self._device = device
        self._h = C.c_void_p()
        status = lib.siafd_b200_create(C.byref(cfg), device, C.byref(self._h))
        if status != capi.OK:
            raise PISMRuntimeError(status, lib.siafd_b200_last_error(None).decode())
        self._host_out = {}
        self._bed_smoother_ready = False

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib.siafd_b200_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- helpers ----------------------------------------------------------------------------
    def _check(self, status):
        if status != capi.OK:
            raise PISMRuntimeError(status, lib.siafd_b200_last_error(self._h).decode())

    def field_shape(self, name):
        w = lib.siafd_b200_field_width(self._h, F[name])
        dof = lib.siafd_b200_field_dof(self._h, F[name])
        s = (self.patch.ym + 2 * w, self.patch.xm + 2 * w)
        return s if dof == 1 else s + (dof,)

    def _host_array(self, name):
        a = self._host_out.get(name)
        if a is None:
            a = np.zeros(self.field_shape(name), dtype=np.float64)
            self._host_out[name] = a
        return a

    def download(self, name):
        """Copy any device field (incl. scratch such as thk_smooth, theta) to a fresh numpy array."""
        a = np.empty(self.field_shape(name), dtype=np.float64)
        self._check(lib.siafd_b200_download(self._h, F[name], a.ctypes.data))
        return a

    def upload(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.field_shape(name), (name, a.shape, self.field_shape(name))
        self._check(lib.siafd_b200_upload(self._h, F[name], a.ctypes.data))

    @property
    def handle(self):
        return self._h

    def launch_count(self):
        return lib.siafd_b200_launch_count(self._h)

    def transfer_bytes(self):
        """(host -> device, device -> host) bytes the host-array calls of this solver have moved since it was made."""
        up, dn = C.c_int64(0), C.c_int64(0)
        self._check(lib.siafd_b200_transfer_bytes(self._h, C.byref(up), C.byref(dn)))
        return up.value, dn.value

    def set_tuning(self, rows_per_cta=0, use_bulk_copy=-1, skip_ice_free_rows=-1):
        self._check(lib.siafd_b200_set_tuning(self._h, rows_per_cta, use_bulk_copy, skip_ice_free_rows))

    # -- the SSB_Modifier interface ------------------------------------------------------------
    def init(self):
        """SIAFD::init (SIAFD.cc:98-118): nothing to allocate lazily here; kept for interface parity."""
        super().init()

    def preprocess_bed(self, global_bed):
        """BedSmoother::preprocess_bed (BedSmoother.cc:99-153) from the global [My, Mx] bed."""
        g = np.ascontiguousarray(global_bed, dtype=np.float64)
        assert g.shape == (self.grid.My, self.grid.Mx)
        self._check(lib.siafd_b200_preprocess_bed(self._h, g.ctypes.data))
        self._bed_smoother_ready = True

    def update(self, sliding_velocity, inputs, full_update):
        """SIAFD::update(sliding_velocity, inputs, full_update), SIAFD.cc:122-155."""
        geo = inputs.geometry
        if inputs.new_bed_elevation and self.config.smoother_range > 0.0:  # SIAFD.cc:130-134
            if self._global_bed is None:
                raise PISMRuntimeError(capi.ERR_BAD_CONFIG,
                                       "bed smoother is on: pass global_bed= to SIAFD() or call preprocess_bed()")
            self.preprocess_bed(self._global_bed)
        device = _is_torch(inputs.enthalpy) and inputs.enthalpy.is_cuda
        cin = capi.Inputs()
        cin.surface, cin.thickness = _as_pd(geo.ice_surface_elevation), _as_pd(geo.ice_thickness)
        cin.mask, cin.bed = _as_pd(geo.cell_type), _as_pd(geo.bed_elevation)
        cin.enthalpy, cin.age = _as_pd(inputs.enthalpy), _as_pd(inputs.age)
        cin.sliding = _as_pd(sliding_velocity)
        cin.current_time = self.current_time
        cin.memory_space = 1 if device else 0
        cin.ghosts_valid = 1
        self._keep = (geo, inputs, sliding_velocity)  # keep arrays alive across the call
        cout = capi.Outputs()
        names = ["h_x", "h_y", "D", "flux"] + (["u", "v"] if full_update else [])
        if device:
            import torch
            for n in names + ["u", "v"]:
                if n not in self._host_out:
                    self._host_out[n] = torch.zeros(self.field_shape(n), dtype=torch.float64,
                                                    device=inputs.enthalpy.device)
            arrs = {n: self._host_out[n] for n in names}
        else:
            arrs = {n: self._host_array(n) for n in names}
        for n in names:
            setattr(cout, n, _as_pd(arrs[n]))
        cout.memory_space = 1 if device else 0
        status = lib.siafd_b200_update(self._h, C.byref(cin), C.byref(cout), 1 if full_update else 0)
        self.m_D_max = lib.siafd_b200_max_diffusivity(self._h)
        self.m_h_x, self.m_h_y, self.m_D = arrs["h_x"], arrs["h_y"], arrs["D"]
        self.m_diffusive_flux = arrs["flux"]
        if full_update:
            self.m_u, self.m_v = arrs["u"], arrs["v"]
        self._check(status)

    # -- the next consumer of u, v (a StressBalance member in the reference) --------------
    def compute_vertical_velocity(self, basal_melt_rate=None, upstream=False):
        """StressBalance::compute_vertical_velocity (StressBalance.cc:283-424) from the u, v, mask of the last
        full update (still resident on the device, ghosts valid).  basal_melt_rate: owned points [ym, xm] or
        None.  Returns w on the owned points, [ym, xm, Mz] (WITHOUT_GHOSTS, StressBalance.cc:142)."""
        if basal_melt_rate is not None:
            self.upload("basal_melt", basal_melt_rate)
        self._check(lib.siafd_b200_compute_vertical_velocity(self._h, 1 if upstream else 0,
                                                             0 if basal_melt_rate is None else 1))
        self.m_w = self.download("w")
        return self.m_w

    # -- SURVEY.md 8(f) N3 (a StressBalance member in the reference) -------------------------------------------
    def compute_volumetric_strain_heating(self, flow_law="gpbld", glen_exponent=3.0, enhancement_factor=1.0):
        """StressBalance::compute_volumetric_strain_heating (StressBalance.cc:426-642) from the thickness, mask,
        enthalpy, u, v of the last full update (resident on the device).  flow_law / glen_exponent /
        enhancement_factor: the SHALLOW stress balance's (`stress_balance.ssa.*`), as in the reference.
        Returns Sigma on the owned points, [ym, xm, Mz]."""
        law = capi.FLOW_LAWS[flow_law] if isinstance(flow_law, str) else int(flow_law)
        self._check(lib.siafd_b200_compute_strain_heating(self._h, law, glen_exponent, enhancement_factor))
        self._check(lib.siafd_b200_finish(self._h))
        self.m_strain_heating = self.download("strain_heating")
        return self.m_strain_heating

    # -- reads of the 3D outputs (util/iceModelVec3.cc:153-240) ------------------------------------------------
    def _value_at_height(self, name, z=None):
        import torch
        dev = self._device if self._device >= 0 else torch.cuda.current_device()
        out = torch.empty((self.patch.ym, self.patch.xm), dtype=torch.float64, device="cuda:%d" % dev)
        if z is None:
            self._check(lib.siafd_b200_surface_values(self._h, F[name], out.data_ptr()))
        else:
            self._check(lib.siafd_b200_hor_slice(self._h, F[name], float(z), out.data_ptr()))
        torch.cuda.synchronize(dev)  # the kernel ran on the handle's stream, not torch's
        return out.cpu().numpy()

    def getSurfaceValues(self, name):
        """IceModelVec3::getSurfaceValues (iceModelVec3.cc:226-240): the 3D field `name` ("u", "v", "w", "enthalpy",
        "age", "strain_heating") at z = ice thickness of the last update, on the owned points [ym, xm]; evaluated
        on the device from the resident fields (two values per column are read, nothing 3D crosses PCIe)."""
        return self._value_at_height(name)

    def getHorSlice(self, name, z):
        """IceModelVec3::getHorSlice (iceModelVec3.cc:209-223): the field at the constant height z."""
        return self._value_at_height(name, z)

    def volumetric_strain_heating(self):
        return getattr(self, "m_strain_heating", None)

    def velocity_w(self):
        return getattr(self, "m_w", None)

    # SIAFD.cc:963-977
    def surface_gradient_x(self):
        return self.m_h_x

    def surface_gradient_y(self):
        return self.m_h_y

    def diffusivity(self):
        return self.m_D

    def high_diffusivity_count(self):
        return lib.siafd_b200_high_diffusivity_count(self._h)


class SIAFD_Regional(SIAFD):
    """stressbalance::SIAFD_Regional (src/regional/SIAFD_Regional.cc): SIAFD whose surface gradient next to
    `no_model_mask` cells is that of the stored `no_model_surface_elevation`.  Single rank,
    host arrays; runs the update through the split form of the C ABI because the override sits between the
    gradient and the flux (SIAFD.cc:137-141)."""

    def init(self):
        super().init()

    def update(self, sliding_velocity, inputs, full_update):
        geo = inputs.geometry
        if inputs.no_model_mask is None or inputs.no_model_surface_elevation is None:
            raise PISMRuntimeError(capi.ERR_BAD_ARGUMENT, "SIAFD_Regional needs no_model_mask and "
                                                          "no_model_surface_elevation (StressBalance.hh:59-62)")
        if inputs.new_bed_elevation and self.config.smoother_range > 0.0:
            self.preprocess_bed(self._global_bed)
        h = self._h
        for name, a in (("surface", geo.ice_surface_elevation), ("thickness", geo.ice_thickness),
                        ("mask", geo.cell_type), ("bed", geo.bed_elevation), ("enthalpy", inputs.enthalpy),
                        ("age", inputs.age), ("sliding", sliding_velocity), ("no_model_mask", inputs.no_model_mask),
                        ("no_model_surface", inputs.no_model_surface_elevation)):
            if a is not None:
                self.upload(name, a)

        def wrap(*names):
            ids = (C.c_int * len(names))(*[F[n] for n in names])
            self._check(lib.siafd_b200_wrap_ghosts_many(h, len(names), ids))

        self._check(lib.siafd_b200_compute_gradient(h))            # SIAFD::compute_surface_gradient
        wrap("h_x", "h_y")
        self._check(lib.siafd_b200_compute_gradient_no_model(h))   # SIAFD_Regional.cc:52-55
        wrap("h_x_no_model", "h_y_no_model")
        self._check(lib.siafd_b200_apply_no_model_gradient(h))     # SIAFD_Regional.cc:63-116
        self._check(lib.siafd_b200_compute_flux_velocity(h, 1 if full_update else 0, self.current_time))
        if full_update:
            wrap("u", "v")
        status = lib.siafd_b200_finish(h)
        self.m_D_max = lib.siafd_b200_max_diffusivity(h)
        self.m_h_x, self.m_h_y, self.m_D = self.download("h_x"), self.download("h_y"), self.download("D")
        self.m_diffusive_flux = self.download("flux")
        if full_update:
            self.m_u, self.m_v = self.download("u"), self.download("v")
        self._check(status)


def computeSIASurfaceVelocities(grid, thk, topg, enthalpy, siasolver=SIAFD, config=None, **overrides):
    """PISM.sia.computeSIASurfaceVelocities (site-packages/PISM/sia.py:24-74): surface horizontal velocities of the
    SIA with zero basal sliding.  `thk`, `topg`: ghosted [ym + 2w, xm + 2w] arrays (w = w_geom, ghosts valid, the
    reference's md.vecs.thk / .topg); `enthalpy`: ghosted [.., Mz].  Sea level 0 and Geometry::ensure_consistency
    with geometry.ice_free_thickness_standard (the configuration's), as there (sia.py:40-45), by the library's mask
    kernel.
    Returns (u_surface, v_surface) on the owned points (the reference's vel_sia components)."""
    import torch
    thk = np.ascontiguousarray(thk, dtype=np.float64)
    topg = np.ascontiguousarray(topg, dtype=np.float64)
    if "global_bed" not in overrides and "patch" not in overrides:
        # one patch = the whole domain: the bed smoother's global bed (BedSmoother.cc:99-153) is topg's interior
        wb = (topg.shape[0] - grid.My) // 2
        overrides["global_bed"] = topg[wb:topg.shape[0] - wb, wb:topg.shape[1] - wb]
    sia = siasolver(grid, config=config, **overrides)
    sia.init()
    dev = "cuda:%d" % torch.cuda.current_device()
    d_thk, d_bed = torch.from_numpy(thk).to(dev), torch.from_numpy(topg).to(dev)
    d_sea = torch.zeros_like(d_thk)
    d_mask, d_surf = torch.empty_like(d_thk), torch.empty_like(d_thk)
    torch.cuda.synchronize()  # the copies above went through torch's stream, the kernel runs on the handle's
    sia._check(lib.siafd_b200_geometry_compute(sia.handle, thk.size, d_sea.data_ptr(), d_bed.data_ptr(),
                                               d_thk.data_ptr(), d_mask.data_ptr(), d_surf.data_ptr()))
    torch.cuda.synchronize()
    geometry = Geometry(topg, thk, d_surf.cpu().numpy(), d_mask.cpu().numpy())
    inputs = Inputs(geometry, np.ascontiguousarray(enthalpy, dtype=np.float64), None)
    ws = sia.config.w_sliding
    sliding = np.zeros((sia.patch.ym + 2 * ws, sia.patch.xm + 2 * ws, 2))
    sia.update(sliding, inputs, True)
    return sia.getSurfaceValues("u"), sia.getSurfaceValues("v")
