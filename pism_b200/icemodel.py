"""Host-side mirror of the reference's time step around the SIAFD path, for SIA-only verification runs.

SURVEY.md 8(f) N1: the caller of `SIAFD::update` and the consumer of its outputs (`diffusive_flux`,
`max_diffusivity`, `velocity_u/v`).  This is the *logic* of
  * IceModel::run / IceModel::step             src/icemodel/IceModel.cc:749-800, :388-640
  * IceModel::max_timestep (+ _diffusivity)    src/icemodel/timestepping.cc:52-68, :111-232
  * Time (calendar "none", pismv.cc:51) and Time::step   src/util/Time.cc:117-131, :153-170, :206-215
  * IceCompModel for tests B / C               src/verification/iceCompModel.cc:60-135, :301-356, :442-583
  * surface::Verification::update_ABCDH        src/verification/PSVerification.cc:165-224
restricted to what `pismv -test C` (BASELINE configs[0]) exercises: SIA stress balance with zero sliding, no
energy / age / bed deformation / calving, explicit mass continuity.  Every array operation is delegated to a
*backend* object (the device backend below drives libsiafd_b200.so through the C ABI and keeps all state in HBM;
the tests drive the same loop with the CPU oracle), so nothing in this file computes physics on fields.

Backend protocol (whole-domain, single rank):
    set_thickness(H_owned[My, Mx]); thickness() -> H_owned; ensure_consistency()
    stress_balance_update(full_update) -> dict(D_max=..., cfl3d_dt=..., cfl2d_dt=...)   [max_dt passed at init]
    flow_step(dt); source_step(dt, smb_owned[My, Mx] in kg m-2 s-1)
"""
import ctypes as C

import numpy as np

from . import verification

SECONDS_PER_YEAR_UDUNITS = 365.242198781 * 86400.0  # UDUNITS-2 "year": convert(sys, 1, "year", "seconds")
SECONDS_PER_YEAR_365_DAY = 365.0 * 86400.0          # Time.cc:163-164, the default "365_day" calendar
ICE_DENSITY = 910.0                                  # constants.ice.density, pism_config.cdl


class Time:
    """pism::Time (Time.cc:117-215).  Year length by calendar (Time.cc:153-170): "365_day" (the configuration default)
    or "none" (what pismv sets, pismv.cc:51: the UDUNITS year)."""

    def __init__(self, start_year=0.0, run_length_years=1000.0, calendar="none"):
        assert calendar in ("none", "365_day")
        self.year_length = SECONDS_PER_YEAR_365_DAY if calendar == "365_day" else SECONDS_PER_YEAR_UDUNITS
        self.m_run_start = self.years_to_seconds(start_year)
        self.m_run_end = self.years_to_seconds(start_year + run_length_years)
        self.m_time_in_seconds = self.m_run_start

    def years_to_seconds(self, y):
        return y * self.year_length

    def seconds_to_years(self, s):
        return s / self.year_length

    def current(self):
        return self.m_time_in_seconds

    def end(self):
        return self.m_run_end

    def step(self, delta_t):
        """Time::step, Time.cc:206-215."""
        self.m_time_in_seconds += delta_t
        if self.m_run_end > self.m_time_in_seconds and self.m_run_end - self.m_time_in_seconds < 1e-3:
            self.m_time_in_seconds = self.m_run_end


class IceCompModel:
    """`pismv -test B|C` as far as the SIAFD path and its mass-continuity consumer go."""

    def __init__(self, backend, grid, testname="C", start_year=0.0, run_length_years=1000.0, max_dt_years=60.0,
                 adaptive_ratio=0.12):
        assert testname in ("B", "C")
        self.backend, self.grid, self.testname = backend, grid, testname
        self.time = Time(start_year, run_length_years)
        # config->get_number("time_stepping.maximum_time_step", "seconds") converts with UDUNITS
        self.max_dt = max_dt_years * SECONDS_PER_YEAR_UDUNITS
        self.adaptive_ratio = adaptive_ratio
        self.m_dt = 0.0
        self.m_adaptive_timestep_reason = "$"
        self.steps = 0
        self.dt_history = []
        self.r = verification.radius(grid)
        self.initialize_2d()

    # iceCompModel.cc:301-356
    def exact(self, t):
        if self.testname == "C":
            return verification.exactC(t, self.r)
        return verification.exactB(t, self.r)

    def initialize_2d(self):
        H, _ = self.exact(self.time.current())
        self.backend.set_thickness(np.ascontiguousarray(H))

    # timestepping.cc:52-68 and :111-232 (hit_multiples = 0, skip off, no reporting restrictions)
    def max_timestep(self, sb):
        restrictions = []
        # submodels: the (inactive) energy model still restricts dt by the 3D CFL, EnergyModel.cc:314-324
        restrictions.append((sb["cfl3d_dt"], "energy"))
        if self.max_dt > 0.0:
            restrictions.append((self.max_dt, "max"))
        time_to_end = self.time.end() - self.time.current()
        if time_to_end > 0.0:
            restrictions.append((time_to_end, "end of the run"))
        restrictions.append((sb["cfl2d_dt"], "2D CFL"))
        D_max = sb["D_max"]
        if D_max > 0.0:
            dx, dy = self.grid.dx, self.grid.dy
            grid_factor = 1.0 / (dx * dx) + 1.0 / (dy * dy)
            restrictions.append((self.adaptive_ratio * 2.0 / (D_max * grid_factor), "diffusivity"))
        else:
            restrictions.append((self.max_dt, "max time step"))
        restrictions.sort(key=lambda r: r[0])
        self.m_adaptive_timestep_reason = "%s (overrides %s)" % (restrictions[0][1], restrictions[1][1])
        return restrictions[0][0]

    # PSVerification.cc:165-224: accumulation [m s-1] at time t, then scale(ice_density)
    def surface_mass_flux(self, t):
        _, M = self.exact(t)
        return np.ascontiguousarray(M * ICE_DENSITY)

    # IceModel.cc:388-640
    def step(self):
        current_time = self.time.current()
        sb = self.backend.stress_balance_update(True)
        self.m_dt = self.max_timestep(sb)
        self.backend.flow_step(self.m_dt)          # flow_step + apply_flux_divergence
        self.backend.ensure_consistency()          # enforce_consistency_of_geometry(DONT_REMOVE_ICEBERGS)
        self.backend.source_step(self.m_dt, self.surface_mass_flux(current_time))
        self.backend.ensure_consistency()          # enforce_consistency_of_geometry(REMOVE_ICEBERGS)
        self.time.step(self.m_dt)
        self.steps += 1
        self.dt_history.append(self.m_dt)

    # IceModel.cc:749-800
    def run(self, max_steps=None):
        self.backend.ensure_consistency()
        while self.time.current() < self.time.end():
            self.step()
            if max_steps is not None and self.steps >= max_steps:
                break

    # iceCompModel.cc:442-583 and :661-667
    def geometry_errors(self):
        """(prcntVOL, maxH, avH, relmaxETA) exactly as `reportErrors` prints them."""
        H = self.backend.thickness()
        Hexact, _ = self.exact(self.time.current())
        g = self.grid
        a = g.dx * g.dy * 1e-3 * 1e-3
        m = (2.0 * 3.0 + 2.0) / 3.0
        vol = volexact = 0.0
        Herr = avHerr = etaerr = 0.0
        domeHexact = 0.0
        for j in range(g.My):          # Points(grid) order: j outer, i inner (sums are order-dependent)
            for i in range(g.Mx):
                h, he = float(H[j, i]), float(Hexact[j, i])
                if h > 0:
                    vol += a * h * 1e-3
                if he > 0:
                    volexact += a * he * 1e-3
                if i == (g.Mx - 1) // 2 and j == (g.My - 1) // 2:
                    domeHexact = he
                Herr = max(Herr, abs(h - he))
                etaerr = max(etaerr, abs(h ** m - he ** m))
                avHerr += abs(h - he)
        volerr = abs(vol - volexact)
        return (100 * volerr / volexact, Herr, avHerr / (g.Mx * g.My), etaerr / domeHexact ** m)

    def report(self):
        return "%12.6f%12.6f%12.6f%12.6f" % self.geometry_errors()


class DeviceBackend:
    """The backend that runs on the B200: all fields stay in the handle's device buffers; per step only D_max,
    the CFL scalars and the (2D) surface mass flux cross PCIe.  `sia` is a pism_b200.sia.SIAFD."""

    def __init__(self, sia, inputs, max_dt_seconds, ice_density=ICE_DENSITY):
        from .capi import F, lib
        self.sia, self.lib, self.F = sia, lib, F
        self.max_dt, self.ice_density = max_dt_seconds, ice_density
        self.w = sia.config.w_geom
        for name in ("bed", "thickness", "enthalpy", "sliding"):
            sia.upload(name, inputs[name])
        self._wrapH = (C.c_int * 1)(F["thickness"])

    def _check(self, status):
        self.sia._check(status)

    def set_thickness(self, H_owned):
        g, w = self.sia.grid, self.w
        a = np.zeros((g.My + 2 * w, g.Mx + 2 * w))
        a[w:-w, w:-w] = H_owned
        self.sia.upload("thickness", a)

    def thickness(self):
        w = self.w
        return self.sia.download("thickness")[w:-w, w:-w]

    def ensure_consistency(self):
        self._check(self.lib.siafd_b200_ensure_consistency(self.sia.handle, 1))

    def stress_balance_update(self, full_update):
        h, lib = self.sia.handle, self.lib
        self._check(lib.siafd_b200_compute_gradient(h))
        names = (C.c_int * 2)(self.F["h_x"], self.F["h_y"])
        self._check(lib.siafd_b200_wrap_ghosts_many(h, 2, names))
        self._check(lib.siafd_b200_compute_flux_velocity(h, 1 if full_update else 0, self.sia.current_time))
        out = (C.c_double * 8)()
        if full_update:
            names = (C.c_int * 2)(self.F["u"], self.F["v"])
            self._check(lib.siafd_b200_wrap_ghosts_many(h, 2, names))
            self._check(lib.siafd_b200_compute_vertical_velocity(h, 0, 0))
        self._check(lib.siafd_b200_cfl(h, self.max_dt, 1 if full_update else 0, out))
        self._check(lib.siafd_b200_finish(h))
        if full_update:
            self._cfl3d = out[0]
        return dict(D_max=lib.siafd_b200_max_diffusivity(h), cfl3d_dt=self._cfl3d, cfl2d_dt=out[4])

    def flow_step(self, dt):
        self._check(self.lib.siafd_b200_mass_flow_step(self.sia.handle, dt))

    def source_step(self, dt, smb_owned):
        self.sia.upload("smb", smb_owned)
        self._check(self.lib.siafd_b200_mass_source_step(self.sia.handle, dt, self.ice_density, 0))
