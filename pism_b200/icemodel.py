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

Backend protocol (global [My, Mx] arrays in and out; a backend of a multi-rank run keeps its own patch of them):
    set_thickness(H); thickness() -> H; ensure_consistency()
    stress_balance_update(full_update) -> dict(D_max=..., cfl3d_dt=..., cfl2d_dt=...), reduced over ranks
    flow_step(dt); source_step(dt, smb in kg m-2 s-1)
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
    """`pismv -test B|C|L` as far as the SIAFD path and its mass-continuity consumer go.  Test L (steady state on a
    non-flat bed, iceCompModel.cc:372-423): the backend must have been given exactL's bed."""

    def __init__(self, backend, grid, testname="C", start_year=0.0, run_length_years=1000.0, max_dt_years=60.0,
                 adaptive_ratio=0.12):
        assert testname in ("B", "C", "L")
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
        if testname == "L":
            # initTestL (iceCompModel.cc:372-423): thickness from the ODE solution, kept as m_HexactL; accumulation of
            # Verification::update_L (PSVerification.cc:99-126): a0 converted by UDUNITS, not exactL's own SperA
            self.m_HexactL, _, _ = verification.exactL(self.r)
            a0 = 0.3 / SECONDS_PER_YEAR_UDUNITS
            self.m_mass_flux_L = a0 * (1.0 - (2.0 * self.r * self.r / (750e3 * 750e3)))
        self.initialize_2d()

    # iceCompModel.cc:301-356
    def exact(self, t):
        if self.testname == "C":
            return verification.exactC(t, self.r)
        if self.testname == "L":
            return self.m_HexactL, self.m_mass_flux_L
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


class Ranks:
    """What a backend needs from a parallel run (one process per GPU / patch): its patch of PISM's decomposition and
    the reductions the reference does with GlobalMax / GlobalMin (src/util/pism_utilities.cc:140-160).  Serial when
    torch.distributed is not initialised."""

    def __init__(self, grid, patches=None, rank=0, group=None):
        self.grid = grid
        self.patches = patches or [grid.whole()]
        self.rank, self.group = rank, group
        self.patch = self.patches[rank]
        self.size = len(self.patches)

    def _reduce(self, x, op):
        if self.size == 1:
            return x
        import torch
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op, group=self.group)
        return float(t.item())

    def global_max(self, x):
        import torch.distributed as dist
        return self._reduce(x, dist.ReduceOp.MAX) if self.size > 1 else x

    def global_min(self, x):
        import torch.distributed as dist
        return self._reduce(x, dist.ReduceOp.MIN) if self.size > 1 else x

    def gather_owned(self, owned):
        """Global [My, Mx] array from every rank's owned part (reports only: small grids)."""
        if self.size == 1:
            return owned
        import torch.distributed as dist
        parts = [None] * self.size
        dist.all_gather_object(parts, np.ascontiguousarray(owned), group=self.group)
        out = np.zeros((self.grid.My, self.grid.Mx))
        for pt, a in zip(self.patches, parts):
            out[pt.ys:pt.ys + pt.ym, pt.xs:pt.xs + pt.xm] = a
        return out

    def owned(self, a_global):
        p = self.patch
        return np.ascontiguousarray(a_global[p.ys:p.ys + p.ym, p.xs:p.xs + p.xm])


class DeviceBackend:
    """The backend that runs on the B200: all fields stay in the handle's device buffers; per step only D_max,
    the CFL scalars and the (2D) surface mass flux cross PCIe.  `sia` is a pism_b200.sia.SIAFD built for this rank's
    patch.  On several ranks (`ranks`, one process per GPU) the ghost updates where the reference has them
    (Geometry.cc:172, SIAFD.cc:498-499, :946-947) go through `halo` (pism_b200.halo.PeerHalo: direct stores into the
    neighbours' arrays over NVLink) and D_max / the CFL scalars are reduced over ranks."""

    def __init__(self, sia, inputs, max_dt_seconds, ice_density=ICE_DENSITY, ranks=None, halo=None):
        from .capi import F, lib
        self.sia, self.lib, self.F = sia, lib, F
        self.max_dt, self.ice_density = max_dt_seconds, ice_density
        self.w = sia.config.w_geom
        self.ranks = ranks or Ranks(sia.grid)
        self.halo = halo
        assert (self.ranks.size == 1) == (halo is None)
        for name in ("bed", "thickness", "enthalpy", "sliding"):
            sia.upload(name, inputs[name])  # this patch's local arrays, ghosts included

    def _check(self, status):
        self.sia._check(status)

    def _ghosts(self, names_widths, phase):
        """IceModelVec::update_ghosts of fields on the device."""
        if self.halo is not None:
            self.halo.exchange(names_widths, phase)
        else:
            ids = (C.c_int * len(names_widths))(*[self.F[n] for n, _ in names_widths])
            self._check(self.lib.siafd_b200_wrap_ghosts_many(self.sia.handle, len(names_widths), ids))

    def set_thickness(self, H_global):
        from . import grid as G
        self.sia.upload("thickness", G.global_to_local(np.ascontiguousarray(H_global), self.ranks.patch, self.w))

    def thickness(self):
        w = self.w
        return self.ranks.gather_owned(self.sia.download("thickness")[w:-w, w:-w])

    def ensure_consistency(self):
        self._ghosts([("thickness", self.w)], 0)
        self._check(self.lib.siafd_b200_ensure_consistency(self.sia.handle, 0))

    def stress_balance_update(self, full_update):
        h, lib = self.sia.handle, self.lib
        self._check(lib.siafd_b200_compute_gradient(h))
        self._ghosts([("h_x", 1), ("h_y", 1)], 1)
        self._check(lib.siafd_b200_compute_flux_velocity(h, 1 if full_update else 0, self.sia.current_time))
        out = (C.c_double * 8)()
        if full_update:
            self._ghosts([("u", 1), ("v", 1)], 2)
            self._check(lib.siafd_b200_compute_vertical_velocity(h, 0, 0))
        self._check(lib.siafd_b200_cfl(h, self.max_dt, 1 if full_update else 0, out))
        self._check(lib.siafd_b200_finish(h))
        R = self.ranks
        if full_update:
            self._cfl3d = R.global_min(out[0])
        return dict(D_max=R.global_max(lib.siafd_b200_max_diffusivity(h)), cfl3d_dt=self._cfl3d,
                    cfl2d_dt=R.global_min(out[4]))

    def flow_step(self, dt):
        self._check(self.lib.siafd_b200_mass_flow_step(self.sia.handle, dt))

    def source_step(self, dt, smb_global):
        self.sia.upload("smb", self.ranks.owned(smb_global))
        self._check(self.lib.siafd_b200_mass_source_step(self.sia.handle, dt, self.ice_density, 0))
