"""The CPU-oracle backend of pism_b200.icemodel.IceCompModel -- TEST INFRASTRUCTURE ONLY.

Drives the reference's time step (pism_b200/icemodel.py, host logic only) with the oracle's restatements of
SIAFD::update, compute_vertical_velocity, the CFL reductions and GeometryEvolution, so that the oracle can be pinned
against the golden output of `pismv -test C` (test/regression/test_15.sh)."""
import ctypes as C

import numpy as np

import cases
import oracle_lib as O
from pism_b200 import grid as G


class OracleBackend:
    """One rank's patch of the run on the CPU oracle; with `ranks` (a pism_b200.icemodel.Ranks of a torch.distributed
    run) ghosts move through pism_b200.halo.HaloExchanger (gloo) and the scalars are reduced over ranks."""

    def __init__(self, grid, cfg, inputs, max_dt_seconds, ice_density=910.0, ranks=None):
        from pism_b200 import icemodel
        self.grid, self.cfg = grid, cfg
        self.ranks = ranks or icemodel.Ranks(grid)
        self.p = cfg.oracle_params(grid, self.ranks.patch if self.ranks.size > 1 else None)
        self.run = O.Run(self.p, inputs)
        self.a = self.run.a
        for k in ("thickness", "surface", "mask"):
            assert self.a[k].flags["WRITEABLE"]
        self.max_dt, self.ice_density = max_dt_seconds, ice_density
        self.w = cfg.w_geom
        n = (self.p.ym, self.p.xm)
        self.wv = np.zeros(n + (grid.Mz,))
        self.divQ, self.dH, self.cons = np.zeros(n), np.zeros(n), np.zeros(n)
        self.eff_smb, self.eff_bmb = np.zeros(n), np.zeros(n)
        self._cfl3d = max_dt_seconds
        self.ex = None
        if self.ranks.size > 1:
            from pism_b200.halo import HaloExchanger
            self.ex = HaloExchanger(self.ranks.patch, group=self.ranks.group)

    def _ghosts(self, name, w):
        if self.ex is None:
            G.wrap_ghosts(self.a[name], w)
        else:
            import torch
            self.ex.exchange(name, torch.from_numpy(self.a[name]), w)

    def set_thickness(self, H_global):
        self.a["thickness"][...] = G.global_to_local(np.ascontiguousarray(H_global), self.ranks.patch, self.w)

    def thickness(self):
        w = self.w
        return self.ranks.gather_owned(self.a["thickness"][w:-w, w:-w])

    def ensure_consistency(self):
        """Geometry::ensure_consistency (Geometry.cc:121-187): mask and surface from H, then ghosts."""
        a, w = self.a, self.w
        self._ghosts("thickness", w)
        sea = np.zeros_like(a["thickness"])
        O.lib().orc_geometry_compute(C.byref(self.p), a["thickness"].size, O.dptr(sea), O.dptr(a["bed"]),
                                     O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["surface"]))
        # smoother off: topgsmooth is the ghosted copy of the bed (BedSmoother.cc:101-109); the bed is constant here

    def stress_balance_update(self, full_update):
        L, p, a, R = O.lib(), self.p, self.a, self.ranks
        assert self.run.gradient() == 0                       # SIAFD.cc:137
        self._ghosts("h_x", 1), self._ghosts("h_y", 1)         # :498-499
        st = self.run.flux_velocity(full_update)              # :141-153
        assert st in (0, 5), st                               # (5 = D_max above the limit on this rank: checked globally)
        out = (C.c_double * 4)()
        if full_update:
            self._ghosts("u", 1), self._ghosts("v", 1)         # :946-947
            st = L.orc_vertical_velocity(C.byref(p), O.dptr(a["mask"]), O.dptr(a["u"]), O.dptr(a["v"]), None, 0,
                                         O.dptr(self.wv))
            assert st == 0
            st = L.orc_cfl_3d(C.byref(p), self.max_dt, O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["u"]),
                              O.dptr(a["v"]), O.dptr(self.wv), out)
            assert st == 0
            self._cfl3d = R.global_min(out[0])
        st = L.orc_cfl_2d(C.byref(p), self.max_dt, O.dptr(a["mask"]), O.dptr(a["sliding"]), out)
        assert st == 0
        D_max = R.global_max(self.run.D_max)
        assert D_max <= self.cfg.D_limit
        return dict(D_max=D_max, cfl3d_dt=self._cfl3d, cfl2d_dt=R.global_min(out[0]))

    def flow_step(self, dt):
        a = self.a
        vel = O.dptr(a["sliding"]) if self.p.w_sliding >= 1 else None
        st = O.lib().orc_mass_flow_step(C.byref(self.p), dt, None, O.dptr(a["bed"]), O.dptr(a["thickness"]), vel,
                                        None, None, O.dptr(a["Q"]), O.dptr(self.divQ), O.dptr(self.dH),
                                        O.dptr(self.cons))
        assert st == 0, st

    def source_step(self, dt, smb_global):
        a = self.a
        smb = self.ranks.owned(np.asarray(smb_global, dtype=np.float64))
        st = O.lib().orc_mass_source_step(C.byref(self.p), dt, self.ice_density, 0, O.dptr(a["thickness"]),
                                          O.dptr(a["mask"]), None, O.dptr(smb), None, O.dptr(self.eff_smb),
                                          O.dptr(self.eff_bmb))
        assert st == 0, st


def pismv_model(testname, M, start_year=0.0, run_length_years=5000.0, max_dt_years=60.0, backend_factory=None,
                ranks=None):
    """`pismv -test B|C|L -Mx M -My M -Mz 31 -ys .. -y .. [-max_dt ..]` on the oracle (or, with backend_factory(grid,
    cfg, inputs, max_dt_seconds[, ranks]), on the GPU).  Test B shares test C's set-up except the domain half-width
    (pismv.cc:88-102).  `ranks`: a pism_b200.icemodel.Ranks for a multi-rank run (this rank gets its patch)."""
    from pism_b200 import icemodel
    grid, cfg, _, _ = cases.case("C1_%d" % M)
    if testname == "B":
        grid = G.Grid(M, M, 31, 1200e3, 1200e3, 4000.0, spacing="quadratic")
    patch = ranks.patch if ranks is not None and ranks.size > 1 else None
    if testname == "L":  # pismv.cc:103-110: the 1800 km box of tests F, G, L
        grid = G.Grid(M, M, 31, 900e3, 900e3, 4000.0, spacing="quadratic")
        from pism_b200 import synthetic as S, verification as V
        inputs = cases.to_numpy(S.test_C_state(grid, patch or grid.whole(), cfg))
    else:
        _, _, inputs, _ = cases.case("C1_%d" % M, patch=patch)
    inputs = {k: np.array(v, dtype=np.float64, copy=True) for k, v in inputs.items()}
    if testname == "L":  # initTestL (iceCompModel.cc:372-423): the bed of exactL; the thickness follows in initialize_2d
        _, bed, _ = V.exactL(V.radius(grid))
        inputs["bed"] = np.ascontiguousarray(G.global_to_local(bed, patch or grid.whole(), cfg.w_geom))
    max_dt = max_dt_years * icemodel.SECONDS_PER_YEAR_UDUNITS
    kw = {} if ranks is None else {"ranks": ranks}
    backend = (backend_factory or OracleBackend)(grid, cfg, inputs, max_dt, **kw)
    return icemodel.IceCompModel(backend, grid, testname, start_year, run_length_years, max_dt_years)


# the reference's golden output: test/regression/test_15.sh:19-28 (pismv -test C -Mbz 1 -Mz 31 -y 5000, Mx = My = 31, 41)
import json as _json
import os as _os

with open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "reference_kats.json")) as _f:
    _kats = _json.load(_f)
    TEST_15_GOLDEN = {int(k): v for k, v in _kats["pismv_test_C"]["rows"].items()}
    # test/regression/test_16.sh:17-28 (pismv -test L -Mbz 1 -Mz 31 -y 1000, Mx = My = 21, 31)
    TEST_16_GOLDEN = {int(k): v for k, v in _kats["pismv_test_L"]["rows"].items()}
