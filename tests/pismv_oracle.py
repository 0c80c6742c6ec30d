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
    def __init__(self, grid, cfg, inputs, max_dt_seconds, ice_density=910.0):
        self.grid, self.cfg = grid, cfg
        self.p = cfg.oracle_params(grid)
        self.run = O.Run(self.p, inputs)
        self.a = self.run.a
        for k in ("thickness", "surface", "mask"):
            assert self.a[k].flags["WRITEABLE"]
        self.max_dt, self.ice_density = max_dt_seconds, ice_density
        self.w = cfg.w_geom
        n = (grid.My, grid.Mx)
        self.wv = np.zeros(n + (grid.Mz,))
        self.divQ, self.dH, self.cons = np.zeros(n), np.zeros(n), np.zeros(n)
        self.eff_smb, self.eff_bmb = np.zeros(n), np.zeros(n)
        self._cfl3d = max_dt_seconds

    def set_thickness(self, H_owned):
        w = self.w
        self.a["thickness"][w:-w, w:-w] = H_owned
        G.wrap_ghosts(self.a["thickness"], w)

    def thickness(self):
        w = self.w
        return self.a["thickness"][w:-w, w:-w]

    def ensure_consistency(self):
        """Geometry::ensure_consistency (Geometry.cc:121-187): mask and surface from H, then ghosts."""
        a, w = self.a, self.w
        G.wrap_ghosts(a["thickness"], w)
        sea = np.zeros_like(a["thickness"])
        O.lib().orc_geometry_compute(C.byref(self.p), a["thickness"].size, O.dptr(sea), O.dptr(a["bed"]),
                                     O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["surface"]))
        # smoother off: topgsmooth is the ghosted copy of the bed (BedSmoother.cc:101-109); the bed is constant here

    def stress_balance_update(self, full_update):
        L, p, a = O.lib(), self.p, self.a
        st = self.run.update_single(full_update)
        assert st == 0, st
        out = (C.c_double * 4)()
        if full_update:
            st = L.orc_vertical_velocity(C.byref(p), O.dptr(a["mask"]), O.dptr(a["u"]), O.dptr(a["v"]), None, 0,
                                         O.dptr(self.wv))
            assert st == 0
            st = L.orc_cfl_3d(C.byref(p), self.max_dt, O.dptr(a["thickness"]), O.dptr(a["mask"]), O.dptr(a["u"]),
                              O.dptr(a["v"]), O.dptr(self.wv), out)
            assert st == 0
            self._cfl3d = out[0]
        st = L.orc_cfl_2d(C.byref(p), self.max_dt, O.dptr(a["mask"]), O.dptr(a["sliding"]), out)
        assert st == 0
        return dict(D_max=self.run.D_max, cfl3d_dt=self._cfl3d, cfl2d_dt=out[0])

    def flow_step(self, dt):
        a = self.a
        vel = O.dptr(a["sliding"]) if self.p.w_sliding >= 1 else None
        st = O.lib().orc_mass_flow_step(C.byref(self.p), dt, None, O.dptr(a["bed"]), O.dptr(a["thickness"]), vel,
                                        None, None, O.dptr(a["Q"]), O.dptr(self.divQ), O.dptr(self.dH),
                                        O.dptr(self.cons))
        assert st == 0, st

    def source_step(self, dt, smb_owned):
        a = self.a
        smb = np.ascontiguousarray(smb_owned, dtype=np.float64)
        st = O.lib().orc_mass_source_step(C.byref(self.p), dt, self.ice_density, 0, O.dptr(a["thickness"]),
                                          O.dptr(a["mask"]), None, O.dptr(smb), None, O.dptr(self.eff_smb),
                                          O.dptr(self.eff_bmb))
        assert st == 0, st


def pismv_model(testname, M, start_year=0.0, run_length_years=5000.0, max_dt_years=60.0, backend_factory=None):
    """`pismv -test B|C -Mx M -My M -Mz 31 -ys .. -y .. [-max_dt ..]` on the oracle (or, with backend_factory(grid,
    cfg, inputs, max_dt_seconds), on the GPU).  Test B shares test C's set-up except the domain half-width
    (pismv.cc:88-102)."""
    from pism_b200 import icemodel
    grid, cfg, inputs, _ = cases.case("C1_%d" % M)
    if testname == "B":
        grid = G.Grid(M, M, 31, 1200e3, 1200e3, 4000.0, spacing="quadratic")
    inputs = {k: np.array(v, dtype=np.float64, copy=True) for k, v in inputs.items()}
    max_dt = max_dt_years * icemodel.SECONDS_PER_YEAR_UDUNITS
    backend = (backend_factory or OracleBackend)(grid, cfg, inputs, max_dt)
    return icemodel.IceCompModel(backend, grid, testname, start_year, run_length_years, max_dt_years)


# the reference's golden output: test/regression/test_15.sh:19-28 (pismv -test C -Mbz 1 -Mz 31 -y 5000, Mx = My = 31, 41)
import json as _json
import os as _os

with open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "reference_kats.json")) as _f:
    TEST_15_GOLDEN = {int(k): v for k, v in _json.load(_f)["pismv_test_C"]["rows"].items()}
