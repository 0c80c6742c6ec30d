"""Helpers shared by the -m gpu tests: run the CUDA path through the C ABI and compare with the oracle."""
import ctypes as C

import numpy as np

import cases
from pism_b200 import capi
from pism_b200.sia import SIAFD, Geometry, Inputs

TOL = 1e-10  # BASELINE.json north_star: u / v / D within 1e-10 relative (max-norm)


def make_sia(grid, cfg, gb=None, patch=None, current_time=0.0):
    return SIAFD(grid, patch=patch, global_bed=gb, current_time=current_time, **cfg.overrides())


def gpu_update(sia, inputs, full=True):
    geo = Geometry(inputs["bed"], inputs["thickness"], inputs["surface"], inputs["mask"])
    sia.update(inputs["sliding"], Inputs(geo, inputs["enthalpy"], inputs.get("age")), full)
    return sia


def compare_with_oracle(sia, run, cfg, full, exact_gradient=True, report=None):
    """Assert parity of one whole-domain update (ghost rings included) and return the measured errors."""
    errs = {}
    hx, hy = sia.surface_gradient_x(), sia.surface_gradient_y()
    if exact_gradient:
        assert np.array_equal(hx, run.a["h_x"]), "h_x not bit-identical"
        assert np.array_equal(hy, run.a["h_y"]), "h_y not bit-identical"
        errs["h_x"] = errs["h_y"] = 0.0
    else:
        errs["h_x"], errs["h_y"] = cases.rel_max(hx, run.a["h_x"]), cases.rel_max(hy, run.a["h_y"])
    assert np.array_equal(sia.download("thk_smooth"), run.a["work2d_0"]), "thk_smooth not bit-identical"
    errs["D"] = cases.rel_max(sia.diffusivity(), run.a["D"])
    errs["flux"] = cases.rel_max(sia.diffusive_flux(), run.a["Q"])
    errs["D_max"] = abs(sia.max_diffusivity() - run.D_max) / max(run.D_max, 1e-300)
    if full:
        errs["u"] = cases.rel_max(sia.velocity_u(), run.a["u"])
        errs["v"] = cases.rel_max(sia.velocity_v(), run.a["v"])
    if report is not None:
        report.update(errs)
    for k, e in errs.items():
        assert e <= TOL, (k, e)
    # exact zeros must stay exact zeros (ice-free columns, edge override): same sparsity pattern
    assert np.array_equal(sia.diffusivity() == 0.0, run.a["D"] == 0.0)
    return errs
