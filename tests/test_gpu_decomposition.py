"""Decomposition independence of the CUDA path on ONE GPU (mirrors test/regression/test_02.sh): the domain is cut into
patches with PISM's rule and with unequal -procs_x / -procs_y ranges, every patch gets its own handle, ghost updates
are emulated through host memory where the reference has them, and every result -- SIAFD::update and its consumers
(vertical velocity, CFL scalars, strain heating, flow step, ensure_consistency, source step) -- must equal the
single-patch result BIT FOR BIT.  This is what exercises the kernels with xs, ys != 0 and xm < Mx."""
import ctypes as C

import numpy as np
import pytest

import cases
import gpu_util as U
from pism_b200 import grid as G
from pism_b200.capi import F, lib

pytestmark = pytest.mark.gpu

SEC = 365.242198781 * 86400.0


def _sliding(grid, w, amp):
    x = np.arange(-w, grid.Mx + w)[None, :] % grid.Mx
    y = np.arange(-w, grid.My + w)[:, None] % grid.My
    s = np.zeros((grid.My + 2 * w, grid.Mx + 2 * w, 2))
    s[..., 0] = amp * np.sin(2 * np.pi * x / grid.Mx) * np.cos(2 * np.pi * y / grid.My)
    s[..., 1] = -amp * np.cos(4 * np.pi * x / grid.Mx) * np.sin(2 * np.pi * y / grid.My)
    return s


class Ranks:
    """One handle per patch on the same device; `exchange` is IceModelVec::update_ghosts through host memory."""

    def __init__(self, grid, cfg, patches, inputs_global):
        self.grid, self.cfg, self.patches = grid, cfg, patches
        self.sias = [U.make_sia(grid, cfg, None, patch=pt) for pt in patches]
        for sia, pt in zip(self.sias, patches):
            for name in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding"):
                w = lib.siafd_b200_field_width(sia.handle, F[name])
                sia.upload(name, G.global_to_local(inputs_global[name], pt, w))

    def each(self, fn):
        return [fn(sia, pt) for sia, pt in zip(self.sias, self.patches)]

    def gather(self, name):
        """Global array of the owned values of a field."""
        out = None
        for sia, pt in zip(self.sias, self.patches):
            a = sia.download(name)
            w = lib.siafd_b200_field_width(sia.handle, F[name])
            own = cases.interior(a, w)
            if out is None:
                out = np.zeros((self.grid.My, self.grid.Mx) + own.shape[2:])
            out[pt.ys:pt.ys + pt.ym, pt.xs:pt.xs + pt.xm] = own
        return out

    def exchange(self, *names):
        for name in names:
            g = self.gather(name)
            for sia, pt in zip(self.sias, self.patches):
                w = lib.siafd_b200_field_width(sia.handle, F[name])
                sia.upload(name, G.global_to_local(g, pt, w))

    def check(self, st, sia):
        sia._check(st)


def _run(grid, cfg, patches, inputs_global, smb_global, dt, dt2):
    R = Ranks(grid, cfg, patches, inputs_global)
    res = {}
    R.each(lambda s, p: s._check(lib.siafd_b200_compute_gradient(s.handle)))
    R.exchange("h_x", "h_y")                                                            # SIAFD.cc:498-499
    R.each(lambda s, p: s._check(lib.siafd_b200_compute_flux_velocity(s.handle, 1, 0.0)))
    R.exchange("u", "v")                                                                # SIAFD.cc:946-947
    R.each(lambda s, p: s._check(lib.siafd_b200_finish(s.handle)))
    res["D_max"] = max(R.each(lambda s, p: lib.siafd_b200_max_diffusivity(s.handle)))  # SIAFD.cc:748
    R.each(lambda s, p: s._check(lib.siafd_b200_compute_vertical_velocity(s.handle, 0, 0)))
    cfl = []
    for s in R.sias:
        out = (C.c_double * 8)()
        s._check(lib.siafd_b200_cfl(s.handle, 60.0 * SEC, 1, out))
        cfl.append(list(out))
    cfl = np.array(cfl)
    res["cfl"] = [cfl[:, 0].min(), cfl[:, 1].max(), cfl[:, 2].max(), cfl[:, 3].max(), cfl[:, 4].min(), cfl[:, 5].max(),
                  cfl[:, 6].max()]
    R.each(lambda s, p: s._check(lib.siafd_b200_compute_strain_heating(s.handle, 2, 3.0, 1.0)))
    for name in ("h_x", "h_y", "D", "flux", "u", "v", "w", "strain_heating"):
        res[name] = R.gather(name)
    R.each(lambda s, p: s._check(lib.siafd_b200_mass_flow_step(s.handle, dt)))
    res["flux_div"], res["thk_change"] = R.gather("flux_div"), R.gather("thk_change")
    R.exchange("thickness")                                                             # Geometry.cc:172
    R.each(lambda s, p: s._check(lib.siafd_b200_ensure_consistency(s.handle, 0)))
    res["H_flow"], res["mask_flow"], res["surface_flow"] = R.gather("thickness"), R.gather("mask"), R.gather("surface")
    R.each(lambda s, p: s.upload("smb", np.ascontiguousarray(smb_global[p.ys:p.ys + p.ym, p.xs:p.xs + p.xm])))
    R.each(lambda s, p: s._check(lib.siafd_b200_mass_source_step(s.handle, dt2, 910.0, 0)))
    R.exchange("thickness")
    R.each(lambda s, p: s._check(lib.siafd_b200_ensure_consistency(s.handle, 0)))
    R.each(lambda s, p: s._check(lib.siafd_b200_finish(s.handle)))
    res["H_source"], res["mask_source"] = R.gather("thickness"), R.gather("mask")
    res["eff_smb"] = R.gather("eff_smb")
    return res


@pytest.mark.parametrize("name,decomp", [
    ("C4s_nosmooth", dict(size=4)),                                               # PISM's rule: 2 x 2 (61 x 113)
    ("C4s_nosmooth", dict(size=6, Nx=2, Ny=3, procs_x=[40, 21], procs_y=[50, 13, 50])),   # unequal ranges
    ("dome_64_21", dict(size=8)),                                                 # 2 x 4
    ("dome_35_101", dict(size=3, Nx=3, Ny=1, procs_x=[17, 3, 15])),               # a 3-column patch, Mz = 101
])
def test_patches_reproduce_the_single_patch_run_bitwise(name, decomp):
    grid, cfg, inputs, gb = cases.case(name)
    cfg.w_sliding = 1
    inputs = dict(inputs)
    inputs["sliding"] = _sliding(grid, 1, 2e-5)
    glob = {k: np.ascontiguousarray(cases.interior(np.asarray(v), (v.shape[0] - grid.My) // 2)) for k, v in inputs.items()
            if k in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding")}
    smb = (np.random.default_rng(1).random((grid.My, grid.Mx)) - 0.6) * 3e-3
    dt, dt2 = 0.5 * SEC, 150.0 * SEC
    one = _run(grid, cfg, [grid.whole()], glob, smb, dt, dt2)
    size = decomp.pop("size")
    patches = G.decompose(grid.Mx, grid.My, size, **decomp)
    many = _run(grid, cfg, patches, glob, smb, dt, dt2)
    assert many["D_max"] == one["D_max"]
    assert many["cfl"] == one["cfl"], (many["cfl"], one["cfl"])
    for k in one:
        if k in ("D_max", "cfl"):
            continue
        assert np.array_equal(many[k], one[k]), k
    assert np.abs(one["w"]).max() > 0 and one["strain_heating"].max() > 0 and np.abs(one["thk_change"]).max() > 0
