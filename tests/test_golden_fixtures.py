"""Frozen vectors (tests/golden/oracle_fixture_*.npz, made by tools/make_golden.py).

CPU: the oracle must still reproduce them bit for bit (a change in the oracle shows up here).
GPU: the CUDA path, through the C ABI, on the stored inputs against the stored outputs -- no oracle
involved at run time; tolerance 1e-10 relative (max-norm) as BASELINE.json's north_star states,
gradients and thk_smooth bit-exact."""
import glob
import os

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "oracle_fixture_*.npz")))
NAMES = [os.path.basename(f)[len("oracle_fixture_"):-4] for f in FIXTURES]


def load(name):
    d = np.load(os.path.join(HERE, "golden", "oracle_fixture_%s.npz" % name))
    inputs = {k[3:]: d[k] for k in d.files if k.startswith("in_")}
    outs = {k[4:]: d[k] for k in d.files if k.startswith("out_")}
    return inputs, outs, (d["global_bed"] if "global_bed" in d.files else None)


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_fixture_bitwise(name):
    grid, cfg, _, _ = cases.case(name)
    inputs, outs, gb = load(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert run.status == 0
    for k in ("h_x", "h_y", "D", "Q", "u", "v", "work2d_0", "work2d_1"):
        assert np.array_equal(run.a[k], outs[k]), k
    assert run.D_max == float(outs["D_max"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_fixture(name):
    import gpu_util as U
    grid, cfg, _, _ = cases.case(name)
    inputs, outs, gb = load(name)
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    assert np.array_equal(sia.surface_gradient_x(), outs["h_x"])
    assert np.array_equal(sia.surface_gradient_y(), outs["h_y"])
    assert np.array_equal(sia.download("thk_smooth"), outs["work2d_0"])
    for got, want in ((sia.diffusivity(), outs["D"]), (sia.diffusive_flux(), outs["Q"]),
                      (sia.velocity_u(), outs["u"]), (sia.velocity_v(), outs["v"])):
        assert cases.rel_max(got, want) <= 1e-10
    assert abs(sia.max_diffusivity() - float(outs["D_max"])) <= 1e-10 * float(outs["D_max"])


# ---- SURVEY 8(f) rows on the frozen C4s update (tests/golden/oracle_consumers_C4s.npz, tools/make_golden.py) -------
def _consumers():
    return np.load(os.path.join(HERE, "golden", "oracle_consumers_C4s.npz"))


def test_oracle_reproduces_consumer_fixture_bitwise():
    import subprocess
    import sys
    import tempfile
    d = _consumers()
    # regenerate into a scratch copy of the tree's golden directory and compare array by array
    sys.path.insert(0, os.path.join(HERE, "..", "tools"))
    import make_golden
    old_root = make_golden.ROOT
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "tests", "golden"))
        make_golden.ROOT = tmp
        try:
            make_golden.consumers()
        finally:
            make_golden.ROOT = old_root
        new = np.load(os.path.join(tmp, "tests", "golden", "oracle_consumers_C4s.npz"))
        for k in d.files:
            assert np.array_equal(d[k], new[k]), k


@pytest.mark.gpu
def test_gpu_consumers_match_fixture():
    """w, CFL scalars, strain heating, flow step, masks, source step on the B200 against committed vectors: no oracle at
    run time.  Thickness, divergence, masks bit-exact when fed the stored flux; w, Sigma within 1e-10."""
    import ctypes as C
    import gpu_util as U
    from pism_b200.capi import lib
    d = _consumers()
    grid, cfg, _, _ = cases.case("C4s")
    inputs, outs, gb = load("C4s")
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    for f, k in (("flux", "Q"), ("u", "u"), ("v", "v")):  # the stored update, so that later differences are the consumers'
        sia.upload(f, outs[k])
    w = sia.compute_vertical_velocity()
    assert cases.rel_max(w, d["w"]) <= 1e-10
    sia.upload("w", d["w"])
    out = (C.c_double * 8)()
    sia._check(lib.siafd_b200_cfl(sia.handle, float(d["max_dt"]), 1, out))
    assert list(out[0:4]) == list(d["cfl3d"]) and list(out[4:7]) == list(d["cfl2d"][0:3])
    sig = sia.compute_volumetric_strain_heating("gpbld", 3.0, 1.0)
    assert cases.rel_max(sig, d["sigma"]) <= 1e-10
    sia._check(lib.siafd_b200_mass_flow_step(sia.handle, float(d["dt"])))
    wg = cfg.w_geom
    assert np.array_equal(sia.download("flux_div"), d["flux_div"])
    assert np.array_equal(sia.download("thk_change"), d["thk_change"])
    assert np.array_equal(sia.download("thickness")[wg:-wg, wg:-wg], d["H_after_flow"])
    sia._check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
    assert np.array_equal(sia.download("mask"), d["mask_after_flow"])
    sia.upload("smb", d["smb"])
    sia._check(lib.siafd_b200_mass_source_step(sia.handle, float(d["dt2"]), 910.0, 0))
    assert np.array_equal(sia.download("thickness")[wg:-wg, wg:-wg], d["H_after_source"])


# ---- reads of the 3D outputs and the spreading-disc flow steps (tools/make_golden.py::reads_and_transport) ---------
def test_oracle_reproduces_reads_and_mass_transport_fixtures_bitwise():
    import oracle_lib as O
    grid, cfg, _, _ = cases.case("C4s")
    inputs, outs, gb = load("C4s")
    d = np.load(os.path.join(HERE, "golden", "oracle_reads_C4s.npz"))
    p = cfg.oracle_params(grid)
    H, wg = inputs["thickness"], cfg.w_geom
    assert np.array_equal(O.value_at_height(p, outs["u"], cfg.w_uv, H, wg), d["u_surface"])
    assert np.array_equal(O.value_at_height(p, outs["v"], cfg.w_uv, H, wg), d["v_surface"])
    assert np.array_equal(O.value_at_height(p, inputs["enthalpy"], cfg.w_3d_in, z0=float(d["z_slice"])),
                          d["enthalpy_slice"])
    assert np.abs(d["u_surface"]).max() > 0

    from test_gpu_mass_transport import oracle_flow_steps, spreading_disc_setup
    S = spreading_disc_setup(51)
    oracle_flow_steps(S, 12)
    m = np.load(os.path.join(HERE, "golden", "oracle_mass_transport_51.npz"))
    assert np.array_equal(S["H"], m["thickness"]) and np.array_equal(S["mask"], m["mask"])
    assert S["dt"] == float(m["dt_last"])
