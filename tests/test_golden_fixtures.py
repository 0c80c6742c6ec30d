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
