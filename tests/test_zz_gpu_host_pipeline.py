"""-m gpu: the pipelined host-array form of siafd_b200_update (bands of row segments on three streams, sparse
transfers, host fill) gives the same bits as the plain upload / compute / download form, for every band length --
including the shapes where the last band holds no owned row (31 rows with 32-row segments) or a single band covers
everything.  (tools/e2e_check.py as a test; the knobs are read from the environment at siafd_b200_create.  Written
after round 1's GPU minutes were spent -- tools/e2e_check.py itself passed on C4s with these settings -- and therefore
placed last in collection order.)"""
import os

import numpy as np
import pytest

import cases
import gpu_util as U

pytestmark = pytest.mark.gpu

KNOBS = ("SIAFD_B200_PIPELINE", "SIAFD_B200_BAND", "SIAFD_B200_ROWS", "SIAFD_B200_SPARSE")


@pytest.fixture
def knobs():
    saved = {k: os.environ.get(k) for k in KNOBS}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _run(name, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["SIAFD_B200_" + k] = str(v)
    grid, cfg, inputs, gb = cases.case(name)
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    out = {k: np.array(v, copy=True) for k, v in (("u", sia.velocity_u()), ("v", sia.velocity_v()),
                                                  ("D", sia.diffusivity()), ("Q", sia.diffusive_flux()))}
    out["D_max"] = sia.max_diffusivity()
    return out


@pytest.mark.parametrize("name", ["Fs", "C4s", "dome_96_31"])
def test_pipelined_host_update_equals_plain(name, knobs):
    plain = _run(name, PIPELINE=0)
    settings = [dict(), dict(BAND=1, ROWS=32), dict(BAND=1, ROWS=16), dict(BAND=3, ROWS=8), dict(BAND=100, ROWS=64),
                dict(BAND=2, ROWS=32, SPARSE=0)]
    for env in settings:
        got = _run(name, PIPELINE=1, **env)
        for k in ("u", "v", "D", "Q"):
            assert np.array_equal(got[k], plain[k]), (name, env, k)
        assert got["D_max"] == plain["D_max"], (name, env)
