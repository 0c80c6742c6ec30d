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

KNOBS = ("SIAFD_B200_PIPELINE", "SIAFD_B200_BAND", "SIAFD_B200_ROWS", "SIAFD_B200_SPARSE", "SIAFD_B200_LEVEL_CUT",
         "SIAFD_B200_CUT_COLS", "SIAFD_B200_CUT_ROWS", "SIAFD_B200_REPL_THREADS", "SIAFD_B200_FILL_THREADS",
         "SIAFD_B200_ZERO_COPY")


@pytest.fixture
def knobs():
    saved = {k: os.environ.get(k) for k in KNOBS}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _sliding(grid, cfg, inputs):
    """A sliding velocity that is nonzero under the ice and over part of the ice-free ground (periodic in x, y)."""
    w = cfg.w_sliding
    jj, ii = np.meshgrid(np.arange(-w, grid.My + w) % grid.My, np.arange(-w, grid.Mx + w) % grid.Mx, indexing="ij")
    s = np.zeros(inputs["sliding"].shape)
    s[..., 0] = 3e-7 * np.sin(2 * np.pi * ii / grid.Mx) * (jj > grid.My // 3)
    s[..., 1] = -2e-7 * np.cos(2 * np.pi * jj / grid.My) * (ii < 2 * grid.Mx // 3)
    return s


def _run(name, sliding=False, pinned=False, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["SIAFD_B200_" + k] = str(v)
    grid, cfg, inputs, gb = cases.case(name)
    if sliding:
        inputs["sliding"] = _sliding(grid, cfg, inputs)
    sia = U.make_sia(grid, cfg, gb)
    keep = []
    if pinned:  # u, v in pinned (hence device-mapped) host memory, poisoned: every cell has to be written by the call
        import torch
        for n in ("u", "v"):
            t = torch.full(sia.field_shape(n), float("nan"), dtype=torch.float64).pin_memory()
            keep.append(t)
            sia._host_out[n] = t.numpy()
    U.gpu_update(sia, inputs, True)
    out = {k: np.array(v, copy=True) for k, v in (("u", sia.velocity_u()), ("v", sia.velocity_v()),
                                                  ("D", sia.diffusivity()), ("Q", sia.diffusive_flux()))}
    out["D_max"] = sia.max_diffusivity()
    out["bytes"] = sia.transfer_bytes()
    out["launches"] = sia.launch_count()
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


@pytest.mark.parametrize("name", ["Fs", "dome_96_31", "dome_96_31_rough", "dome_64_31_quadratic", "C4s_nosmooth"])
def test_level_cut_of_the_sparse_host_update_is_bit_identical(name, knobs):
    """The level cut (include/siafd_b200.h, siafd_b200_host_levels_needed): only the levels up to the thickest ice
    next to a column cross PCIe, the host replicates the top value of u, v.  Same bits as the plain form and as the
    pipeline without the cut, with and without sliding, for several chunk widths; and it does move fewer bytes."""
    for sliding in (False, True):
        plain = _run(name, sliding, PIPELINE=0)
        uncut = _run(name, sliding, PIPELINE=1, ROWS=16, LEVEL_CUT=0)
        settings = [dict(ROWS=16), dict(ROWS=16, CUT_COLS=8, REPL_THREADS=1), dict(ROWS=8, BAND=2, CUT_COLS=16, REPL_THREADS=3),
                    dict(ROWS=32, CUT_COLS=1000), dict(ROWS=16, LEVEL_CUT=2, FILL_THREADS=1), dict(ROWS=16, CUT_ROWS=1, CUT_COLS=16),
                    dict(ROWS=32, CUT_ROWS=5, CUT_COLS=32)]
        for env in settings:
            got = _run(name, sliding, PIPELINE=1, **env)
            for k in ("u", "v", "D", "Q"):
                assert np.array_equal(got[k], plain[k]), (name, sliding, env, k)
            assert got["D_max"] == plain["D_max"], (name, env)
            if env.get("ROWS") == 16 and "BAND" not in env:  # (the same bands as `uncut`)
                assert got["bytes"][0] < uncut["bytes"][0] and got["bytes"][1] < uncut["bytes"][1], (name, env, got["bytes"], uncut["bytes"])


def test_level_cut_is_off_with_the_bed_smoother(knobs):
    """thk_smooth = usurf - topgsmooth can exceed H by the roughness of the bed (BedSmoother.cc:306-320): no cut."""
    a = _run("C4s", PIPELINE=1, ROWS=16, LEVEL_CUT=1)
    b = _run("C4s", PIPELINE=1, ROWS=16, LEVEL_CUT=0)
    assert a["bytes"] == b["bytes"]
    for k in ("u", "v", "D", "Q"):
        assert np.array_equal(a[k], b[k])


@pytest.mark.parametrize("name", ["Fs", "dome_96_31_rough", "C4s_nosmooth", "C4s"])
def test_zero_copy_stores_into_pinned_host_arrays_are_bit_identical(name, knobs):
    """SIAFD_B200_ZERO_COPY=1: with u, v in pinned host memory a kernel stores the pieces there itself (k_store_pieces)
    instead of the copy engine's strided copies -- same bits, with and without the level cut, for coarse and fine pieces;
    pageable arrays keep the copies."""
    for sliding in (False, True):
        plain = _run(name, sliding, PIPELINE=0)
        copies = _run(name, sliding, pinned=True, ROWS=16)
        settings = [dict(ROWS=16), dict(ROWS=16, LEVEL_CUT=0), dict(ROWS=16, CUT_COLS=8, CUT_ROWS=1), dict(ROWS=8, BAND=2, CUT_COLS=24, CUT_ROWS=3),
                    dict(ROWS=32, SPARSE=0)]
        for env in settings:
            got = _run(name, sliding, pinned=True, ZERO_COPY=1, **env)
            for k in ("u", "v", "D", "Q"):
                assert np.array_equal(got[k], plain[k]), (name, sliding, env, k)
            assert got["D_max"] == plain["D_max"], (name, env)
        zc = _run(name, sliding, pinned=True, ZERO_COPY=1, ROWS=16)
        assert zc["launches"] > copies["launches"] and zc["bytes"] == copies["bytes"]
        pageable = _run(name, sliding, ZERO_COPY=1, ROWS=16)  # numpy arrays: the copy engine
        assert pageable["launches"] == copies["launches"]
        for k in ("u", "v"):
            assert np.array_equal(pageable[k], plain[k])
