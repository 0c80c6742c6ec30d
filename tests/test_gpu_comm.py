"""The decomposed update (siafd_b200_update_decomposed + the siafd_b200_comm_* communicator) on ONE GPU: the domain is cut
into patches, every patch gets its own handle, the handles form a same-process communicator (siafd_b200_comm_init_local)
and run the whole step -- input ghosts, gradient with the fused h_x / h_y ghost update (SIAFD.cc:498-499), the fused
kernel with the u / v ghost update (:946-947), the all-rank reduction of D_max / error bits / counter (:748-750) -- with
no host help.  Every owned AND ghost value must equal the single-patch result bit for bit (test/regression/test_02.sh),
a negative thickness on one rank must fail EVERY rank (ParallelSection, util/error_handling.cc:189-214), and the generic
ghost update / all-reduce must do what IceModelVec::update_ghosts / GlobalMax do."""
import ctypes as C

import numpy as np
import pytest

import cases
import gpu_util as U
from pism_b200 import capi
from pism_b200 import grid as G
from pism_b200.capi import F, lib

pytestmark = pytest.mark.gpu

INPUTS = ("surface", "thickness", "mask", "bed", "enthalpy", "sliding")


def _global(grid, inputs):
    return {k: np.ascontiguousarray(cases.interior(np.asarray(v), (v.shape[0] - grid.My) // 2)) for k, v in inputs.items()
            if k in INPUTS}


def _handles(grid, cfg, patches, glob, poison_ghosts):
    sias = [U.make_sia(grid, cfg, None, patch=pt) for pt in patches]
    for sia, pt in zip(sias, patches):
        for name in INPUTS:
            w = lib.siafd_b200_field_width(sia.handle, F[name])
            a = G.global_to_local(glob[name], pt, w)
            if poison_ghosts and w > 0 and name != "sliding":
                own = cases.interior(a, w).copy()
                a = np.full_like(a, np.nan)
                a[w:-w, w:-w] = own
            sia.upload(name, a)
    hs = (C.c_void_p * len(sias))(*[s.handle for s in sias])
    st = lib.siafd_b200_comm_init_local(hs, len(sias))
    assert st == 0, lib.siafd_b200_last_error(sias[0].handle)
    return sias


def _local(glob_field, pt, w):
    return G.global_to_local(glob_field, pt, w)


def _gather(grid, sias, patches, name):
    out = None
    for sia, pt in zip(sias, patches):
        a = sia.download(name)
        w = lib.siafd_b200_field_width(sia.handle, F[name])
        own = cases.interior(a, w)
        if out is None:
            out = np.zeros((grid.My, grid.Mx) + own.shape[2:])
        out[pt.ys:pt.ys + pt.ym, pt.xs:pt.xs + pt.xm] = own
    return out


@pytest.mark.parametrize("name,decomp,full,graph", [
    ("C4s_nosmooth", dict(size=1), True, 1),                                      # one rank: periodic self-wrap
    ("C4s_nosmooth", dict(size=4), True, 1),                                      # PISM's rule: 2 x 2 (61 x 113)
    ("C4s_nosmooth", dict(size=6, Nx=2, Ny=3, procs_x=[40, 21], procs_y=[50, 13, 50]), True, 0),
    ("dome_64_21", dict(size=8), True, 1),                                        # 2 x 4
    ("dome_35_101", dict(size=3, Nx=3, Ny=1, procs_x=[17, 3, 15]), True, 1),      # a 3-column patch, Mz = 101
    ("dome_64_21", dict(size=2), False, 1),                                       # full_update = false
    ("C1_mahaffy", dict(size=4), True, 1),                                        # no mid exchange (mahaffy)
])
def test_decomposed_update_equals_single_patch_bitwise(name, decomp, full, graph, monkeypatch):
    monkeypatch.setenv("SIAFD_B200_GRAPH", str(graph))
    grid, cfg, inputs, gb = cases.case(name)
    cfg.w_sliding = 1
    inputs = dict(inputs)
    rng = np.random.default_rng(3)
    sl = np.zeros((grid.My, grid.Mx, 2))
    sl[...] = 1e-6 * rng.standard_normal(sl.shape)
    inputs["sliding"] = G.global_to_local(sl, grid.whole(), 1)
    glob = _global(grid, inputs)
    # reference: one patch, the split calls with periodic self-wraps
    one = U.make_sia(grid, cfg, None)
    U.gpu_update(one, inputs, full)
    ref = {k: one.download(k) for k in ("h_x", "h_y", "D", "flux") + (("u", "v") if full else ())}
    ref_dmax = one.max_diffusivity()

    size = decomp.pop("size")
    patches = G.decompose(grid.Mx, grid.My, size, **decomp)
    sias = _handles(grid, cfg, patches, glob, poison_ghosts=True)
    for rep in range(3):  # (the captured graph is replayed from the second call on)
        for s in sias:
            s._check(lib.siafd_b200_update_decomposed(s.handle, 1 if full else 0, 0.0, 1))
        for s in sias:
            s._check(lib.siafd_b200_finish(s.handle))
    for s, pt in zip(sias, patches):
        assert lib.siafd_b200_max_diffusivity(s.handle) == ref_dmax  # global max on every rank (SIAFD.cc:748)
        for k, r in ref.items():
            w = lib.siafd_b200_field_width(s.handle, F[k])
            got = s.download(k)
            glob_k = cases.interior(r, w)
            want = _local(np.ascontiguousarray(glob_k), pt, w)
            if k in ("D", "flux"):
                # computed locally on owned + 1 (SIAFD.cc:620, :781): compare the owned points and the ring; the
                # domain-edge override makes ghosts of an interior patch differ from the wrapped single-patch ones
                # only where the reference's would too, so compare owned points here
                assert np.array_equal(cases.interior(got, w), cases.interior(want, w)), (k, pt)
            else:
                assert np.array_equal(got, want), (k, pt)  # ghosts included: they came from the neighbours
        for k in ("surface", "thickness", "mask", "bed", "enthalpy"):  # input ghosts arrived too
            w = lib.siafd_b200_field_width(s.handle, F[k])
            assert np.array_equal(s.download(k), _local(glob[k], pt, w)), (k, pt)


def test_error_is_collective():
    """A negative thickness in ONE patch: every rank's finish returns the same status (the reference's ParallelSection)."""
    grid, cfg, inputs, gb = cases.case("dome_64_21")
    glob = _global(grid, dict(inputs, sliding=np.zeros((grid.My + 2, grid.Mx + 2, 2))))
    glob["thickness"] = glob["thickness"].copy()
    glob["thickness"][3, 5] = -1.0  # owned by rank 0 only, far from every other patch's ghosts
    cfg.w_sliding = 1
    patches = G.decompose(grid.Mx, grid.My, 8)
    sias = _handles(grid, cfg, patches, glob, poison_ghosts=False)
    for s in sias:
        assert lib.siafd_b200_update_decomposed(s.handle, 1, 0.0, 0) == 0
    st = [lib.siafd_b200_finish(s.handle) for s in sias]
    assert st == [capi.ERR_NEGATIVE_THICKNESS] * len(sias), st
    # and the flag is cleared once it has been reported: after repairing the input the next step is clean
    for s, pt in zip(sias, patches):
        t = glob["thickness"].copy()
        t[3, 5] = 0.0
        s.upload("thickness", G.global_to_local(t, pt, 2))
    for s in sias:
        assert lib.siafd_b200_update_decomposed(s.handle, 1, 0.0, 0) == 0
    assert [lib.siafd_b200_finish(s.handle) for s in sias] == [0] * len(sias)


def test_generic_exchange_and_allreduce():
    grid, cfg, inputs, gb = cases.case("dome_64_21")
    cfg.w_sliding = 1
    glob = _global(grid, dict(inputs, sliding=np.zeros((grid.My + 2, grid.Mx + 2, 2))))
    patches = G.decompose(grid.Mx, grid.My, 6, Nx=3, Ny=2, procs_x=[20, 30, 14], procs_y=[40, 24])
    sias = _handles(grid, cfg, patches, glob, poison_ghosts=True)
    fields = (C.c_int * 3)(F["thickness"], F["enthalpy"], F["mask"])
    for width in (1, 2):
        widths = (C.c_int * 3)(width, 2, width)
        for s in sias:
            assert lib.siafd_b200_comm_exchange(s.handle, 3, fields, widths) == 0, lib.siafd_b200_last_error(s.handle)
        for s in sias:
            s._check(lib.siafd_b200_finish(s.handle))
    for s, pt in zip(sias, patches):
        for k in ("thickness", "enthalpy", "mask"):
            assert np.array_equal(s.download(k), _local(glob[k], pt, 2)), (k, pt)
    # all-reduce: needs every rank in flight at once, so one host thread per rank (as one process per GPU would be)
    import threading
    out = [None] * len(sias)

    def work(q):
        v = (C.c_double * 3)(float(q + 1), -float(q), 0.5 * (q + 1))
        assert lib.siafd_b200_comm_allreduce(sias[q].handle, 0, 3, v) == 0
        w = (C.c_double * 2)(float(q + 1), 0.1 * (q + 1))
        assert lib.siafd_b200_comm_allreduce(sias[q].handle, 2, 2, w) == 0
        m = (C.c_double * 1)(float(q + 1))
        assert lib.siafd_b200_comm_allreduce(sias[q].handle, 1, 1, m) == 0
        out[q] = (list(v), list(w), list(m))

    ts = [threading.Thread(target=work, args=(q,)) for q in range(len(sias))]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    n = len(sias)
    s01 = 0.0
    for q in range(n):
        s01 += 0.1 * (q + 1)
    for q in range(n):
        assert out[q] == ([float(n), 0.0, 0.5 * n], [n * (n + 1) / 2.0, s01], [1.0]), out[q]


@pytest.mark.parametrize("name,decomp,full", [
    ("dome_64_21", dict(size=4), True),
    ("C4s_nosmooth", dict(size=6, Nx=2, Ny=3, procs_x=[40, 21], procs_y=[50, 13, 50]), True),
    ("C1", dict(size=2), True),
    ("dome_64_21", dict(size=2), False),
])
def test_host_arrays_through_update_on_every_rank(name, decomp, full):
    """siafd_b200_update with HOST arrays (valid ghosts, as PISM's are) on every rank of a decomposed run: the drop-in call
    of a multi-rank PISM.  One host thread per rank (a rank's call returns when its neighbours have delivered their
    ghosts).  Every output array must equal the single-rank call's, ghosts included."""
    import threading
    import torch
    from pism_b200.sia import Geometry, Inputs

    def pinned(a):
        # page-locked host arrays: a blocking copy from pageable memory holds a per-process staging lock, which would
        # serialise the ranks of this ONE-process test against each other (separate processes, as under MPI, do not share it)
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

    grid, cfg, inputs, gb = cases.case(name)
    cfg.w_sliding = 1
    inputs = dict(inputs)
    inputs["sliding"] = np.zeros((grid.My + 2, grid.Mx + 2, 2))
    glob = _global(grid, inputs)
    one = U.make_sia(grid, cfg, None)
    U.gpu_update(one, inputs, full)
    names = ("h_x", "h_y", "D", "flux") + (("u", "v") if full else ())
    ref = {k: one.download(k) for k in names}
    size = decomp.pop("size")
    patches = G.decompose(grid.Mx, grid.My, size, **decomp)
    sias = [U.make_sia(grid, cfg, None, patch=pt) for pt in patches]
    hs = (C.c_void_p * size)(*[s.handle for s in sias])
    assert lib.siafd_b200_comm_init_local(hs, size) == 0
    errs = [None] * size
    keep = []
    locs = []
    for s, pt in zip(sias, patches):
        loc = {}
        for k in INPUTS:
            loc[k], t = pinned(G.global_to_local(glob[k], pt, lib.siafd_b200_field_width(s.handle, F[k])))
            keep.append(t)
        for k in ("h_x", "h_y", "D", "flux", "u", "v"):
            s._host_out[k], t = pinned(np.zeros(s.field_shape(k)))
            keep.append(t)
        locs.append(loc)

    def work(q):
        try:
            s, loc = sias[q], locs[q]
            for rep in range(2):
                s.update(loc["sliding"], Inputs(Geometry(loc["bed"], loc["thickness"], loc["surface"], loc["mask"]), loc["enthalpy"]),
                         full)
        except Exception as e:  # noqa: BLE001
            errs[q] = e

    ts = [threading.Thread(target=work, args=(q,)) for q in range(size)]
    [t.start() for t in ts]
    [t.join(120) for t in ts]
    assert errs == [None] * size, errs
    for s, pt in zip(sias, patches):
        assert s.max_diffusivity() == one.max_diffusivity()
        got = {"h_x": s.surface_gradient_x(), "h_y": s.surface_gradient_y(), "D": s.diffusivity(), "flux": s.diffusive_flux()}
        if full:
            got["u"], got["v"] = s.velocity_u(), s.velocity_v()
        for k in names:
            w = lib.siafd_b200_field_width(s.handle, F[k])
            want = _local(np.ascontiguousarray(cases.interior(ref[k], w)), pt, w)
            if k in ("D", "flux"):
                assert np.array_equal(cases.interior(got[k], w), cases.interior(want, w)), (k, pt)
            else:
                assert np.array_equal(got[k], want), (k, pt)


def test_step_breakdown_times_the_sections_of_a_step():
    """siafd_b200_step_breakdown_ms (bench.py's roofline.step_breakdown_ms): in kernel-timing mode every ungraphed step of
    siafd_b200_update_decomposed leaves five section times; the fused kernel's section agrees with
    siafd_b200_kernel_time_ms, and reading the breakdown resets it."""
    grid, cfg, inputs, gb = cases.case("dome_64_21")
    glob = _global(grid, dict(inputs))
    patches = G.decompose(grid.Mx, grid.My, 2)
    sias = _handles(grid, cfg, patches, glob, poison_ghosts=False)
    for s in sias:
        s._check(lib.siafd_b200_kernel_timing(s.handle, 1))
    for rep in range(3):
        for s in sias:
            s._check(lib.siafd_b200_update_decomposed(s.handle, 1, 0.0, 1))
        for s in sias:
            s._check(lib.siafd_b200_finish(s.handle))
    for s in sias:
        sec, n = (C.c_double * 5)(), C.c_int(0)
        s._check(lib.siafd_b200_step_breakdown_ms(s.handle, sec, C.byref(n)))
        nk = C.c_int(0)
        k_ms = lib.siafd_b200_kernel_time_ms(s.handle, C.byref(nk))
        assert n.value == 3 and nk.value == 3
        assert all(v >= 0.0 for v in sec) and sec[3] > 0.0
        assert abs(sec[3] - k_ms / 3) <= 0.5 * sec[3] + 0.05  # the same kernel between neighbouring events
        s._check(lib.siafd_b200_step_breakdown_ms(s.handle, sec, C.byref(n)))
        assert n.value == 0
        s._check(lib.siafd_b200_kernel_timing(s.handle, 0))
