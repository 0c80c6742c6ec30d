"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle
on identical inputs.  Integer / mask / index work bit-exact; D, flux, u, v within 1e-10 relative
(max-norm), the tolerance BASELINE.json's north_star states."""
import ctypes as C

import numpy as np
import pytest

import cases
import gpu_util as U
import oracle_lib as O
from pism_b200 import capi, grid as G
from test_oracle_known_answers import FLOW_TABLE, bed_smoother_case, flow_table_inputs, ghosted

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


# name, full_update, gradient bit-exact?
CASES = [
    ("C1", True, True),               # pismv -test C 61x61x31 (BASELINE configs[0])
    ("C1_mahaffy", True, True),
    ("C1_eta", True, False),
    ("C2", True, True),               # pismv -test G 121x121x61 (configs[1])
    ("C3", True, True),               # EISMINT II F-shaped 151x151x101, pb (configs[2])
    ("C4", True, True),               # Greenland-shaped 301x561x101, gpbld, smoother (configs[3])
    ("C4s_limit", True, True),        # diffusivity cap active
    ("C4s_nosmooth", False, True),    # full_update = false path
    ("dome_96_31", True, True),
    ("dome_96_31", False, True),
    ("dome_80_101_pb", True, True),
    ("dome_64_31_hooke", True, True),
    ("dome_64_31_gk", True, True),
    ("dome_64_31_arrwarm", True, True),
    ("dome_64_31_n4", True, True),          # generic Glen exponent: pow() path
    ("dome_64_41_quadratic", True, True),   # unequal vertical spacing
    ("dome_64_31_rough_eta", True, False),  # eta gradient over a rough bed
    ("dome_33_13", True, True),             # ragged: strip/segment remainders, Mz < 16
    ("dome_50_17_mahaffy", True, True),
    ("dome_48_20", True, True),             # even Mz: padded shared-memory columns, 8-byte cp.async row loader
    ("dome_48_20", False, True),
    ("dome_24_401", True, True),            # tall grid: the 8-lane-column configuration of the fused kernel
]


@pytest.mark.parametrize("name,full,exact_grad", CASES)
def test_update_matches_oracle(name, full, exact_grad, record_property):
    grid, cfg, inputs, gb = cases.case(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=full)
    assert run.status == 0, "oracle status %d" % run.status
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, full)
    errs = U.compare_with_oracle(sia, run, cfg, full, exact_grad)
    for k, v in errs.items():
        record_property(k, v)
    print(name, "full" if full else "flux-only", {k: "%.2e" % v for k, v in errs.items()})
    if cfg.smoother_range > 0:
        assert cases.rel_max(sia.download("theta"), run.a["work2d_1"]) < 1e-14
        for f in ("topgsmooth", "maxtl", "C2", "C3", "C4"):
            assert np.array_equal(sia.download(f), run.a[f]), f   # preprocess_bed: bit-exact
    if cfg.limit_diffusivity:
        assert sia.high_diffusivity_count() == run.f.high_diffusivity_counter > 0
    if not full:
        assert sia.velocity_u() is None   # G10: u, v untouched


@pytest.mark.parametrize("name", ["C4", "dome_256"])
def test_pointwise_relative_error_of_the_velocities(name):
    """The north_star's bar is a MAX-NORM relative error (1e-10 of the largest |u|); that hides what happens at thin margins
    where |u| is a millionth of the maximum.  Here every point on its own: |u_gpu - u_oracle| / |u_oracle| over all points
    with |u_oracle| > 1e-12 max|u_oracle| (below that the values are rounding noise of the sums themselves), for the
    Greenland-shaped config at full size (301 x 561 x 101, all four mask values, bed smoother on) and a C5-shaped dome.
    Measured on the B200: C4 2.0e-10 at the worst point (where the four I h terms nearly cancel: its value is 1e-6 of the
    maximum), 2.6e-13 at the 99.99th percentile; dome 2.2e-14; D 1.9e-14.  Asserted at 1e-9."""
    grid, cfg, inputs, gb = cases.case(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert run.status == 0
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    for k, got in (("u", sia.velocity_u()), ("v", sia.velocity_v()), ("D", sia.diffusivity())):
        ref = run.a[k]
        sel = np.abs(ref) > 1e-12 * np.abs(ref).max()
        rel = np.abs(got - ref)[sel] / np.abs(ref)[sel]
        print("%s %s: pointwise relative error max %.2e, 99.99th percentile %.2e, over %d of %d points; max-norm %.2e" %
              (name, k, rel.max(), np.quantile(rel, 0.9999), sel.sum(), ref.size, cases.rel_max(got, ref)))
        assert rel.max() < 1e-9, (k, rel.max())
        assert np.all(got[~sel & (ref == 0.0)] == 0.0)  # exact zeros stay exact zeros


@pytest.mark.parametrize("name,w_sliding", [("dome_96_31", 1), ("C4s", 1), ("C4s_nosmooth", 0), ("dome_33_13", 2)])
def test_nonzero_sliding_velocity(name, w_sliding):
    """u = u_b - ... with a sliding velocity that differs at every point (the reference adds it for all Mz levels,
    SIAFD.cc:935-942), including the columns and rows without ice, where u, v are the sliding velocity alone."""
    grid, cfg, inputs, gb = cases.case(name)
    cfg.w_sliding = w_sliding
    My, Mx = grid.My, grid.Mx
    jj, ii = np.meshgrid(np.arange(-w_sliding, My + w_sliding), np.arange(-w_sliding, Mx + w_sliding), indexing="ij")
    sl = np.zeros((My + 2 * w_sliding, Mx + 2 * w_sliding, 2))
    sl[..., 0] = 1e-6 * (1.0 + np.sin(0.37 * ii) * np.cos(0.11 * jj)) + 1e-9 * (ii + 1000 * jj)
    sl[..., 1] = -2e-6 * np.cos(0.23 * ii + 0.05 * jj) + 1e-9 * (7 * ii - jj)
    inputs["sliding"] = np.ascontiguousarray(sl)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert run.status == 0
    for rows in (64, 5):   # long and short row segments: the lane-distributed sliding FIFO wraps several times
        sia = U.make_sia(grid, cfg, gb)
        sia.set_tuning(rows_per_cta=rows)
        U.gpu_update(sia, inputs, True)
        U.compare_with_oracle(sia, run, cfg, True)
        # columns without any ice around them carry exactly the sliding velocity
        u, ub = cases.interior(sia.velocity_u(), cfg.w_uv), cases.interior(sl[..., 0], w_sliding)
        free = cases.interior(sia.download("thk_smooth"), cfg.w_geom) == 0
        free &= np.roll(free, 1, 0) & np.roll(free, -1, 0) & np.roll(free, 1, 1) & np.roll(free, -1, 1)
        assert free.any() and np.array_equal(u[free], np.repeat(ub[free][:, None], grid.Mz, axis=1))


@pytest.mark.parametrize("name,upstream,melt", [("C2", False, False), ("C4s", False, True), ("C4s", True, True),
                                                ("dome_96_31", True, False), ("dome_33_13", False, True),
                                                ("dome_64_41_quadratic", False, False)])
def test_vertical_velocity_matches_oracle(name, upstream, melt):
    """StressBalance::compute_vertical_velocity (StressBalance.cc:283-424) on the u, v of a full
    update, against the oracle run on the oracle's own u, v: centered and upstream differences, one-sided at ice
    margins (all four mask values in C4s), with and without a basal melt rate, unequal dz."""
    grid, cfg, inputs, gb = cases.case(name)
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
    assert run.status == 0
    p = run.p
    for a in (run.a["u"], run.a["v"]):  # the ghost exchange of SIAFD.cc:946-947
        O.lib().orc_wrap_ghosts(grid.Mx, grid.My, cfg.w_uv, grid.Mz, O.dptr(a))
    bmr = None
    if melt:
        jj, ii = np.meshgrid(np.arange(grid.My), np.arange(grid.Mx), indexing="ij")
        bmr = np.ascontiguousarray(1e-9 * (1.0 + np.sin(0.3 * ii) * np.cos(0.2 * jj)))
    w_o = np.zeros((grid.My, grid.Mx, grid.Mz))
    st = O.lib().orc_vertical_velocity(C.byref(p), O.dptr(run.a["mask"]), O.dptr(run.a["u"]), O.dptr(run.a["v"]),
                                       O.dptr(bmr) if melt else None, 1 if upstream else 0, O.dptr(w_o))
    assert st == 0
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    w_g = sia.compute_vertical_velocity(bmr, upstream)
    assert np.max(np.abs(w_o)) > 0
    assert cases.rel_max(w_g, w_o) <= U.TOL
    assert np.array_equal(w_g[..., 0], -bmr if melt else np.zeros((grid.My, grid.Mx)))  # w at the base, exactly


def test_age_coupling_matches_oracle():
    """e_age_coupling and grain_size_age_coupling (SIAFD.cc:649-675) with the gk law."""
    grid, cfg, inputs, gb = cases.case("dome_64_31_gk")
    cfg.grain_size_age_coupling, cfg.e_age_coupling = 1, 1
    cfg.fl_e, cfg.fl_e_interglacial = 3.0, 1.0
    z = grid.z[None, None, :]
    secpera = 365.242198781 * 86400.0
    H = np.maximum(cases.interior(inputs["thickness"], 0), 1.0)[..., None]
    inputs["age"] = np.ascontiguousarray(150.0e3 * secpera * np.clip(1.0 - z / H, 0.0, 1.0) ** 2 +
                                         np.zeros_like(inputs["enthalpy"]))
    run = cases.oracle_run(grid, cfg, inputs, gb, full=True, current_time=0.0)
    assert run.status == 0
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    U.compare_with_oracle(sia, run, cfg, True)


@pytest.mark.parametrize("law", sorted(FLOW_TABLE))
def test_flow_law_known_answers_on_device(law):
    """flowlaw_test (test/miscellaneous.py:598-671) evaluated by the device flow laws."""
    p = O.default_params()
    p.flow_law = O.FLOW_LAWS[law]
    rows = np.array(flow_table_inputs(p))
    grid = G.Grid(16, 16, 11, 1e5, 1e5, 1000.0)
    sia = U.make_sia(grid, cases.Cfg(flow_law=law, smoother_range=0.0))
    dev = [torch.tensor(rows[:, c].copy(), dtype=torch.float64, device="cuda") for c in range(4)]
    out = torch.zeros(len(rows), dtype=torch.float64, device="cuda")
    st = capi.lib.siafd_b200_flow_n(sia.handle, len(rows), *[d.data_ptr() for d in dev], out.data_ptr())
    assert st == 0
    got, want = out.cpu().numpy(), np.array(FLOW_TABLE[law])
    assert np.max(np.abs(got - want)) < 1e-16
    assert np.max(np.abs(got - want) / want) < 1e-8
    ora = np.array([O.lib().orc_flow(C.byref(p), *r) for r in rows])
    assert np.max(np.abs(got - ora) / ora) < 1e-13


def test_cell_type_mask_bit_exact():
    """GeometryCalculator::compute (Mask.hh:96-133): masks and surface bit-identical to the oracle,
    including thickness exactly at the ice-free threshold and flotation ties."""
    rng = np.random.RandomState(0)
    n = 200000
    bed = rng.uniform(-1500.0, 1500.0, n)
    thk = np.where(rng.rand(n) < 0.3, 0.0, rng.uniform(0.0, 3000.0, n))
    thk[:1000] = 0.01           # == threshold -> ice free
    thk[1000:2000] = np.nextafter(0.01, 1.0)
    sea = rng.uniform(-50.0, 50.0, n)
    p = O.default_params()
    alpha = 1 - p.ec_rho_i / p.sea_water_density
    bed[2000:3000] = sea[2000:3000] + alpha * thk[2000:3000] - thk[2000:3000]   # flotation tie
    m_o, s_o = np.zeros(n), np.zeros(n)
    O.lib().orc_geometry_compute(C.byref(p), n, O.dptr(sea), O.dptr(bed), O.dptr(thk), O.dptr(m_o), O.dptr(s_o))
    sia = U.make_sia(G.Grid(16, 16, 11, 1e5, 1e5, 1000.0), cases.Cfg(smoother_range=0.0))
    d = [torch.tensor(a, device="cuda") for a in (sea, bed, thk)]
    m_g, s_g = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    assert capi.lib.siafd_b200_geometry_compute(sia.handle, n, *[t.data_ptr() for t in d], m_g.data_ptr(),
                                                s_g.data_ptr()) == 0
    assert np.array_equal(m_g.cpu().numpy(), m_o) and np.array_equal(s_g.cpu().numpy(), s_o)
    assert set(np.unique(m_o)) == {0.0, 2.0, 3.0, 4.0}


def test_bed_smoother_known_answer_on_device():
    """bed_smoother_test (test/bed_smoother.py:120-146) through siafd_b200_preprocess_bed + theta."""
    p, topg = bed_smoother_case()
    grid = G.Grid(81, 81, 3, 1200e3, 1200e3, 2000.0)
    cfg = cases.Cfg(smoother_range=50.0e3, flow_law="isothermal_glen", D_limit=1e30)
    sia = U.make_sia(grid, cfg, gb=topg)
    w = 2
    usurf = np.full((81 + 2 * w, 81 + 2 * w), 1000.0)
    inputs = dict(surface=usurf, thickness=usurf.copy(), mask=np.full_like(usurf, 2.0), bed=ghosted(topg, w),
                  enthalpy=np.zeros((85, 85, 3)), sliding=np.zeros((83, 83, 2)))
    U.gpu_update(sia, inputs, False)
    ts = sia.download("topgsmooth")[w:-w, w:-w]
    th = sia.download("theta")[w:-w, w:-w]
    assert abs(ts.min() + 372.9924735817933) < 1e-16 and abs(ts.max() - 372.9924735817933) < 1e-16
    # theta goes through pow(): allow 2 ulp on the reference's 1e-16
    assert abs(th.min() - 0.7147300652935706) < 3e-16 and abs(th.max() - 0.9884843647808601) < 3e-16


def test_error_conditions_are_reported_like_the_reference():
    grid, cfg, inputs, gb = cases.case("dome_33_13")
    from pism_b200.sia import PISMRuntimeError
    # negative thickness: BedSmoother.cc:303-305
    bad = dict(inputs)
    bad["thickness"] = inputs["thickness"].copy()
    bad["thickness"][10, 10] = -1.0
    sia = U.make_sia(grid, cfg, gb)
    with pytest.raises(PISMRuntimeError) as ei:
        U.gpu_update(sia, bad, True)
    assert ei.value.status == capi.ERR_NEGATIVE_THICKNESS
    # thickness above the top of the grid: IceGrid.cc:434-437
    bad = dict(inputs)
    bad["surface"] = inputs["surface"] + 5000.0
    bad["thickness"] = inputs["thickness"] + 5000.0
    bad["mask"] = np.full_like(inputs["mask"], 2.0)
    with pytest.raises(PISMRuntimeError) as ei:
        U.gpu_update(U.make_sia(grid, cfg, gb), bad, True)
    assert ei.value.status == capi.ERR_HEIGHT_ABOVE_TOP
    # D_max > D_limit without limiting: SIAFD.cc:752-760
    cfg.D_limit = 1e-3
    assert cases.oracle_run(grid, cfg, inputs, gb, full=False).status == 5
    with pytest.raises(PISMRuntimeError) as ei:
        U.gpu_update(U.make_sia(grid, cfg, gb), inputs, False)
    assert ei.value.status == capi.ERR_DIFFUSIVITY and "too high" in str(ei.value)


def test_bitwise_reproducible_and_tiling_independent():
    """Same inputs -> same bits, whatever the CTA row-segment length or the copy engine (mirrors the
    reference's decomposition-independence requirement, test/regression/test_02.sh)."""
    grid, cfg, inputs, gb = cases.case("C4s")
    outs = []
    for rows, bulk in ((64, 0), (64, 0), (7, 0), (16, 1), (64, 1)):
        sia = U.make_sia(grid, cfg, gb)
        sia.set_tuning(rows_per_cta=rows, use_bulk_copy=bulk)
        U.gpu_update(sia, inputs, True)
        outs.append({k: np.array(v, copy=True) for k, v in (("u", sia.velocity_u()), ("v", sia.velocity_v()),
                                                           ("D", sia.diffusivity()), ("Q", sia.diffusive_flux()))})
        outs[-1]["Dmax"] = sia.max_diffusivity()
    for o in outs[1:]:
        for k in ("u", "v", "D", "Q"):
            assert np.array_equal(o[k], outs[0][k]), k
        assert o["Dmax"] == outs[0]["Dmax"]


def test_device_resident_update_equals_host_path():
    """memory_space = 1 (torch CUDA tensors bound as the field storage) gives the same bits."""
    grid, cfg, inputs, gb = cases.case("dome_96_31")
    host = U.gpu_update(U.make_sia(grid, cfg, gb), inputs, True)
    dev_in = {k: torch.tensor(v, device="cuda") for k, v in inputs.items()}
    dev = U.gpu_update(U.make_sia(grid, cfg, gb), dev_in, True)
    torch.cuda.synchronize()
    for a, b in ((host.velocity_u(), dev.velocity_u()), (host.velocity_v(), dev.velocity_v()),
                 (host.diffusivity(), dev.diffusivity()), (host.diffusive_flux(), dev.diffusive_flux())):
        assert np.array_equal(a, b.cpu().numpy())
    assert host.max_diffusivity() == dev.max_diffusivity()
