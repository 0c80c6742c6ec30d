"""N > 1 host logic on CPU: the two-stage halo exchange (pism_b200/halo.py) over the gloo backend must
reproduce the periodic BOX-stencil ghost update of PISM's DMDA (IceGrid.cc:863-885, iceModelVec.cc:630-643)
for PISM's own decomposition (IceGrid.cc:443-499), and the oracle run on the exchanged patches must be
bitwise identical to the single-patch run (test/regression/test_02.sh)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker_exchange(rank, world, port, Mx, My, q, procs=None):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pism_b200 import grid as G
        from pism_b200.halo import HaloExchanger, global_max
        patches = G.decompose(Mx, My, world, **(procs or {}))
        pt = patches[rank]
        ex = HaloExchanger(pt)
        rng = np.random.RandomState(7)
        ok = True
        for w, dof in ((2, 1), (1, 2), (2, 5), (1, 7)):
            shape = (My, Mx) if dof == 1 else (My, Mx, dof)
            g = rng.rand(*shape)
            want = G.global_to_local(g, pt, w)
            a = np.full_like(want, np.nan)
            a[w:-w, w:-w] = want[w:-w, w:-w]
            t = torch.from_numpy(a)
            ex.exchange("enthalpy", t if dof > 1 else t, w)
            ok = ok and np.array_equal(a, want)
        m = global_max(float(rank + 1), "cpu")
        ok = ok and m == float(world)
        q.put((rank, ok, None))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,Mx,My,ranges", [(2, 12, 16, None), (2, 16, 9, None), (4, 13, 11, None), (3, 30, 7, None),
                                                 # PISM's -Nx / -Ny / -procs_x / -procs_y: unequal ownership ranges
                                                 (4, 13, 11, dict(Nx=2, Ny=2, procs_x=[9, 4], procs_y=[3, 8])),
                                                 (3, 9, 20, dict(Nx=1, Ny=3, procs_y=[11, 2, 7]))])
def test_halo_exchange_matches_periodic_box_ghosts(world, Mx, My, ranges):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, world, port, Mx, My, q, ranges)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in res:
        assert ok, (rank, err)


def _worker_oracle(rank, world, port, q):
    """Each rank runs the ORACLE on its patch with ghost exchanges through HaloExchanger: the multi-rank
    driver logic of bench.py (gradient -> exchange h_x,h_y -> flux/velocity -> exchange u,v -> global max)."""
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import cases
        import oracle_lib as O
        from pism_b200 import grid as G
        from pism_b200.halo import HaloExchanger, global_max
        grid, cfg, inputs, gb = cases.case("C4s")
        whole = cases.oracle_run(grid, cfg, inputs, gb, full=True)
        pt = G.decompose(grid.Mx, grid.My, world)[rank]
        _, _, loc, _ = cases.case("C4s", patch=pt)     # closed-form inputs of this patch, ghosts included
        sm = O.preprocess_bed(cfg.oracle_params(grid), gb)
        smoothed = {k: G.global_to_local(sm[k], pt, cfg.w_geom) for k in ("topgsmooth", "maxtl", "C2", "C3", "C4")}
        smoothed["active"] = sm["active"]
        run = O.Run(cfg.oracle_params(grid, pt), loc, smoothed)
        ex = HaloExchanger(pt)
        assert run.gradient() == 0
        for f in ("h_x", "h_y"):
            ex.exchange(f, torch.from_numpy(run.a[f]), 1)
        assert run.flux_velocity(True) == 0
        for f in ("u", "v"):
            ex.exchange(f, torch.from_numpy(run.a[f]), 1)
        dmax = global_max(run.D_max, "cpu")
        ok = dmax == whole.D_max
        for f, w in (("u", 1), ("v", 1), ("h_x", 1), ("h_y", 1)):   # ghosts valid after the exchange
            ok = ok and np.array_equal(run.a[f], G.global_to_local(cases.interior(whole.a[f], w), pt, w))
        for f in ("D", "Q"):                                        # owned + locally computed ghost ring
            ok = ok and np.array_equal(run.a[f], G.global_to_local(cases.interior(whole.a[f], 1), pt, 1))
        q.put((rank, bool(ok), None))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_two_rank_update_is_bitwise_identical_to_one_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_oracle, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in res:
        assert ok, (rank, err)
