"""N > 1 host logic of the time step on CPU (gloo): `pismv -test C` with the domain cut into patches, one process per
patch, ghosts of H / h_x, h_y / u, v exchanged where the reference exchanges them and D_max / CFL reduced over
ranks (pism_b200.icemodel.Ranks + pism_b200.halo.HaloExchanger) -- every rank must print the reference's golden
rows of test/regression/test_15.sh, i.e. the run does not depend on the decomposition (test_02.sh).  The field
arithmetic is the oracle's here; on GPUs the same driver runs DeviceBackend + PeerHalo (tools/pismv_multi_gpu.py)."""
import os
import socket
import sys

import pytest
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


def _worker(rank, world, port, M, decomp, q, testname="C", years=5000.0):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import pismv_oracle as PO
        from pism_b200 import grid as G
        from pism_b200 import icemodel
        import cases
        grid, _, _, _ = cases.case("C1_%d" % M)  # (only Mx, My matter to Ranks; pismv_model builds the test's own grid)
        ranks = icemodel.Ranks(grid, G.decompose(grid.Mx, grid.My, world, **decomp), rank)
        m = PO.pismv_model(testname, M, run_length_years=years, ranks=ranks)
        m.run()
        q.put((rank, m.steps, m.report(), None))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, -1, "", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,M,decomp", [(2, 31, {}), (4, 41, dict(Nx=2, Ny=2, procs_x=[25, 16], procs_y=[12, 29]))])
def test_pismv_test_C_on_patches_prints_the_golden_rows(world, M, decomp):
    import pismv_oracle as PO
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, decomp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, steps, report, err in res:
        assert err is None, (rank, err)
        assert steps == 84
        assert report == PO.TEST_15_GOLDEN[M], (rank, report)


def test_pismv_test_L_on_two_ranks_like_test_16():
    """test/regression/test_16.sh runs `mpiexec -n 2 pismv -test L`: two patches (PISM's 1 x 2 split), non-flat bed,
    diffusivity-limited steps from the GLOBAL D_max; both ranks print the golden rows."""
    import pismv_oracle as PO
    ctx = mp.get_context("spawn")
    for M in (21, 31):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, M, {}, q, "L", 1000.0)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=600) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        for rank, steps, report, err in res:
            assert err is None, (rank, err)
            assert report == PO.TEST_16_GOLDEN[M], (M, rank, report)

