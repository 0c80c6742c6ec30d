#!/usr/bin/env python3
"""`pismv -test C` (test/regression/test_15.sh) time-stepped on SEVERAL GPUs: one process per GPU, PISM's decomposition,
every field resident in its rank's HBM, ghost updates by direct stores into the neighbours' arrays (PeerHalo),
D_max / CFL reduced with NCCL.  Prints the reference's report; exit code 1 if it is not the golden row.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pismv_multi_gpu.py [M ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import cases
import gpu_util as U
import pismv_oracle as PO
from pism_b200 import grid as G
from pism_b200 import icemodel
from pism_b200.halo import PeerHalo


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for M in [int(a) for a in sys.argv[1:]] or [31, 41]:
        grid, cfg, _, _ = cases.case("C1_%d" % M)
        patches = G.decompose(grid.Mx, grid.My, world)
        ranks = icemodel.Ranks(grid, patches, rank)

        def factory(grid_, cfg_, inputs, max_dt, ranks=None):
            sia = U.make_sia(grid_, cfg_, None, patch=ranks.patch)
            for name in ("thickness", "h_x", "h_y", "u", "v"):   # PeerHalo maps handle-owned storage
                sia.upload(name, torch.zeros(sia.field_shape(name), dtype=torch.float64).numpy())
            halo = PeerHalo(ranks.patch, patches, sia, ["thickness", "h_x", "h_y", "u", "v"]) if world > 1 else None
            return icemodel.DeviceBackend(sia, inputs, max_dt, ranks=ranks, halo=halo)

        t0 = time.perf_counter()
        m = PO.pismv_model("C", M, backend_factory=factory, ranks=ranks)
        m.run()
        rep = m.report()
        if rank == 0:
            good = rep == PO.TEST_15_GOLDEN[M]
            ok = ok and good
            print("M = %d, %d ranks (%dx%d), %d steps, %.1f s:%s  %s" % (
                M, world, patches[0].Nx, patches[0].Ny, m.steps, time.perf_counter() - t0, rep,
                "== golden row of test_15.sh" if good else "!= " + PO.TEST_15_GOLDEN[M]), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
