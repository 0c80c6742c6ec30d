"""Sweep of the host-side knobs of the pipelined host-buffer update (siafd_b200_update with pinned HOST arrays,
bench.py's `e2e`): number of host fill threads and row segments per band.  One process, one set of pinned
buffers, one fresh handle per setting (the knobs are read from the environment at siafd_b200_create).
Usage (GPU box): python tools/e2e_sweep.py [--size 4096] [--steps 2] > gpurun_out/e2e_sweep.json"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--mz", type=int, default=101)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--settings", default="8:4,16:4,32:4,16:2,16:8", help="fill_threads:band, comma-separated")
    ap.add_argument("--comm", action="store_true", help="with a one-rank communicator (what bench.py's e2e does at N = 1)")
    args = ap.parse_args()

    import torch
    from pism_b200 import capi, grid as G, synthetic as S
    from pism_b200.capi import lib
    from pism_b200.sia import SIAFD

    dev = torch.device("cuda", 0)
    M, Mz = args.size, args.mz
    L = (M - 1) / 2.0 * 5000.0
    grid = G.Grid(M, M, Mz, L, L, 4000.0)
    cfg = capi.default_config()
    cfg.smoother_range = 0.0
    probe = SIAFD(grid, config=cfg, device=0)
    shapes = {n: probe.field_shape(n) for n in ("h_x", "h_y", "D", "flux", "u", "v")}
    in_shapes = {n: probe.field_shape(n) for n in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding")}
    inp = S.dome(grid, grid.whole(), probe.config, device=dev)
    del probe
    host = {}
    for name, shp in in_shapes.items():
        host[name] = torch.zeros(shp, dtype=torch.float64).pin_memory() if name not in inp else \
            torch.empty(shp, dtype=torch.float64, pin_memory=True)
        if name in inp:
            host[name].copy_(inp[name])
    del inp
    for name, shp in shapes.items():
        host[name] = torch.empty(shp, dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    def pd(t):
        return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_double))

    cin, cout = capi.Inputs(), capi.Outputs()
    cin.surface, cin.thickness, cin.mask, cin.bed = pd(host["surface"]), pd(host["thickness"]), pd(host["mask"]), pd(host["bed"])
    cin.enthalpy, cin.sliding = pd(host["enthalpy"]), pd(host["sliding"])
    cin.current_time, cin.memory_space, cin.ghosts_valid = 0.0, 0, 1
    cout.h_x, cout.h_y, cout.D, cout.flux = pd(host["h_x"]), pd(host["h_y"]), pd(host["D"]), pd(host["flux"])
    cout.u, cout.v = pd(host["u"]), pd(host["v"])
    cout.memory_space = 0

    results = []
    for s in args.settings.split(","):
        ft, band = s.split(":")
        os.environ["SIAFD_B200_FILL_THREADS"] = ft
        os.environ["SIAFD_B200_BAND"] = band
        sia = SIAFD(grid, config=cfg, device=0)
        if args.comm:
            hs = (C.c_void_p * 1)(sia.handle)
            assert lib.siafd_b200_comm_init_local(hs, 1) == 0
        ms = []
        for it in range(args.steps + 1):
            t0 = time.perf_counter()
            st = lib.siafd_b200_update(sia.handle, C.byref(cin), C.byref(cout), 1)
            dmax = lib.siafd_b200_max_diffusivity(sia.handle)
            t1 = time.perf_counter()
            assert st == 0, lib.siafd_b200_last_error(sia.handle).decode()
            if it > 0:
                ms.append((t1 - t0) * 1e3)
        b = (C.c_int64(), C.c_int64())
        lib.siafd_b200_transfer_bytes(sia.handle, C.byref(b[0]), C.byref(b[1]))
        results.append({"comm": bool(args.comm), "fill_threads": int(ft), "band": int(band), "ms": ms, "best_ms": min(ms), "D_max": dmax,
                        "h2d_bytes_per_step": b[0].value // (args.steps + 1),
                        "d2h_bytes_per_step": b[1].value // (args.steps + 1),
                        "sum_abs_u_row_2048": float(host["u"][M // 2].abs().sum())})
        print(json.dumps(results[-1]), flush=True)
        del sia
    print(json.dumps({"size": M, "mz": Mz, "cpu_count": os.cpu_count(), "results": results}))


if __name__ == "__main__":
    main()
