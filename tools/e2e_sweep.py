"""Sweep of the host-side knobs of the pipelined host-buffer update (siafd_b200_update with pinned HOST arrays,
bench.py's `e2e`): host fill / replicate threads, row segments per band, level cut and its chunk width.  One process,
one set of pinned buffers, one fresh handle per setting (the knobs are read from the environment at
siafd_b200_create).  A setting is "KNOB=value;KNOB=value" with the SIAFD_B200_ prefix left off ("" = the defaults);
the old "fill_threads:band" form still works.  Before every setting the host's u, v are overwritten with NaN, and a
checksum of a sample of their rows (bit patterns) must agree between all settings.
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
    ap.add_argument("--settings", default="LEVEL_CUT=0,,CUT_COLS=64,CUT_COLS=256,REPL_THREADS=2,REPL_THREADS=8",
                    help="comma-separated settings, each KNOB=value;KNOB=value (or fill_threads:band)")
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
    knobs = ("FILL_THREADS", "BAND", "LEVEL_CUT", "CUT_COLS", "CUT_ROWS", "REPL_THREADS", "SPARSE", "ROWS", "PIPELINE", "TRACE", "ZERO_COPY")
    sums = []
    for s in args.settings.split(","):
        for k in knobs:
            os.environ.pop("SIAFD_B200_" + k, None)
        env = {}
        if ":" in s:
            ft, band = s.split(":")
            env = {"FILL_THREADS": ft, "BAND": band}
        elif s:
            env = dict(kv.split("=") for kv in s.split(";"))
        for k, v in env.items():
            assert k in knobs, k
            os.environ["SIAFD_B200_" + k] = v
        host["u"].fill_(float("nan"))
        host["v"].fill_(float("nan"))
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
        rows = slice(0, None, 5)
        chk = [int(host[n][rows].view(torch.int64).sum().item()) for n in ("u", "v")]
        sums.append(chk)
        results.append({"comm": bool(args.comm), "setting": s, "ms": ms, "best_ms": min(ms), "D_max": dmax,
                        "h2d_bytes_per_step": b[0].value // (args.steps + 1),
                        "d2h_bytes_per_step": b[1].value // (args.steps + 1),
                        "checksum_u_v_every_5th_row": chk, "same_bits_as_first_setting": chk == sums[0]})
        print(json.dumps(results[-1]), flush=True)
        del sia
    print(json.dumps({"size": M, "mz": Mz, "cpu_count": os.cpu_count(), "results": results}))


if __name__ == "__main__":
    main()
