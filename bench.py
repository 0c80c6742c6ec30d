#!/usr/bin/env python
"""bench.py -- SIAFD column-updates/s on B200s, with roofline, end-to-end and CPU-baseline figures.

    python bench.py --gpus N --steps K --warmup W                 # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Workload (BASELINE.json configs[4]): synthetic dome 4096 x 4096 x 101, gpbld flow law, haseloff
gradient, bed smoother off, full_update = true; strong scaling: the same grid split over N GPUs with
PISM's DMDA decomposition (IceGrid.cc:443-499), ghost exchange over NCCL.
One "step" = one SIAFD::update() of every column of the grid.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SIAFD column-updates/sec"
UNIT = "column-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="Mx = My of the dome")
    ap.add_argument("--mz", type=int, default=101)
    ap.add_argument("--flux-only", action="store_true", help="time full_update = false instead")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-domain", type=int, default=256, help="edge of each CPU-baseline sample domain")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--rows-per-cta", type=int, default=0)
    ap.add_argument("--bulk", type=int, default=-1)
    ap.add_argument("--with-w", action="store_true", help="(default at N = 1) kept for compatibility")
    ap.add_argument("--no-consumers", action="store_true",
                    help="skip timing the consumers of the update (SURVEY 8f: vertical velocity + CFL, strain heating, "
                         "mass-continuity step); they are reported under 'vertical_velocity' / 'consumers', after the "
                         "metric's timed region and not part of it")
    ap.add_argument("--uniform", action="store_true",
                    help="N > 1: PISM's default (equal) ownership ranges instead of the load-balanced -procs_x / -procs_y")
    ap.add_argument("--procs-x", default="", help="PISM's -procs_x: comma-separated ownership ranges in x")
    ap.add_argument("--procs-y", default="", help="PISM's -procs_y: comma-separated ownership ranges in y")
    ap.add_argument("--regime", default="dome", choices=["dome", "icefree", "allice"],
                    help="icefree: zero thickness everywhere (the write-only regime of the fused kernel; diagnostic)")
    ap.add_argument("--no-input-exchange", action="store_true",
                    help="skip the per-step width-2 exchange of the inputs' ghosts (N > 1)")
    return ap.parse_args()


def algorithmic_bytes_per_column(Mz, full):
    """read E, write u and v (3D) + 48 B of 2D reads + 64 B of 2D writes."""
    return (24 * Mz + 112) if full else (8 * Mz + 112)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (rank 0)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference; PETSc/MPI are not installable here)
# ------------------------------------------------------------------------------------------------
def cpu_arm(domain, Mz, nthreads, min_seconds, steps=None, warmup=1, full=True):
    """Time the CPU restatement on `nthreads` independent domain x domain x Mz dome domains, one per
    host thread (OpenMP), same flow law / gradient / smoother settings as the GPU workload.
    Returns (column-updates/s, seconds per step, steps, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import cases
    import oracle_lib as O
    grid, cfg, inputs, _ = cases.case("dome_%d_%d" % (domain, Mz))
    runs = []
    for _ in range(nthreads):
        inp = {k: np.array(v, copy=True) for k, v in inputs.items()}
        runs.append(O.Run(cfg.oracle_params(grid), inp))
    P = (O.Params * nthreads)(*[r.p for r in runs])
    Fa = (O.Fields * nthreads)(*[r.f for r in runs])
    L = O.lib()
    cols = grid.Mx * grid.My * nthreads

    def one():
        st = L.orc_siafd_update_many(nthreads, P, Fa, 1 if full else 0, nthreads)
        assert st == 0, st

    for _ in range(warmup):
        one()
    times = []
    t_begin = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_begin >= min_seconds and len(times) >= 2:
            break
    total = sum(times)
    sample = ("%d independent dome domains %dx%dx%d (one per host thread, OpenMP), %d updates each; "
              "restated reference (oracle port), PETSc/MPI unavailable" % (nthreads, domain, domain, Mz, len(times)))
    return cols * len(times) / total, total / len(times), len(times), sample


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = os.cpu_count() or 1
    full = not args.flux_only
    value, sec, steps, sample = cpu_arm(args.cpu_domain, args.mz, nthreads, 0.0, steps=args.steps,
                                        warmup=max(args.warmup, 1), full=full)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, "host cores only"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, decomposition):
    return {
        "workload": "synthetic dome %dx%dx%d SIAFD::update full_update=%s (BASELINE configs[4])" %
                    (args.size, args.size, args.mz, "false" if args.flux_only else "true"),
        "flow_law": "gpbld", "gradient": "haseloff", "bed_smoother": "off", "dx_m": 5000.0,
        "decomposition": decomposition,
        "l2": "inputs exceed L2 (%.1f GB enthalpy read + %.1f GB u,v written per step vs 126 MB L2)" %
              (args.size ** 2 * args.mz * 8 / 1e9, 2 * args.size ** 2 * args.mz * 8 / 1e9),
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return reference_main(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pism_b200 import capi, grid as G, synthetic as S
    from pism_b200.capi import F, lib
    from pism_b200.halo import PeerHalo, device_view, global_max
    from pism_b200.sia import SIAFD, PISMRuntimeError

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL announces its version on stdout at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "NONE"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = world
    full = not args.flux_only

    M, Mz = args.size, args.mz
    L = (M - 1) / 2.0 * 5000.0
    grid = G.Grid(M, M, Mz, L, L, 4000.0)
    # PISM's processor grid (IceGrid.cc:443-484).  Ownership ranges: what a PISM user passes as -procs_x / -procs_y
    # (IceGrid.cc:549-580) to balance the work -- here derived from the ice cover of the synthetic dome with the
    # measured cost ratio of an icy to an ice-free column -- or PISM's default equal ranges with --uniform.
    Nx_, Ny_ = G.compute_nprocs(M, M, N)
    procs_x = [int(v) for v in args.procs_x.split(",")] if args.procs_x else None
    procs_y = [int(v) for v in args.procs_y.split(",")] if args.procs_y else None
    ranges_note = "PISM DMDA rule"
    if N > 1 and not args.uniform and procs_x is None and procs_y is None:
        xs_ = torch.as_tensor(grid.x, dtype=torch.float64)
        icy = S.dome_2d(grid, capi.default_config(), xs_, torch.as_tensor(grid.y, dtype=torch.float64))["thickness"] > 0
        cost = np.where(icy.numpy(), 2.7, 1.0)  # 1.16 vs 3.2 G column-updates/s in the two regimes
        procs_x, procs_y = G.balanced_ownership_ranges(cost, Nx_, Ny_)
        del icy, cost
    if procs_x is not None or procs_y is not None:
        ranges_note = "PISM DMDA, -procs_x %s -procs_y %s" % (",".join(map(str, procs_x or G.ownership_ranges(M, Nx_))),
                                                             ",".join(map(str, procs_y or G.ownership_ranges(M, Ny_))))
        if not (args.procs_x or args.procs_y):
            ranges_note += " (balanced by ice cover)"
    patches = G.decompose(M, M, N, Nx_, Ny_, procs_x, procs_y)
    patch = patches[rank]
    cfg = capi.default_config()
    cfg.smoother_range = 0.0
    sia = SIAFD(grid, config=cfg, patch=patch, device=local_rank)
    if args.rows_per_cta or args.bulk >= 0:
        sia.set_tuning(args.rows_per_cta, args.bulk, -1)
    # One explicit stream carries everything: the library's kernels, torch's tensor ops and the NCCL transfers
    # (torch.distributed orders its communication stream against the CURRENT stream).  The legacy default
    # stream (handle 0) cannot be handed to the library: 0 means "use the handle's own stream" there.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    assert lib.siafd_b200_set_stream(sia.handle, stream.cuda_stream) == 0

    # ---- inputs resident in HBM, outputs too; bound as the handle's field storage ----
    t_gen = time.perf_counter()
    # the fields live in the handle's own device storage (exportable to the neighbours over CUDA IPC); torch
    # sees them as views
    inp = S.dome(grid, patch, sia.config, device=dev, Rfrac=1.5 if args.regime == "allice" else 0.75)
    fields = {}
    for name in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding", "h_x", "h_y", "D", "flux", "u", "v"):
        fields[name] = device_view(sia, name, sia.field_shape(name), dev)
        if name in inp:
            fields[name].copy_(inp[name])
    del inp
    if args.regime == "icefree":
        fields["thickness"].zero_()
        fields["mask"].zero_()
        fields["surface"].copy_(fields["bed"])
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    wg, we, ws = sia.config.w_geom, sia.config.w_3d_in, sia.config.w_sliding
    multi = N > 1
    halo = PeerHalo(patch, patches, sia, ["surface", "thickness", "mask", "bed", "enthalpy", "h_x", "h_y", "u", "v"]) \
        if multi else None

    def check(st):
        if st != 0:
            raise PISMRuntimeError(st, lib.siafd_b200_last_error(sia.handle).decode())

    def step_device():
        """SIAFD::update on device-resident fields, ghost exchanges where the reference has them."""
        if multi and not args.no_input_exchange:
            halo.exchange([("surface", wg), ("thickness", wg), ("mask", wg), ("bed", wg), ("enthalpy", we)], 0)
        check(lib.siafd_b200_compute_gradient(sia.handle))
        if multi:                            # SIAFD.cc:498-499
            halo.exchange([("h_x", 1), ("h_y", 1)], 1)
        else:
            check(lib.siafd_b200_wrap_ghosts_many(sia.handle, 2, (C.c_int * 2)(F["h_x"], F["h_y"])))
        check(lib.siafd_b200_compute_flux_velocity(sia.handle, 1 if full else 0, 0.0))
        if full:                             # SIAFD.cc:946-947
            if multi:
                halo.exchange([("u", 1), ("v", 1)], 2)
            else:
                check(lib.siafd_b200_wrap_ghosts_many(sia.handle, 2, (C.c_int * 2)(F["u"], F["v"])))
        check(lib.siafd_b200_finish(sia.handle))              # error flags + D_max (host sync, as in PISM)
        return global_max(lib.siafd_b200_max_diffusivity(sia.handle), dev)   # SIAFD.cc:748

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if multi:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # ---- timed region: K steps, device-resident ----
    W = max(args.warmup, 3)
    for _ in range(W):
        step_device()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    check(lib.siafd_b200_kernel_timing(sia.handle, 1))
    launches0 = sia.launch_count()
    if sampler:
        sampler.start()
    ms, dmax = timed(step_device, args.steps, 0)
    clocks = sampler.stop() if sampler else None
    launches = sia.launch_count() - launches0
    nk = C.c_int(0)
    kernel_ms = lib.siafd_b200_kernel_time_ms(sia.handle, C.byref(nk))
    check(lib.siafd_b200_kernel_timing(sia.handle, 0))
    cols_total = M * M
    value = cols_total * args.steps / (ms / 1e3)

    # decomposition check: sums over the OWNED points of |u|, |v| and |Q| (whatever the ranks and their ranges, these
    # agree to rounding of the summation order; D_max agrees exactly)
    checksum = None
    if full:
        wuv, wst = sia.config.w_uv, sia.config.w_stag
        parts = torch.stack([fields["u"][wuv:-wuv, wuv:-wuv].abs().sum(), fields["v"][wuv:-wuv, wuv:-wuv].abs().sum(),
                             fields["flux"][wst:-wst, wst:-wst].abs().sum()])
        if multi:
            dist.all_reduce(parts, op=dist.ReduceOp.SUM)
        checksum = {"sum_abs_u": float(parts[0]), "sum_abs_v": float(parts[1]), "sum_abs_flux": float(parts[2])}

    # roofline of the dominant (fused) kernel on this rank; report the slowest rank's
    B = algorithmic_bytes_per_column(Mz, full)
    k_avg_ms = kernel_ms / max(nk.value, 1)
    if multi:
        t = torch.tensor([k_avg_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_avg_ms = float(t.item())
    peak, peak_src = measured_peak()
    achieved = B * patch.xm * patch.ym / (k_avg_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "k_sia_slab", "kernel_ms": k_avg_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_column": B, "columns_per_launch": patch.xm * patch.ym,
                "whole_step_frac": B * cols_total / N / (ms / args.steps / 1e3) / 1e9 / peak}
    tr = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tr) and N == 1 and full:
        try:
            roofline["traffic"] = json.load(open(tr)).get("bytes_per_launch_%d" % M)
        except Exception:
            pass

    vertical = None
    consumers = None
    if (args.with_w or not args.no_consumers) and full and not multi:
        def w_step():
            check(lib.siafd_b200_compute_vertical_velocity(sia.handle, 0, 0))
        ms_w, _ = timed(w_step, args.steps, 3)
        bw = 24 * Mz  # read u, v once, write w
        vertical = {"kernel": "k_vvel_slab (w + fused 3D CFL maxima)", "ms": ms_w / args.steps,
                    "algorithmic_bytes_per_column": bw,
                    "achieved_GBps": bw * cols_total / (ms_w / args.steps / 1e3) / 1e9,
                    "frac": bw * cols_total / (ms_w / args.steps / 1e3) / 1e9 / peak}
        # the rest of a mass-continuity time step, device-resident
        out8 = (C.c_double * 8)()
        smb = torch.zeros((patch.ym, patch.xm), dtype=torch.float64, device=dev)
        check(lib.siafd_b200_bind(sia.handle, F["smb"], smb.data_ptr()))

        def cfl_fused():
            check(lib.siafd_b200_compute_vertical_velocity(sia.handle, 0, 0))
            check(lib.siafd_b200_cfl(sia.handle, 1.9e9, 1, out8))
        ms_cf, _ = timed(cfl_fused, args.steps, 3)

        def mass_step():  # dt = 0: the geometry stays what it is, the traffic is the same
            check(lib.siafd_b200_mass_flow_step(sia.handle, 0.0))
            check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
            check(lib.siafd_b200_mass_source_step(sia.handle, 0.0, 910.0, 0))
            check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
        ms_m, _ = timed(mass_step, args.steps, 3)
        def heat_step():
            check(lib.siafd_b200_compute_strain_heating(sia.handle, 2, 3.0, 1.0))  # gpbld, n = 3, e = 1 (ssa defaults)
        ms_h, _ = timed(heat_step, args.steps, 3)
        bh = 32 * Mz  # read E, u, v, write Sigma
        consumers = {"vertical_velocity_plus_cfl_ms": ms_cf / args.steps, "cfl3d_dt_s": out8[0],
                     "strain_heating": {"kernel": "k_strain_heating<gpbld>", "ms": ms_h / args.steps,
                                        "algorithmic_bytes_per_column": bh,
                                        "achieved_GBps": bh * cols_total / (ms_h / args.steps / 1e3) / 1e9,
                                        "frac": bh * cols_total / (ms_h / args.steps / 1e3) / 1e9 / peak,
                                        "note": "FP64-bound where there is ice (exp + 2 cbrt per level), "
                                                "write-bound elsewhere"},
                     "mass_continuity_step_ms": ms_m / args.steps,
                     "mass_continuity_launches": 8,
                     "note": "flow step + ensure_consistency + source step + ensure_consistency (2D fields only)"}

    # ---- end to end through the reference-facing call with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        host = {}
        h2d = d2h = 0
        for name in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding"):
            host[name] = torch.empty(fields[name].shape, dtype=torch.float64, pin_memory=True)
            host[name].copy_(fields[name])
            h2d += host[name].numel() * 8
        outs = ["h_x", "h_y", "D", "flux"] + (["u", "v"] if full else [])
        for name in outs:
            host[name] = torch.empty(fields[name].shape, dtype=torch.float64, pin_memory=True)
            d2h += host[name].numel() * 8
        torch.cuda.synchronize()

        def pd(t):
            return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_double))

        cin, cout = capi.Inputs(), capi.Outputs()
        cin.surface, cin.thickness, cin.mask, cin.bed = pd(host["surface"]), pd(host["thickness"]), pd(host["mask"]), pd(host["bed"])
        cin.enthalpy, cin.sliding = pd(host["enthalpy"]), pd(host["sliding"])
        cin.current_time, cin.memory_space, cin.ghosts_valid = 0.0, 0, 1
        cout.h_x, cout.h_y, cout.D, cout.flux = pd(host["h_x"]), pd(host["h_y"]), pd(host["D"]), pd(host["flux"])
        if full:
            cout.u, cout.v = pd(host["u"]), pd(host["v"])
        cout.memory_space = 0

        def step_e2e():
            if not multi:
                check(lib.siafd_b200_update(sia.handle, C.byref(cin), C.byref(cout), 1 if full else 0))
                return lib.siafd_b200_max_diffusivity(sia.handle)
            for name in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding"):
                check(lib.siafd_b200_upload(sia.handle, F[name], host[name].data_ptr()))
            d = step_device()
            for name in outs:
                check(lib.siafd_b200_download(sia.handle, F[name], host[name].data_ptr()))
            return d

        b0, b1 = (C.c_int64(), C.c_int64()), (C.c_int64(), C.c_int64())
        step_e2e()  # (warm-up, before the byte counters are read)
        lib.siafd_b200_transfer_bytes(sia.handle, C.byref(b0[0]), C.byref(b0[1]))
        ms_e, dmax_e = timed(step_e2e, args.e2e_steps, 0)
        lib.siafd_b200_transfer_bytes(sia.handle, C.byref(b1[0]), C.byref(b1[1]))
        # bytes the library actually moved per step (it skips the parts of the 3D arrays that hold no ice: no enthalpy
        # is read there and u = v = sliding velocity, which it fills in on the host); "dense" = the arrays' full sizes
        moved = [(b1[q].value - b0[q].value) // args.e2e_steps for q in (0, 1)]
        if multi:
            t = torch.tensor(moved, dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            moved = [int(t[0].item()), int(t[1].item())]
        e2e = {"value": cols_total * args.e2e_steps / (ms_e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(moved[0]), "d2h_bytes_per_step": int(moved[1]),
               "dense_h2d_bytes_per_step": int(h2d * N), "dense_d2h_bytes_per_step": int(d2h * N),
               "ms_per_step": ms_e / args.e2e_steps, "steps": args.e2e_steps,
               "api": "siafd_b200_update(host pointers)" if not multi else "upload + split update + download",
               "host_memory": "pinned"}
        assert dmax_e == dmax
        if full:  # the host arrays are the device-resident result, bit for bit (ghosts included)
            for name in ("u", "v", "flux"):
                chunk = 256
                for j0 in range(0, host[name].shape[0], chunk):
                    assert torch.equal(host[name][j0:j0 + chunk].to(dev), fields[name][j0:j0 + chunk]), (name, j0)
            e2e["verified"] = "host u, v, flux == device-resident u, v, flux (bitwise)"
        del host

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        nthreads = os.cpu_count() or 1
        v, sec, steps_c, sample = cpu_arm(args.cpu_domain, Mz, nthreads, args.cpu_seconds, full=full)
        cpu = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, "%dx%d (%s)" % (patch.Nx, patch.Ny, ranges_note)),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "D_max": dmax, "checksum": checksum, "vertical_velocity": vertical, "consumers": consumers, "halo_bytes_per_step_per_rank": (halo.bytes_sent // max(args.steps + W + 1, 1)) if halo else 0,
            "halo_transport": "direct stores into CUDA-IPC-mapped neighbour arrays (NVLink), 3 phases/step" if halo else None,
            "input_generation_s": t_gen,
        }
        print(json.dumps(line))
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
