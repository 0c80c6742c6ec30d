#!/usr/bin/env python
"""bench.py -- SIAFD column-updates/s on B200s, with roofline, end-to-end and CPU-baseline figures.

    python bench.py --gpus N --steps K --warmup W                 # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Workload (BASELINE.json configs[4]): synthetic dome 4096 x 4096 x 101, gpbld flow law, haseloff
gradient, bed smoother off, full_update = true; strong scaling: the same grid split over N GPUs with
PISM's DMDA decomposition (IceGrid.cc:443-499), ghost updates by direct stores into the neighbours'
arrays over NVLink (the library's own communicator, siafd_b200_comm_*: no NCCL on the data path;
torch.distributed is only the launcher's barrier / max-over-ranks plumbing).
One "step" = one SIAFD::update() of every column of the grid.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import hashlib
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
    ap.add_argument("--cpu-domain", type=int, default=0,
                    help="edge of the CPU arm's domain (0: 2048 if the host has >= 40 GB of RAM, else 1024)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--rows-per-cta", type=int, default=0)
    ap.add_argument("--bulk", type=int, default=-1)
    ap.add_argument("--no-consumers", action="store_true",
                    help="skip timing the consumers of the update (SURVEY 8f: vertical velocity + CFL, strain heating, "
                         "mass-continuity step); they are reported under 'vertical_velocity' / 'consumers', after the "
                         "metric's timed region and not part of it")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra records: full_update = false, the all-ice regime (2048^2), and at N = 8 PISM's "
                         "default (equal) ownership ranges beside the balanced ones")
    ap.add_argument("--uniform", action="store_true",
                    help="N > 1: PISM's default (equal) ownership ranges as the headline instead of the balanced ones")
    ap.add_argument("--procs-x", default="", help="PISM's -procs_x: comma-separated ownership ranges in x")
    ap.add_argument("--procs-y", default="", help="PISM's -procs_y: comma-separated ownership ranges in y")
    ap.add_argument("--regime", default="dome", choices=["dome", "icefree", "allice"],
                    help="icefree: zero thickness everywhere (the write-only regime of the fused kernel); allice: margin "
                         "radius 1.5 Lx, every column carries ice (diagnostics)")
    ap.add_argument("--no-input-exchange", action="store_true",
                    help="skip the per-step width-2 exchange of the inputs' ghosts (N > 1)")
    ap.add_argument("--cost-ratio", type=float, default=0.0,
                    help="constant cost of an icy column relative to an ice-free one for the balanced ownership ranges "
                         "(0: the thickness-proportional model)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one instead of a CUDA graph")
    return ap.parse_args()


def algorithmic_bytes_per_column(Mz, full):
    """read E, write u and v (3D) + 48 B of 2D reads + 64 B of 2D writes."""
    return (24 * Mz + 112) if full else (8 * Mz + 112)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """Identifies the fused kernel's source: a DRAM-traffic figure taken under ncu is only quoted for the same source."""
    h = hashlib.sha256()
    for f in ("siafd_slab.cu", "siafd_math.cuh", "siafd_device.cuh"):
        h.update(open(os.path.join(ROOT, "pism_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


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
def native_oracle():
    """The oracle built for timing: -O3 -march=native on THIS machine (oracle/Makefile `native`); the parity tests keep
    the -O2 -ffp-contract=off build.  Returns (ctypes library with oracle_lib's signatures, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    path = os.path.join(ROOT, "oracle", "_native", "liboracle_siafd_native.so")
    try:
        subprocess.run(["make", "-B", "-C", os.path.join(ROOT, "oracle"), "native"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(path)
        PP, FP = C.POINTER(O.Params), C.POINTER(O.Fields)
        L.orc_siafd_update_many.argtypes = [C.c_int, PP, FP, C.c_int, C.c_int]
        L.orc_siafd_update_decomposed.argtypes = [C.c_int, PP, FP, C.c_int, C.c_int]
        return L, "g++ -O3 -march=native -fopenmp (built on this host)"
    except Exception as e:  # no compiler on the box: the parity build, and say so
        return O.lib(), "g++ -O2 -ffp-contract=off -fopenmp (parity build; native build failed: %s)" % type(e).__name__


def host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().total / 2 ** 30
    except Exception:
        return 0.0


def cpu_arm(domain, Mz, nthreads, min_seconds, steps=None, warmup=1, full=True, tiles=False):
    """Time the CPU restatement of SIAFD::update on all host cores, same flow law / gradient / smoother settings as the
    GPU workload.  Default: ONE domain x domain x Mz dome split into `nthreads` patches by PISM's own rule
    (compute_nprocs, IceGrid.cc:443-499), one OpenMP thread per patch standing for one MPI rank, ghost updates of
    h_x, h_y and u, v as copies between the patches (orc_siafd_update_decomposed).  tiles=True: `nthreads` independent
    domain^2 domes instead (round 1's sample).  Returns (column-updates/s, s per step, steps, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import cases
    import oracle_lib as O
    from pism_b200 import grid as G, synthetic as S
    L, how = native_oracle()
    runs = []
    if tiles:
        grid, cfg, inputs, _ = cases.case("dome_%d_%d" % (domain, Mz))
        for _ in range(nthreads):
            inp = {k: np.array(v, copy=True) for k, v in inputs.items()}
            runs.append(O.Run(cfg.oracle_params(grid), inp))
        cols = grid.Mx * grid.My * nthreads
        what = "%d independent dome domains %dx%dx%d (one per host thread)" % (nthreads, domain, domain, Mz)
    else:
        Lxy = (domain - 1) / 2.0 * 5000.0
        grid = G.Grid(domain, domain, Mz, Lxy, Lxy, 4000.0)
        cfg = cases.Cfg(flow_law="gpbld", smoother_range=0.0)
        Nx, Ny = G.compute_nprocs(domain, domain, nthreads)
        patches = G.decompose(domain, domain, nthreads, Nx, Ny)
        for pt in patches:
            inp = cases.to_numpy(S.dome(grid, pt, cfg, device="cpu"))
            runs.append(O.Run(cfg.oracle_params(grid, pt), inp))
        cols = domain * domain
        what = ("ONE dome %dx%dx%d split into %d x %d patches by PISM's compute_nprocs, one OpenMP thread per patch "
                "(= MPI rank), ghost copies of h_x, h_y and u, v between patches" % (domain, domain, Mz, Nx, Ny))
    n = len(runs)
    P = (O.Params * n)(*[r.p for r in runs])
    Fa = (O.Fields * n)(*[r.f for r in runs])

    def one():
        if tiles:
            st = L.orc_siafd_update_many(n, P, Fa, 1 if full else 0, nthreads)
        else:
            st = L.orc_siafd_update_decomposed(n, P, Fa, 1 if full else 0, nthreads)
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
    sample = "%s, %d updates; restated reference (oracle port; PETSc/MPI unavailable), %s" % (what, len(times), how)
    return cols * len(times) / total, total / len(times), len(times), sample


def cpu_domain(args):
    if args.cpu_domain:
        return args.cpu_domain
    return 2048 if host_ram_gb() >= 40.0 else 1024  # 7 live 3D fields: 56 Mz B / column = 24 GB at 2048^2


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = os.cpu_count() or 1
    full = not args.flux_only
    dom = cpu_domain(args)
    value, sec, steps, sample = cpu_arm(dom, args.mz, nthreads, 0.0, steps=args.steps, warmup=max(args.warmup, 1), full=full)
    cfgw = workload_config(args, "host cores only: %d OpenMP threads, one patch each" % nthreads)
    cfgw["workload"] = ("synthetic dome %dx%dx%d SIAFD::update full_update=%s on host cores (bounded sample of BASELINE "
                        "configs[4], which is 4096x4096x%d: per-column rates compare)" %
                        (dom, dom, args.mz, "false" if args.flux_only else "true", args.mz))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfgw,
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
class Rank:
    """One rank's handle, device-resident fields and communicator for a given decomposition of a given dome."""

    def __init__(self, M, Mz, regime, patches, rank, local_rank, stream, args):
        import torch
        from pism_b200 import capi, grid as G, synthetic as S
        from pism_b200.capi import F, lib
        from pism_b200.halo import PeerHalo, device_view
        from pism_b200.sia import SIAFD
        self.lib, self.F = lib, F
        self.M, self.Mz, self.patches, self.patch = M, Mz, patches, patches[rank]
        self.N = len(patches)
        dev = torch.device("cuda", local_rank)
        L = (M - 1) / 2.0 * 5000.0
        self.grid = G.Grid(M, M, Mz, L, L, 4000.0)
        cfg = capi.default_config()
        cfg.smoother_range = 0.0
        self.sia = SIAFD(self.grid, config=cfg, patch=self.patch, device=local_rank)
        if args.rows_per_cta or args.bulk >= 0:
            self.sia.set_tuning(args.rows_per_cta, args.bulk, -1)
        assert lib.siafd_b200_set_stream(self.sia.handle, stream.cuda_stream) == 0
        t0 = time.perf_counter()
        inp = S.dome(self.grid, self.patch, self.sia.config, device=dev, Rfrac=1.5 if regime == "allice" else 0.75)
        self.fields = {}
        for name in ("surface", "thickness", "mask", "bed", "enthalpy", "sliding", "h_x", "h_y", "D", "flux", "u", "v"):
            self.fields[name] = device_view(self.sia, name, self.sia.field_shape(name), dev)
            if name in inp:
                self.fields[name].copy_(inp[name])
        del inp
        if regime == "icefree":
            self.fields["thickness"].zero_()
            self.fields["mask"].zero_()
            self.fields["surface"].copy_(self.fields["bed"])
        torch.cuda.synchronize()
        self.t_gen = time.perf_counter() - t0
        if self.N > 1:
            self.halo = PeerHalo(self.patch, patches, self.sia)  # the library's communicator (CUDA IPC, files)
        else:
            hs = (C.c_void_p * 1)(self.sia.handle)
            assert lib.siafd_b200_comm_init_local(hs, 1) == 0, lib.siafd_b200_last_error(self.sia.handle)
            self.halo = None
        self.exchange_inputs = 1 if (self.N > 1 and not args.no_input_exchange) else 0

    def check(self, st):
        from pism_b200.sia import PISMRuntimeError
        if st != 0:
            raise PISMRuntimeError(st, self.lib.siafd_b200_last_error(self.sia.handle).decode())

    def step(self, full=True):
        """SIAFD::update on device-resident fields, every ghost update and the D_max / status reduction inside
        (siafd_b200_update_decomposed: one CUDA-graph launch), then the one host synchronisation PISM has too."""
        h = self.sia.handle
        self.check(self.lib.siafd_b200_update_decomposed(h, 1 if full else 0, 0.0, self.exchange_inputs))
        self.check(self.lib.siafd_b200_finish(h))
        return self.lib.siafd_b200_max_diffusivity(h)   # SIAFD.cc:748: already the maximum over all ranks


def main():
    args = parse()
    if args.impl == "reference":
        return reference_main(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pism_b200 import capi, grid as G, synthetic as S
    from pism_b200.capi import F, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if args.no_graph:
        os.environ["SIAFD_B200_GRAPH"] = "0"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = world
    multi = N > 1
    full = not args.flux_only
    M, Mz = args.size, args.mz

    # PISM's processor grid (IceGrid.cc:443-484).  Ownership ranges: what a PISM user passes as -procs_x / -procs_y
    # (IceGrid.cc:549-580) to balance the work -- here derived from the ice cover of the synthetic dome with the
    # measured cost ratio of an icy to an ice-free column -- or PISM's default equal ranges (--uniform; also timed
    # beside the balanced ones and reported under "uniform_ranges" whenever the two differ).
    Nx_, Ny_ = G.compute_nprocs(M, M, N)
    procs_x = [int(v) for v in args.procs_x.split(",")] if args.procs_x else None
    procs_y = [int(v) for v in args.procs_y.split(",")] if args.procs_y else None
    ranges_note = "PISM DMDA rule"
    if multi and not args.uniform and procs_x is None and procs_y is None and args.regime == "dome":
        Lh = (M - 1) / 2.0 * 5000.0
        g0 = G.Grid(M, M, Mz, Lh, Lh, 4000.0)
        H0 = S.dome_2d(g0, capi.default_config(), torch.as_tensor(g0.x, dtype=torch.float64),
                       torch.as_tensor(g0.y, dtype=torch.float64))["thickness"].numpy()
        # cost of a column relative to an ice-free one: fitted to the per-rank times of the fused kernel
        # measured at 8 ranks (profiles/scale8_variants_r02.json: 1.045 ms on the outer, 1.333 ms on the central patches
        # of -procs_y 1238,810,810,1238): an ice-free column 0.274 ns, an icy one 0.299 ns + 0.639 ns x H / mean(H), in
        # proportion to the levels below the surface.  A constant ratio (--cost-ratio 2.7: round 1) leaves the central
        # ranks 28 % slower than the outer ones.
        if args.cost_ratio > 0:
            cost = np.where(H0 > 0, args.cost_ratio, 1.0)
        else:
            cost = np.where(H0 > 0, 1.09 + 2.34 * H0 / H0[H0 > 0].mean(), 1.0)
        procs_x, procs_y = G.balanced_ownership_ranges(cost, Nx_, Ny_)
        del H0, cost
    uniform_x, uniform_y = G.ownership_ranges(M, Nx_), G.ownership_ranges(M, Ny_)
    if procs_x is not None or procs_y is not None:
        ranges_note = "PISM DMDA, -procs_x %s -procs_y %s" % (",".join(map(str, procs_x or uniform_x)),
                                                             ",".join(map(str, procs_y or uniform_y)))
        if not (args.procs_x or args.procs_y):
            ranges_note += " (balanced by ice cover)"
    patches = G.decompose(M, M, N, Nx_, Ny_, procs_x, procs_y)
    differs_from_uniform = multi and ((procs_x or uniform_x) != uniform_x or (procs_y or uniform_y) != uniform_y)

    # One explicit stream carries everything: the library's kernels and torch's tensor ops.  The legacy default stream
    # (handle 0) cannot be handed to the library: 0 means "use the handle's own stream" there.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    R = Rank(M, Mz, args.regime, patches, rank, local_rank, stream, args)
    sia, fields, patch = R.sia, R.fields, R.patch

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), out

    per_rank_kernel_ms = []
    step_breakdown = {}

    def kernel_only_ms(Rk, steps, fullk=True):
        """CUDA events around the fused kernel alone, on its own stream (ungraphed launches of the same step)."""
        Rk.check(lib.siafd_b200_kernel_timing(Rk.sia.handle, 1))
        barrier()
        for _ in range(steps):
            Rk.step(fullk)
        barrier()
        nk = C.c_int(0)
        kms = lib.siafd_b200_kernel_time_ms(Rk.sia.handle, C.byref(nk))
        sec, ns = (C.c_double * 5)(), C.c_int(0)
        Rk.check(lib.siafd_b200_step_breakdown_ms(Rk.sia.handle, sec, C.byref(ns)))
        Rk.check(lib.siafd_b200_kernel_timing(Rk.sia.handle, 0))
        mine = kms / max(nk.value, 1)
        # the sections of an (ungraphed) step, per rank; reported as the maximum over the ranks of each section
        names = ("input_ghosts_2d", "gradient_pass", "wait_before_fused_kernel", "fused_kernel", "uv_arrival_and_reduction")
        vals = torch.tensor(list(sec), dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        step_breakdown.clear()
        step_breakdown.update({n: float(v) for n, v in zip(names, vals.tolist())})
        if multi:
            t = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(N)]
            dist.all_gather(t, torch.tensor([mine], dtype=torch.float64, device=dev))
            per_rank_kernel_ms[:] = [float(x.item()) for x in t]
        return max_over_ranks(mine)

    peak, peak_src = measured_peak()

    def measure(Rk, steps, fullk=True):
        """Whole-step time (max over ranks) and the fused kernel's own time for one Rank."""
        barrier()
        for _ in range(3):
            Rk.step(fullk)
        k_ms = kernel_only_ms(Rk, min(steps, 10), fullk)
        ms_, dmax_ = timed(lambda: Rk.step(fullk), steps, 1)
        Bc = algorithmic_bytes_per_column(Rk.Mz, fullk)
        cols = Rk.M * Rk.M
        ach = Bc * Rk.patch.xm * Rk.patch.ym / (k_ms / 1e3) / 1e9
        return {"ms_per_step": ms_ / steps, "value": cols * steps / (ms_ / 1e3), "kernel_ms": k_ms,
                "kernel_frac": ach / peak, "whole_step_frac": Bc * cols / Rk.N / (ms_ / steps / 1e3) / 1e9 / peak,
                "D_max": dmax_}

    # ---- timed region: K steps, device-resident ----
    W = max(args.warmup, 3)
    for _ in range(W):
        R.step(full)
    k_avg_ms = kernel_only_ms(R, min(args.steps, 10), full)
    kernel_ms_by_rank = list(per_rank_kernel_ms)
    breakdown = dict(step_breakdown)
    R.step(full)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = sia.launch_count()
    if sampler:
        sampler.start()
    ms, dmax = timed(lambda: R.step(full), args.steps, 0)
    clocks = sampler.stop() if sampler else None
    launches = sia.launch_count() - launches0
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

    # roofline of the dominant (fused) kernel on the slowest rank
    B = algorithmic_bytes_per_column(Mz, full)
    achieved = B * patch.xm * patch.ym / (k_avg_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "dram_frac": None, "kernel": "k_sia_slab", "kernel_ms": k_avg_ms,
                "peak_source": peak_src, "algorithmic_bytes_per_column": B, "columns_per_launch": patch.xm * patch.ym,
                "whole_step_frac": B * cols_total / N / (ms / args.steps / 1e3) / 1e9 / peak,
                "step_minus_kernel_ms": ms / args.steps - k_avg_ms, "kernel_ms_by_rank": kernel_ms_by_rank or None,
                "step_breakdown_ms": breakdown or None,
                "step_breakdown_note": "CUDA events between the launches of an ungraphed step (max over ranks per "
                                       "section); the timed region replays the same launches as one CUDA graph"}
    tr = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tr) and N == 1 and full and args.regime == "dome":
        try:  # quoted only if it was captured (ncu --set full) on this very kernel source
            t = json.load(open(tr))
            if t.get("kernel_source_sha16") == kernel_source_hash() and t.get("bytes_per_launch_%d" % M):
                roofline["traffic"] = t["bytes_per_launch_%d" % M]
                roofline["dram_frac"] = roofline["traffic"] / (k_avg_ms / 1e3) / 1e9 / peak
        except Exception:
            pass

    # ---- extra records (not the headline): full_update = false, the all-ice regime, PISM's default ranges ----
    extras = {}
    if not args.no_extras and full and args.regime == "dome":
        extras["full_update_false"] = measure(R, args.steps, False)
        extras["full_update_false"]["algorithmic_bytes_per_column"] = algorithmic_bytes_per_column(Mz, False)
        for _ in range(2):
            R.step(True)  # u, v current again for what follows
        if differs_from_uniform:
            Ru = Rank(M, Mz, args.regime, G.decompose(M, M, N, Nx_, Ny_), rank, local_rank, stream, args)
            extras["uniform_ranges"] = measure(Ru, args.steps, True)
            extras["uniform_ranges"]["decomposition"] = "%dx%d, PISM's default ownership ranges (-procs_x %s -procs_y %s)" % (
                Nx_, Ny_, ",".join(map(str, uniform_x)), ",".join(map(str, uniform_y)))
            assert extras["uniform_ranges"]["D_max"] == dmax
            del Ru
        if not multi:
            Ma = min(M, 2048)
            La = (Ma - 1) / 2.0 * 5000.0
            Ra = Rank(Ma, Mz, "allice", [G.Grid(Ma, Ma, Mz, La, La, 4000.0).whole()], 0, local_rank, stream, args)
            extras["allice_regime"] = measure(Ra, args.steps, True)
            extras["allice_regime"]["workload"] = ("dome %dx%dx%d with the margin at 1.5 Lx: every column carries ice "
                                                   "(the compute-heavy regime of the fused kernel)" % (Ma, Ma, Mz))
            del Ra
        torch.cuda.empty_cache()

    vertical = None
    consumers = None
    if not args.no_consumers and full and not multi:
        def w_step():
            R.check(lib.siafd_b200_compute_vertical_velocity(sia.handle, 0, 0))
        ms_w, _ = timed(w_step, args.steps, 3)
        bw = 24 * Mz  # read u, v once, write w
        vertical = {"kernel": "k_vvel_slab (w + fused 3D CFL maxima)", "ms": ms_w / args.steps,
                    "algorithmic_bytes_per_column": bw,
                    "achieved_GBps": bw * cols_total / (ms_w / args.steps / 1e3) / 1e9,
                    "frac": bw * cols_total / (ms_w / args.steps / 1e3) / 1e9 / peak}
        # the rest of a mass-continuity time step, device-resident
        out8 = (C.c_double * 8)()
        smb = torch.zeros((patch.ym, patch.xm), dtype=torch.float64, device=dev)
        R.check(lib.siafd_b200_bind(sia.handle, F["smb"], smb.data_ptr()))

        def cfl_fused():
            R.check(lib.siafd_b200_compute_vertical_velocity(sia.handle, 0, 0))
            R.check(lib.siafd_b200_cfl(sia.handle, 1.9e9, 1, out8))
        ms_cf, _ = timed(cfl_fused, args.steps, 3)

        def mass_step():  # dt = 0: the geometry stays what it is, the traffic is the same
            R.check(lib.siafd_b200_mass_flow_step(sia.handle, 0.0))
            R.check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
            R.check(lib.siafd_b200_mass_source_step(sia.handle, 0.0, 910.0, 0))
            R.check(lib.siafd_b200_ensure_consistency(sia.handle, 1))
        ms_m, _ = timed(mass_step, args.steps, 3)

        def heat_step():
            R.check(lib.siafd_b200_compute_strain_heating(sia.handle, 2, 3.0, 1.0))  # gpbld, n = 3, e = 1 (ssa defaults)
        ms_h, _ = timed(heat_step, args.steps, 3)
        bh = 32 * Mz  # read E, u, v, write Sigma
        consumers = {"vertical_velocity_plus_cfl_ms": ms_cf / args.steps, "cfl3d_dt_s": out8[0],
                     "strain_heating": {"kernel": "k_strain_heating<gpbld>", "ms": ms_h / args.steps,
                                        "algorithmic_bytes_per_column": bh,
                                        "achieved_GBps": bh * cols_total / (ms_h / args.steps / 1e3) / 1e9,
                                        "frac": bh * cols_total / (ms_h / args.steps / 1e3) / 1e9 / peak,
                                        "note": "FP64-bound where there is ice (one exp, one cbrt per level), "
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
            # the drop-in call: host arrays in (ghosts valid, as PISM's are), host arrays out, one call per rank
            R.check(lib.siafd_b200_update(sia.handle, C.byref(cin), C.byref(cout), 1 if full else 0))
            return lib.siafd_b200_max_diffusivity(sia.handle)

        b0, b1 = (C.c_int64(), C.c_int64()), (C.c_int64(), C.c_int64())
        barrier()   # (pinning gigabytes of host memory takes the ranks of one host seconds apart)
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
               "api": "siafd_b200_update(host pointers), one call per rank: every rank streams its own patch over its "
                      "own PCIe link (row bands, only the parts near ice; on one rank only the levels up to the thickest "
                      "ice nearby, the host replicates u, v above), ghost updates between the GPUs",
               "host_memory": "pinned"}
        assert dmax_e == dmax, (dmax_e, dmax)
        if full:  # the host arrays are the device-resident result, bit for bit (ghosts included)
            R.step(True)
            for name in ("u", "v", "flux"):
                chunk = 256
                for j0 in range(0, host[name].shape[0], chunk):
                    assert torch.equal(host[name][j0:j0 + chunk].to(dev), fields[name][j0:j0 + chunk]), (name, j0)
            e2e["verified"] = "host u, v, flux == device-resident u, v, flux (bitwise, ghosts included)"
        del host

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    cpu_tiles = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        nthreads = os.cpu_count() or 1
        v, sec, steps_c, sample = cpu_arm(cpu_domain(args), Mz, nthreads, args.cpu_seconds, full=full)
        cpu = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample}
        if not args.no_extras:
            v2, _, _, sample2 = cpu_arm(256, Mz, nthreads, 4.0, full=full, tiles=True)
            cpu_tiles = {"value": v2, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample2}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, "%dx%d (%s)" % (patch.Nx, patch.Ny, ranges_note)),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "D_max": dmax, "checksum": checksum, "extras": extras,
            "vertical_velocity": vertical, "consumers": consumers, "cpu_baseline_independent_tiles": cpu_tiles,
            "step": "siafd_b200_update_decomposed: one CUDA-graph launch of %d kernels per step, one host sync" %
                    (launches // max(args.steps, 1)),
            "halo_transport": ("stores into CUDA-IPC-mapped neighbour arrays over NVLink, issued by the producing kernels "
                               "(h_x, h_y: gradient kernel; u, v: fused kernel); D_max / status reduced through the same "
                               "pads; no NCCL on the data path") if multi else "periodic self-wrap inside the kernels",
            "input_generation_s": R.t_gen,
        }
        print(json.dumps(line))
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
