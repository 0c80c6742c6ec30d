#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel) into text: headline metrics, stall mix, and where the executed
instructions / stall samples sit in the SASS.  Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [block]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 80


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in m:
        print("%-70s %s %s" % (k, m[k][0], m[k][1]))
print("\nstall reasons (warps stalled per issue-active cycle):")
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v[0]))
      for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and v[0]]
for name, v in sorted(st, key=lambda x: -x[1])[:10]:
    print("  %-24s %.3f" % (name, v))

src = page("source")
h2 = src[1]
ia, isrc, isamp, iex = h2.index("Address"), h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
data = [(r[isrc], int(r[isamp] or 0), int(r[iex] or 0)) for r in src[2:] if len(r) > iex]
tot_ex, tot_s = sum(d[2] for d in data), sum(d[1] for d in data)
print("\nSASS: %d instructions, %d warp-instructions executed, %d stall samples" % (len(data), tot_ex, tot_s))
ops_all = {}
for d in data:
    t = d[0].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops_all[op] = ops_all.get(op, 0) + d[2]
print("executed mix: " + "  ".join("%s %.1f%%" % (k, 100.0 * v / tot_ex) for k, v in sorted(ops_all.items(), key=lambda x: -x[1])[:18]))
print("\nblock     exec%  samp%  top ops")
for b in range(0, len(data), blk):
    ex = sum(d[2] for d in data[b:b + blk])
    s = sum(d[1] for d in data[b:b + blk])
    if ex < 0.004 * tot_ex and s < 0.004 * tot_s:
        continue
    ops = {}
    for d in data[b:b + blk]:
        t = d[0].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + d[2]
    top = sorted(ops.items(), key=lambda x: -x[1])[:7]
    print("%5d-%-5d %5.1f %6.1f  %s" % (b, b + blk, 100.0 * ex / tot_ex, 100.0 * s / tot_s,
                                       " ".join("%s:%.1f" % (k, 100.0 * v / tot_ex) for k, v in top)))
