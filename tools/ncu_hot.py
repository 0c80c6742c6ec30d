#!/usr/bin/env python3
"""Hot spots of an ncu source-page CSV: basic blocks by executed share and stall-sample share, and the top
stalled instructions with their dominant stall reason.  Usage: tools/ncu_hot.py source.csv [ntop]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = rows[2:]
totx = sum(int(r[iex] or 0) for r in data); tots = sum(int(r[isamp] or 0) for r in data)
runs = []; start = 0
for n in range(1, len(data) + 1):
    if n == len(data) or data[n][iex] != data[start][iex]:
        ex = int(data[start][iex] or 0) * (n - start); s = sum(int(data[m][isamp] or 0) for m in range(start, n))
        runs.append((start, n, ex, s)); start = n
print("basic blocks (>=0.4%% exec or >=0.6%% samples); total exec %d samples %d" % (totx, tots))
for a, b, ex, s in runs:
    if ex > 0.004 * totx or s > 0.006 * tots:
        ops = {}
        for m in range(a, b):
            t = data[m][isrc].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:6]
        print("%5d-%-5d n=%3d each=%.3f%% exec=%5.1f%% samp=%5.1f%%  %s" % (a, b, b - a, 100.0 * int(data[a][iex] or 0) / totx, 100.0 * ex / totx, 100.0 * s / tots, " ".join("%s:%d" % kv for kv in top)))
print("\ntop stalled instructions")
top = sorted(range(len(data)), key=lambda n: -int(data[n][isamp] or 0))[:ntop]
for n in sorted(top):
    r = data[n]
    st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("%5d %-70s %6s %5.2f%%  %s" % (n, r[isrc][:70], r[isamp], 100.0 * int(r[isamp]) / tots, " ".join("%s:%d" % (b, a) for a, b in st)))
