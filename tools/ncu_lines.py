#!/usr/bin/env python3
"""Per SOURCE LINE executed warp-instructions and stall samples of one kernel: joins the SASS listing of an ncu source
page (ncu -i rep --page source --csv, one row per instruction, in address order) with the line table of the same object
file (nvdisasm -gi).  Usage: tools/ncu_lines.py source.csv object.o 'mangled-name-prefix' [top]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

src_csv, obj, prefix = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + prefix))
lines = []
cur = None
for l in sass[start + 1:]:
    if l.startswith("//------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l):
        lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
ex, sm = collections.Counter(), collections.Counter()
st = collections.defaultdict(collections.Counter)
for ln, r in zip(lines, data):
    e, s = int(r[iex] or 0), int(r[isamp] or 0)
    ex[ln] += e
    sm[ln] += s
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            st[ln][h[6:]] += v
te, ts = sum(ex.values()), sum(sm.values())
print("total warp-instructions %d, samples %d" % (te, ts))
srcs = {}
def text(ln):
    if ln is None:
        return ""
    f = ln[0]
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return srcs[f][ln[1] - 1].strip()[:90] if ln[1] - 1 < len(srcs[f]) else ""
print("%6s %6s  %-22s %s" % ("exec%", "samp%", "line", "top stalls | source"))
for ln, e in sorted(ex.items(), key=lambda x: -(x[1] / te + sm[x[0]] / ts))[:top]:
    tops = " ".join("%s:%.1f" % (k, 100.0 * v / ts) for k, v in st[ln].most_common(3))
    print("%6.2f %6.2f  %-22s %s | %s" % (100.0 * e / te, 100.0 * sm[ln] / ts, "%s:%d" % ln if ln else "?", tops, text(ln)))
