#!/bin/bash
# default bench command: the JSON line, then the ncu launch list of the SAME command (one pass, no replay)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 2000 --csv --log-file gpurun_out/launches_r01_default.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_default.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_r01_default.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > mv:
        name = r[kn].split("(")[0][:60]
        t[name][0] += 1
        t[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in t.values())
for k, v in sorted(t.items(), key=lambda x: -x[1][1]):
    print("%-62s n=%4d  %10.1f us  %5.1f %%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
PY
