#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-x}
for R in allice dome; do
python bench.py --size 2048 --regime $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${TAG}_${R}_2048.json 2> gpurun_out/r2_${TAG}_${R}.err
done
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${TAG}_dome_4096.json 2> gpurun_out/r2_${TAG}_dome4096.err
python - <<P
import json
for r in ("allice_2048","dome_2048","dome_4096"):
    d=json.load(open("gpurun_out/r2_${TAG}_%s.json"%r)); print(r, "step %.3f ms kernel %.3f ms frac %.3f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
P
