#!/bin/bash
# round 2: quick parity + timing of the fused kernel (all-ice and dome, 2048^2)
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -x -q -m gpu > gpurun_out/r2_${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_${TAG}_tests.log
for R in allice dome; do
python bench.py --size 2048 --regime $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers > gpurun_out/r2_${TAG}_${R}_2048.json 2> gpurun_out/r2_${TAG}_${R}.err
done
python bench.py --size 2048 --regime allice --flux-only --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers > gpurun_out/r2_${TAG}_allice_flux_2048.json 2>> gpurun_out/r2_${TAG}_allice.err
tail -3 gpurun_out/r2_${TAG}_tests.log
python - <<P
import json
for r in ("allice","dome","allice_flux"):
    d=json.load(open("gpurun_out/r2_${TAG}_%s_2048.json"%r)); print(r, "step %.3f ms kernel %.3f ms frac %.3f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
P
