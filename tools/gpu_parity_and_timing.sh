#!/bin/bash
# round 2: parity of the trimmed fused kernel + timing (2048^2 all-ice / dome, 4096^2 dome) + executed-instruction counts
TAG=${1:-t1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_decomposition.py tests/test_gpu_comm.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/r2_${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_${TAG}_tests.log
tail -4 gpurun_out/r2_${TAG}_tests.log
for R in allice dome; do
python bench.py --size 2048 --regime $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${TAG}_${R}_2048.json 2> gpurun_out/r2_${TAG}_${R}.err
done
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${TAG}_dome_4096.json 2> gpurun_out/r2_${TAG}_dome4096.err
python bench.py --regime icefree --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${TAG}_icefree_4096.json 2>> gpurun_out/r2_${TAG}_allice.err
python bench.py --size 2048 --regime allice --flux-only --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${TAG}_alliceflux_2048.json 2>> gpurun_out/r2_${TAG}_allice.err
python - <<P
import json
for r in ("allice_2048","dome_2048","dome_4096","alliceflux_2048","icefree_4096"):
    try:
        d=json.loads(open("gpurun_out/r2_${TAG}_%s.json"%r).read().strip().split('\n')[-1]); print(r, "step %.3f ms kernel %.3f ms frac %.3f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]), d["roofline"].get("step_breakdown_ms"))
    except Exception as e: print(r, "failed", e)
P
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sia_slab -c 1 --csv --log-file gpurun_out/r2_${TAG}_inst_allice.csv python bench.py --size 2048 --regime allice --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-consumers --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_${TAG}_inst_allice.csv | cut -d, -f5,13- | tail -4
