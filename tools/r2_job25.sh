#!/bin/bash
mkdir -p gpurun_out
./tools/memcpy2d_bw > gpurun_out/memcpy2d_bw_r02.txt 2>&1; cat gpurun_out/memcpy2d_bw_r02.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_comm.py tests/test_gpu_decomposition.py tests/test_golden_fixtures.py tests/test_gpu_pismv.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_t5_dome_4096.json 2> gpurun_out/r2_t5.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_t5_dome_4096.json").read().strip().split('\n')[-1]); print("step %.3f ms kernel %.3f ms frac %.3f launches %d"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"]), d["roofline"]["step_breakdown_ms"])
P
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_grad_haseloff -c 1 --csv --log-file gpurun_out/r2_t5_grad.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-consumers --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_t5_grad.csv | cut -d, -f5,13- | tail -7
