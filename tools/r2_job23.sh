#!/bin/bash
mkdir -p gpurun_out
./tools/store_pattern 4096 > gpurun_out/store_pattern_r02.txt 2>&1; cat gpurun_out/store_pattern_r02.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_comm.py tests/test_gpu_decomposition.py tests/test_zz_gpu_host_pipeline.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_t3_dome_4096.json 2> gpurun_out/r2_t3.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_t3_dome_4096.json").read().strip().split('\n')[-1]); print("step %.3f ms kernel %.3f ms frac %.3f launches %d e2e %.1f ms"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["ms_per_step"]))
P
