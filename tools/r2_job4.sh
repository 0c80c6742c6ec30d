#!/bin/bash
# round 2: the communicator path on one GPU (tests), then a short N=1 bench
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests/test_gpu_comm.py -x -q -m gpu > gpurun_out/r2_comm_tests.log 2>&1; echo "comm tests rc=$?" >> gpurun_out/r2_comm_tests.log
tail -25 gpurun_out/r2_comm_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n1.err
python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("step %.3f ms kernel %.3f frac %.3f launches %d e2e %.1f ms"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["ms_per_step"]))
print(json.dumps(d["extras"])[:1500])
print(d["cpu_baseline"])
P
