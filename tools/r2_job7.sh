#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_comm_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_comm_tests.log
tail -4 gpurun_out/r2_comm_tests.log
bash tools/r2_job6.sh
