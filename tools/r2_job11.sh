#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2_gpu_tests.log
tail -15 gpurun_out/r2_gpu_tests.log
