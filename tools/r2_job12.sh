#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py -x -q -m gpu -k "host_arrays" > gpurun_out/r2_hostcomm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hostcomm.log
tail -30 gpurun_out/r2_hostcomm.log
