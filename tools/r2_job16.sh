#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k pointwise > gpurun_out/r2_pointwise.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pointwise.log
grep -E "pointwise relative|passed|failed|rc=|assert" gpurun_out/r2_pointwise.log | head -20
