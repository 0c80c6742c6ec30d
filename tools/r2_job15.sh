#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact_columns.py -x -q -s -m gpu > gpurun_out/r2_exact_cols.log 2>&1; echo "rc=$?" >> gpurun_out/r2_exact_cols.log
tail -30 gpurun_out/r2_exact_cols.log
