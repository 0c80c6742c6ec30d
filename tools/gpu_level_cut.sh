#!/bin/bash
# the level cut of the host-buffer path on the GPU: (optionally its parity tests, then) the knob sweep at 4096^2
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 300 python -m pytest tests/test_zz_gpu_host_pipeline.py -x -q -m gpu > gpurun_out/r2_level_cut_tests.log 2>&1
  echo "tests rc=$?" >> gpurun_out/r2_level_cut_tests.log; tail -4 gpurun_out/r2_level_cut_tests.log
fi
timeout 420 python tools/e2e_sweep.py --comm --steps 2 --settings "${1:-LEVEL_CUT=0,,CUT_COLS=64,CUT_COLS=256}" \
  > gpurun_out/e2e_sweep_r02g.json 2> gpurun_out/e2e_sweep_r02g.err
echo "sweep rc=$?"; cut -c1-200 gpurun_out/e2e_sweep_r02g.json | head -20; grep -A4 "trace" gpurun_out/e2e_sweep_r02g.err | cut -c1-400
