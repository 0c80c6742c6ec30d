#!/bin/bash
# the level cut of the host-buffer path on the GPU: its parity tests, then the knob sweep at 4096^2 (one process)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_gpu_host_pipeline.py -x -q -m gpu > gpurun_out/r2_level_cut_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_level_cut_tests.log; tail -4 gpurun_out/r2_level_cut_tests.log
timeout 420 python tools/e2e_sweep.py --comm --steps 2 --settings "${1:-LEVEL_CUT=0,,CUT_COLS=64,CUT_COLS=256,CUT_COLS=512,REPL_THREADS=2,REPL_THREADS=6,REPL_THREADS=8;FILL_THREADS=6,FILL_THREADS=2,BAND=2}" \
  > gpurun_out/e2e_sweep_r02.json 2> gpurun_out/e2e_sweep_r02.err
echo "sweep rc=$?"; cut -c1-330 gpurun_out/e2e_sweep_r02.json | head -20; tail -5 gpurun_out/e2e_sweep_r02.err
