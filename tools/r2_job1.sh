#!/bin/bash
# round 2, job 1: FP64 pipe microbenchmark; the all-ice regime of the fused kernel (timing + one ncu capture)
mkdir -p gpurun_out
./tools/fp64_peak > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err
python bench.py --size 2048 --regime allice --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers > gpurun_out/r2_allice_2048.json 2> gpurun_out/r2_allice_2048.err
python bench.py --size 2048 --regime allice --flux-only --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers > gpurun_out/r2_allice_2048_flux.json 2>> gpurun_out/r2_allice_2048.err
python bench.py --size 2048 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers > gpurun_out/r2_dome_2048.json 2>> gpurun_out/r2_allice_2048.err
ncu --set full --clock-control none --import-source on -k regex:k_sia_slab -c 1 -o gpurun_out/prof_r02_allice_k_sia_slab -f \
  python bench.py --size 2048 --regime allice --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-consumers > gpurun_out/ncu_allice.log 2>&1
ls -la gpurun_out/*.ncu-rep
