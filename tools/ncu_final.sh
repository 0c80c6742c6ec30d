#!/bin/bash
# ncu --set full of the three 3D kernels on the 2048^2 dome (one launch each), for profiles/
mkdir -p gpurun_out
for k in k_sia_slab k_vvel_slab k_strain_heating; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_r01_final_$k -f \
    python bench.py --size 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/prof_r01_final_*.ncu-rep
