#!/bin/bash
# One GPU call's worth of round-end measurements (1 GPU): bench line, consumers, launch list, ncu captures.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
tail -c 600 gpurun_out/bench_r01_final.json
python bench.py --steps 10 --warmup 3 --with-w --no-e2e --no-cpu-baseline > gpurun_out/bench_r01_consumers.json 2>> gpurun_out/bench_r01_final.err
python bench.py --steps 10 --warmup 3 --flux-only --no-e2e --no-cpu-baseline > gpurun_out/bench_r01_fluxonly.json 2>> gpurun_out/bench_r01_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv \
  python bench.py --steps 2 --warmup 3 --with-w --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for k in k_vvel_slab k_strain_heating k_sia_slab; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_r01_final_$k \
    python bench.py --size 2048 --steps 1 --warmup 1 --with-w --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
