#!/bin/bash
mkdir -p gpurun_out
for V in nc16 nc8; do
  if [ $V = nc8 ]; then export SIAFD_B200_NC=8; fi
  for R in allice dome; do
    python bench.py --size 2048 --regime $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_${V}_${R}_2048.json 2> gpurun_out/r2_${V}_${R}.err
    python -c "
import json
d=json.loads(open('gpurun_out/r2_${V}_${R}_2048.json').read().strip().split('\n')[-1]); print('$V $R', 'step %.3f ms kernel %.3f ms frac %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  done
done
