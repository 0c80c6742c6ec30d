#!/bin/bash
# variants of the fused kernel: ice-free and dome at 4096^2
mkdir -p gpurun_out
cp pism_b200/libsiafd_b200.so /tmp/lib_orig.so
for V in "$@"; do
  cp variants/lib_$V.so pism_b200/libsiafd_b200.so
  for R in icefree dome; do
    python bench.py --regime $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_v3_${V}_${R}.json 2> gpurun_out/r2_v3_${V}_${R}.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_v3_${V}_${R}.json').read().strip().split('\n')[-1]); print('$V $R', 'step %.3f ms kernel %.3f ms frac %.3f Dmax %r'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['D_max']))
except Exception as e: print('$V $R failed', e)"
  done
done
cp /tmp/lib_orig.so pism_b200/libsiafd_b200.so
