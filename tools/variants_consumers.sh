#!/bin/bash
# variants of the consumers' kernels: parity of the w / CFL tests, then the consumers' timings at 4096^2
mkdir -p gpurun_out
cp pism_b200/libsiafd_b200.so /tmp/lib_orig.so
for V in "$@"; do
  cp variants/lib_$V.so pism_b200/libsiafd_b200.so
  timeout 600 python -m pytest tests/test_gpu_pismv.py tests/test_golden_fixtures.py tests/test_gpu_decomposition.py -x -q -m gpu 2>&1 | tail -1
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_v4_${V}.json 2> gpurun_out/r2_v4_${V}.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_v4_${V}.json').read().strip().split('\n')[-1]); print('$V', 'w %.3f ms (%.3f) w+cfl %.3f Sigma %.3f ms'%(d['vertical_velocity']['ms'], d['vertical_velocity']['frac'], d['consumers']['vertical_velocity_plus_cfl_ms'], d['consumers']['strain_heating']['ms']))
except Exception as e: print('$V failed', e)"
done
cp /tmp/lib_orig.so pism_b200/libsiafd_b200.so
