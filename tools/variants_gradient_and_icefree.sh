#!/bin/bash
# gradient-pass variants (step breakdown of a 4096^2 dome step) and write-only-regime diagnostics of the fused kernel
mkdir -p gpurun_out
cp pism_b200/libsiafd_b200.so /tmp/lib_orig.so
for V in g4b4 g4b6 g2b6 g2b8 g1b8; do
  cp variants/lib_$V.so pism_b200/libsiafd_b200.so
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_v_${V}.json 2> gpurun_out/r2_v_${V}.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_v_${V}.json').read().strip().split('\n')[-1]); print('$V', 'step %.3f ms kernel %.3f gradient pass %.3f ms'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['step_breakdown_ms']['gradient_pass']))
except Exception as e: print('$V failed', e)"
done
for V in cur nodq noflags nodqflags; do
  cp variants/lib_$V.so pism_b200/libsiafd_b200.so
  python bench.py --regime icefree --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-consumers --no-extras > gpurun_out/r2_v_${V}.json 2> gpurun_out/r2_v_${V}.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_v_${V}.json').read().strip().split('\n')[-1]); print('$V icefree', 'step %.3f ms kernel %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms']))
except Exception as e: print('$V failed', e)"
done
cp /tmp/lib_orig.so pism_b200/libsiafd_b200.so
