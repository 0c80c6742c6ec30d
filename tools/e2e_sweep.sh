#!/bin/bash
# e2e (host buffers) for a few settings of the sparse host path: "<fill threads> <band>"
for cfg in "$@"; do
  set -- $cfg
  SIAFD_B200_FILL_THREADS=$1 SIAFD_B200_BAND=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-consumers 2>gpurun_out/e2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('threads $1 band $2:', round(e['ms_per_step'],1), 'ms', round(e['value']/1e6,1), 'M col/s  h2d', round(e['h2d_bytes_per_step']/1e9,2), 'GB d2h', round(e['d2h_bytes_per_step']/1e9,2), 'GB', e.get('verified'))" || tail -3 gpurun_out/e2e.err
done
