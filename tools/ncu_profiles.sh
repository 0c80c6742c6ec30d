#!/bin/bash
# round 2 profiles: (1) the launch list of the default bench command (our kernels only), (2) ncu --set full of the fused
# kernel on the 4096^2 dome (DRAM traffic per launch), on the 2048^2 dome and in the all-ice regime (stall breakdowns), of
# the gradient pass and of the two 3D consumers.  The reports are summarised on the box (tools/ncu_summary.py) and
# deleted: gpurun_out/ must stay under 64 MiB.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
cap() { # tag, kernel regex, bench args...
  TAG=$1; K=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/$TAG -f python bench.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/$TAG.ncu-rep 80 > gpurun_out/ncu_r02_${TAG}_summary.txt 2>&1
  ncu -i /tmp/$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,v=r[0],r[2]
m=dict(zip(h,v))
print({k:m[k] for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','lts__t_bytes.sum') if k in m})" > gpurun_out/ncu_r02_${TAG}_dram.txt
  if [ "$SRC" = "1" ]; then ncu -i /tmp/$TAG.ncu-rep --page source --csv > gpurun_out/ncu_r02_${TAG}_source.csv 2>/dev/null; fi
  rm -f /tmp/$TAG.ncu-rep
}
B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
cap k_sia_slab_4096 k_sia_slab $B --no-consumers
cap k_sia_slab_2048 k_sia_slab --size 2048 $B --no-consumers
SRC=1 cap allice_k_sia_slab_2048 k_sia_slab --size 2048 --regime allice $B --no-consumers
cap k_grad_haseloff_4096 k_grad_haseloff $B --no-consumers
cap k_vvel_slab_2048 k_vvel_slab --size 2048 $B
cap k_strain_heating_2048 k_strain_heating --size 2048 $B
ls -la gpurun_out/ | tail -20; cat gpurun_out/ncu_r02_*_dram.txt
