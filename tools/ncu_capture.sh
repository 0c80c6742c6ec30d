#!/bin/bash
# ncu --set full of one kernel of a bench.py run, summarised on the box: tools/ncu_capture.sh TAG KERNEL_REGEX BLOCK bench-args...
mkdir -p gpurun_out
TAG=$1; K=$2; BLK=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/$TAG -f python bench.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/$TAG.ncu-rep $BLK > gpurun_out/ncu_r02_${TAG}_summary.txt 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv > gpurun_out/ncu_r02_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/ncu_r02_${TAG}_*; head -22 gpurun_out/ncu_r02_${TAG}_summary.txt
