#!/bin/bash
mkdir -p gpurun_out
cd tests/host_cpp
./pismv_test_C -Mx 31 -My 31 -y 600 -trace 1 > ../../gpurun_out/trace1.txt 2>&1
P=/dev/shm/trace_$$
./pismv_test_C -Mx 31 -My 31 -y 600 -trace 1 -rank 0 -size 2 -prefix $P > ../../gpurun_out/trace2.txt 2>&1 &
./pismv_test_C -Mx 31 -My 31 -y 600 -trace 1 -rank 1 -size 2 -prefix $P > ../../gpurun_out/trace2_r1.txt 2>&1
wait
cd ../..
head -12 gpurun_out/trace1.txt; echo ---; head -12 gpurun_out/trace2.txt; echo; tail -3 gpurun_out/trace2_r1.txt
