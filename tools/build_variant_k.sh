#!/bin/bash
# tools/build_variant_k.sh TAG "<nvcc -D flags>": the library with another build of siafd_kernels.cu -> variants/lib_TAG.so
set -e
TAG=$1; shift
cd "$(dirname "$0")/../pism_b200/csrc"
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p ../../variants
$NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c siafd_kernels.cu -o /tmp/kern_$TAG.o 2> ../../variants/ptxas_k_$TAG.log
$NVCC $ARCH -shared -o ../../variants/lib_$TAG.so /tmp/kern_$TAG.o siafd_slab.o siafd_mass.o siafd_capi.o siafd_comm.o
grep -A2 "k_grad_haseloff_quadILb0ELb1" ../../variants/ptxas_k_$TAG.log | grep -i "registers\|spill" | tr '\n' ' '; echo " <- $TAG"
