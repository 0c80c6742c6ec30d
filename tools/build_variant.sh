#!/bin/bash
# tools/build_variant.sh TAG "<nvcc -D flags>": a development build of the library with another fused-kernel variant
# (gpbld only, SLAB_DEV), linked against the in-tree objects -> variants/lib_TAG.so (git-ignored; travels with gpurun)
set -e
TAG=$1; shift
cd "$(dirname "$0")/../pism_b200/csrc"
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p ../../variants
$NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -DSLAB_DEV "$@" -c siafd_slab.cu -o /tmp/slab_$TAG.o 2> ../../variants/ptxas_$TAG.log
$NVCC $ARCH -shared -o ../../variants/lib_$TAG.so siafd_kernels.o /tmp/slab_$TAG.o siafd_mass.o siafd_capi.o siafd_comm.o
grep -A2 "k_sia_slabILi2ELb1ELi16ELi[48]ELb1" ../../variants/ptxas_$TAG.log | grep -i "registers\|spill" | tr '\n' ' '; echo " <- $TAG"
