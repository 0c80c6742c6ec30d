#!/bin/bash
# tools/build_variant_m.sh TAG "<nvcc -D flags>": the library with another build of siafd_mass.cu -> variants/lib_TAG.so
set -e
TAG=$1; shift
cd "$(dirname "$0")/../pism_b200/csrc"
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p ../../variants
$NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c siafd_mass.cu -o /tmp/mass_$TAG.o 2> ../../variants/ptxas_m_$TAG.log
$NVCC $ARCH -shared -o ../../variants/lib_$TAG.so siafd_kernels.o siafd_slab.o /tmp/mass_$TAG.o siafd_capi.o siafd_comm.o
grep -A2 "k_vvel_slab" ../../variants/ptxas_m_$TAG.log | grep -i "registers\|spill" | tr '\n' ' '; echo " <- $TAG"
