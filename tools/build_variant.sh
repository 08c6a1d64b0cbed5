#!/bin/bash
# tools/build_variant.sh NAME "-DMPSK_..." : build an experimental variant of the
# library as mp-sort_b200/variants/libmpsort-b200.NAME.so (select it at run time with
# MPSORT_LIB=<path>). Development aid for kernel tuning sweeps; not part of the product.
set -e
NAME=$1; DEFS=$2
cd "$(dirname "$0")/../mp-sort_b200"
mkdir -p variants/build_$NAME
nvcc -O3 -lineinfo -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $DEFS \
    -I../include -Icsrc -c csrc/mpsort_kernels.cu -o variants/build_$NAME/mpsort_kernels.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libmpsort-b200.$NAME.so \
    variants/build_$NAME/mpsort_kernels.o build/mpsort_host.o build/mpsort_comm.o build/mpsort_layout.o build/mpsort_util.o -lnccl -lpthread
echo built variants/libmpsort-b200.$NAME.so
