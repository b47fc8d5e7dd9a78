#!/bin/bash
# Dev: A/B build of the library with extra compile flags (-D...) on the walk and API units:
#   tools/build_variant.sh <name> <flags...>   ->   variants/liblt_<name>.so   (tools take it with --lib)
set -e
cd "$(dirname "$0")/../light_transport_b200/csrc"
name=$1; shift
mkdir -p ../../variants
C="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 128"
nvcc $C --use_fast_math "$@" -c lt_walk_f32.cu -o ../../variants/walk_f32_$name.o &
nvcc $C -fmad=false "$@" -c lt_walk_f64.cu -o ../../variants/walk_f64_$name.o &
nvcc $C "$@" -c lt_api.cu -o ../../variants/api_$name.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/liblt_$name.so ../../variants/api_$name.o lt_lbvh.o lt_medium.o lt_comm.o ../../variants/walk_f32_$name.o ../../variants/walk_f64_$name.o -lcudart -ldl
echo built variants/liblt_$name.so
