#!/bin/bash
# Dev: A/B build of the float32 walk TU with extra flags:  tools/build_variant.sh <name> <flags...>
# -> variants/liblt_<name>.so   (run with LT_B200_LIB=variants/liblt_<name>.so)
set -e
cd "$(dirname "$0")/../light_transport_b200/csrc"
name=$1; shift
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 128 --use_fast_math "$@" -c lt_walk_f32.cu -o ../../variants/walk_f32_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/liblt_$name.so lt_api.o lt_lbvh.o lt_medium.o ../../variants/walk_f32_$name.o lt_walk_f64.o -lcudart
echo built variants/liblt_$name.so
