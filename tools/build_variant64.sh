#!/bin/bash
# Dev: A/B build of the float64 walk TU with extra flags:  tools/build_variant64.sh <name> <flags...>  -> variants/liblt_<name>.so
set -e
cd "$(dirname "$0")/../light_transport_b200/csrc"
name=$1; shift
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 128 -fmad=false "$@" -c lt_walk_f64.cu -o ../../variants/walk_f64_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/liblt_$name.so lt_api.o lt_lbvh.o lt_medium.o lt_walk_f32.o ../../variants/walk_f64_$name.o -lcudart
echo built variants/liblt_$name.so
