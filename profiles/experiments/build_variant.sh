#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: libscmc_b200.so with scmc_tma.cu compiled under extra flags -> build_variants/NAME/libscmc_b200.so
# (the other objects are reused from the in-tree build; select the library with SCMC_LIB=build_variants/NAME/libscmc_b200.so)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p build_variants/$name
P=semiclassicalmc.jl_b200
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -use_fast_math -Xcompiler -fPIC "$@" -c $P/csrc/scmc_tma.cu -o build_variants/$name/scmc_tma.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/$name/libscmc_b200.so $P/build/scmc_core.cu.o $P/build/scmc_drivers.cu.o $P/build/scmc_resident.cu.o $P/build/scmc_multi.cu.o build_variants/$name/scmc_tma.o -ldl
echo build_variants/$name/libscmc_b200.so
