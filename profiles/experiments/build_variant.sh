#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: libscmc_b200.so with ONE source (SRC=scmc_tma.cu by default, e.g. SRC=scmc_core.cu) compiled under extra
# flags -> build_variants/NAME/libscmc_b200.so; the other objects are reused from the in-tree build.  Select with SCMC_LIB=build_variants/NAME/libscmc_b200.so
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
src=${SRC:-scmc_tma.cu}
mkdir -p build_variants/$name
P=semiclassicalmc.jl_b200
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -use_fast_math -Xcompiler -fPIC "$@" -c $P/csrc/$src -o build_variants/$name/variant.o
objs=""
for s in scmc_core.cu scmc_drivers.cu scmc_resident.cu scmc_tma.cu scmc_multi.cu; do
  if [ "$s" = "$src" ]; then objs="$objs build_variants/$name/variant.o"; else objs="$objs $P/build/$s.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/$name/libscmc_b200.so $objs -ldl
echo build_variants/$name/libscmc_b200.so
