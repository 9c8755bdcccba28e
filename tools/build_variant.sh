#!/bin/bash
# Developer tool: build an A/B variant of libmojo_b200.so with extra nvcc flags for ONE source file, reusing the other
# objects of the regular build.   tools/build_variant.sh <name> <file.cu> <flags...>
#   -> mojo_opset_b200/libmojo_b200_<name>.so   (select with MOJO_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p build/variant_$name
obj=build/variant_$name/$(basename ${src%.cu}).o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden -DMOJO_B200_BUILD "$@" -c mojo_opset_b200/csrc/$src -o $obj
others=$(ls build/obj/*.o | grep -v "/$(basename ${src%.cu}).o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o mojo_opset_b200/libmojo_b200_$name.so $obj $others \
  -cudart static -Xlinker --exclude-libs,ALL
echo mojo_opset_b200/libmojo_b200_$name.so
