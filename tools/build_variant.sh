#!/bin/bash
# usage: tools/build_variant.sh NAME file.cu [-DFLAG ...]   -> glue_factory_colon_b200/lib/var/NAME.so (one TU recompiled)
set -e
cd "$(dirname "$0")/../glue_factory_colon_b200"
name=$1; src=$2; shift 2
mkdir -p lib/var
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$src.cu -o lib/var/$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o lib/var/$name.so lib/var/$name.o $(ls lib/obj/*.o | grep -v "/$src.o")
echo lib/var/$name.so
