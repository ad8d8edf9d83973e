#!/bin/bash
# Ablation / experiment build of the fused layer kernel: compiles csrc/nsf_fused.cu with the given -D macros and
# links it with the other (already built) objects of csrc/build into tools/ubench/libnfk_<name>.bin.
#   tools/ubench/build_variant.sh <name> [-DMACRO ...]
#   NFK_LIB=$PWD/tools/ubench/libnfk_<name>.bin python tools/bench_fused.py --modes fast
set -e
name=$1; shift
cd "$(dirname "$0")/../../normalizingflow_b200/csrc"
make -s -j8 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default \
     "$@" -c nsf_fused.cu -o /tmp/nsf_fused_$name.o
objs=$(ls build/*.o | grep -v nsf_fused.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/ubench/libnfk_$name.bin $objs /tmp/nsf_fused_$name.o
echo "built tools/ubench/libnfk_$name.bin ($*)"
