#!/bin/bash
# Experiment build of one kernel source: compiles csrc/<file>.cu with the given -D macros and links it with the other
# (already built) objects of csrc/build into tools/ubench/libnfk_<name>.bin.
#   tools/ubench/build_variant.sh <name> <file-without-.cu> [-DMACRO ...]
#   NFK_LIB=$PWD/tools/ubench/libnfk_<name>.bin python tools/bench_fused.py --modes fast
set -e
name=$1; file=$2; shift; shift
cd "$(dirname "$0")/../../normalizingflow_b200/csrc"
make -s -j8 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default \
     "$@" -c $file.cu -o /tmp/${file}_$name.o
objs=$(ls build/*.o | grep -v "/$file.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/ubench/libnfk_$name.bin $objs /tmp/${file}_$name.o
echo "built tools/ubench/libnfk_$name.bin ($file: $*)"
