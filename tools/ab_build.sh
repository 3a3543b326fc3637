#!/bin/bash
# A/B build: tools/ab_build.sh <name> <source.cu> [nvcc flags...]  ->  tools/ab/lib_<name>.so
# Recompiles one source of the library with extra flags (e.g. -DOFK_BOX3=160) and links it with the objects of the
# regular build (AB_SRC=/path/to/variant.cu compiles another file in its place); (python -m oflibnumpy_b200.build first). Select a variant at run time with OFK_LIB_PATH=tools/ab/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p tools/ab
L=oflibnumpy_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -fmad=true "$@" \
    -I oflibnumpy_b200/csrc -c ${AB_SRC:-oflibnumpy_b200/csrc/$src} -o tools/ab/${src%.cu}_$name.o
objs=""
for o in $L/*.o; do
    if [ "$(basename $o)" == "${src%.cu}.o" ]; then objs="$objs tools/ab/${src%.cu}_$name.o"; else objs="$objs $o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/ab/lib_$name.so $objs
echo tools/ab/lib_$name.so
