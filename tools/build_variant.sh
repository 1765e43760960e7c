#!/bin/bash
# A/B build of one translation unit with extra nvcc flags: tools/build_variant.sh NAME FILE.cu FLAGS...
# -> build/variants/libpodsdf_NAME.so (the other objects come from the regular in-tree build)
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/pods-digital-filter_b200/csrc
mkdir -p $root/build/variants
obj=$root/build/variants/${name}_${src%.cu}.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $csrc/$src -o $obj
others=$(ls $csrc/_obj/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o $root/build/variants/libpodsdf_$name.so $obj $others -gencode arch=compute_100a,code=sm_100a -ldl
echo built $root/build/variants/libpodsdf_$name.so
