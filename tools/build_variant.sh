#!/bin/bash
# build a tuning variant of libfhe_b200.so that differs in ONE translation unit:
#   tools/build_variant.sh <name> <source.cu> <extra nvcc flags...>   -> fhe_study_b200/variants/lib_<name>.so
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p fhe_study_b200/variants fhe_study_b200/_build/var_$name
o=fhe_study_b200/_build/var_$name/${src%.cu}.o
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --fmad=false "$@" -c fhe_study_b200/csrc/$src -o $o
objs=$(ls fhe_study_b200/_build/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o fhe_study_b200/variants/lib_$name.so $objs $o -gencode arch=compute_100a,code=sm_100a -lcuda
echo fhe_study_b200/variants/lib_$name.so
