#!/bin/bash
# build_exp.sh <name> <extra nvcc flags...>: an experimental build of the library into exp_libs/lib_<name>.so (use with TR_LIB=...)
name=$1; shift
mkdir -p exp_libs
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared --threads 0 "$@" -o exp_libs/lib_$name.so transrot_b200/csrc/transrot_b200.cu
