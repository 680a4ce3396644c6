#!/bin/bash
# devbuild.sh <name> [nvcc flags...]: an experimental build of the library into exp_libs/lib_<name>.so (run with TR_LIB=...).
# Only crew2_launch.cu is recompiled with the extra flags; the other objects come from the regular build (transrot_b200/build/).
name=$1; shift
set -e
[ -f transrot_b200/build/multi.o ] || python transrot_b200/build.py > /dev/null
mkdir -p exp_libs
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DTR_CREW2_DEV "$@" -c transrot_b200/csrc/crew2_launch.cu -o exp_libs/crew2_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o exp_libs/lib_$name.so transrot_b200/build/transrot_b200.o transrot_b200/build/crew_launch.o exp_libs/crew2_$name.o transrot_b200/build/team_pc_launch.o transrot_b200/build/multi.o -ldl
echo exp_libs/lib_$name.so
