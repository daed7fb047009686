#!/bin/bash
# Experiment build of libhrmgpu.so: one source recompiled with extra -D flags, linked with the other (already built) objects.
#   tools/build_variant.sh <name> <source.cu> [-DFLAG ...]   ->  hrm_b200/lib/exp/libhrmgpu_<name>.so   (use: HRMGPU_LIB=<that path>)
set -e
NAME=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/hrm_b200/csrc/build
mkdir -p $ROOT/hrm_b200/lib/exp $OBJ/exp
BASE=$(basename $SRC .cu)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-Wall --expt-relaxed-constexpr "$@" -c $ROOT/hrm_b200/csrc/$BASE.cu -o $OBJ/exp/${BASE}_$NAME.o
OTHERS=$(ls $OBJ/*.o | grep -v "/$BASE.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/hrm_b200/lib/exp/libhrmgpu_$NAME.so $OBJ/exp/${BASE}_$NAME.o $OTHERS
echo $ROOT/hrm_b200/lib/exp/libhrmgpu_$NAME.so
