#!/bin/bash
# Build a variant of libkeep_b200.so that differs from the in-tree library in ONE translation unit compiled with extra
# -D flags (timing experiments, A/B inside one gpurun call through KEEPB200_LIB=<path>):
#   bash tools/build_variant.sh <name> <file.cu> [-DFOO=1 ...]     ->  _ab/libkeep_b200_<name>.so
set -e
NAME=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/_ab
OBJ=$ROOT/_ab/${NAME}_$(basename $SRC .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I $ROOT/include "$@" -c $ROOT/keep_b200/csrc/$SRC -o $OBJ
OTHERS=$(ls $ROOT/keep_b200/_build/*.o | grep -v "/$(basename $SRC .cu).o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/_ab/libkeep_b200_$NAME.so $OTHERS $OBJ
echo $ROOT/_ab/libkeep_b200_$NAME.so
