#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> [-Dflags...]   -> variants/libfemb_<name>.so (the other objects come from build/)
set -e
NAME=$1; SRC=$2; shift 2
PKG=extendablefembase.jl_b200
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr -I include -I $PKG/csrc "$@" -Xptxas -v -c $PKG/csrc/$SRC -o variants/${NAME}.o 2> variants/${NAME}.ptxas
OBJS=$(ls $PKG/build/*.o | grep -v "/${SRC%.cu}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libfemb_${NAME}.so $OBJS variants/${NAME}.o -ldl
grep -A3 "k_h1_gatherILi3" variants/${NAME}.ptxas | grep "registers\|spill" | tr '\n' ' '; echo
