#!/usr/bin/env bash
# Builds a variant of the library with extra nvcc defines: tools/build_variant.sh <suffix> <defines...>
# -> incagg_gnn_b200/csrc/libincagg_b200_<suffix>.so   (select it with INCAGG_B200_LIB=<path>)
set -euo pipefail
SUF=$1; shift
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP=$(mktemp -d)
mkdir -p $TMP/pkg/csrc $TMP/include
cp $ROOT/incagg_gnn_b200/csrc/*.cu $ROOT/incagg_gnn_b200/csrc/*.cuh $ROOT/incagg_gnn_b200/csrc/Makefile $ROOT/incagg_gnn_b200/csrc/metis_shim.c $TMP/pkg/csrc/
cp $ROOT/include/*.h $TMP/include/
make -C $TMP/pkg/csrc -j8 libincagg_b200.so NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xptxas -v $*" > /dev/null
cp $TMP/pkg/csrc/libincagg_b200.so $ROOT/incagg_gnn_b200/csrc/libincagg_b200_$SUF.so
rm -rf $TMP
echo built incagg_gnn_b200/csrc/libincagg_b200_$SUF.so
