#!/bin/bash
# Development: build an experimental variant of the library under build/exp/<name>/ with extra nvcc flags.
#   tools/build_exp.sh <name> "<extra flags>" [make vars, e.g. FAST=1]
# Use it with SGBM_B200_LIB=build/exp/<name>/pkg/libsgbm_b200.so (see _lib.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; flags=$2; shift 2
dst=$ROOT/build/exp/$name
mkdir -p $dst/pkg/csrc $dst/include
cp $ROOT/include/sgbm_b200.h $dst/include/
cp $ROOT/stereo_reconstruction_cv_b200/csrc/*.cu $ROOT/stereo_reconstruction_cv_b200/csrc/*.cuh $ROOT/stereo_reconstruction_cv_b200/csrc/Makefile $dst/pkg/csrc/
cd $dst/pkg/csrc
make -j8 -s NVCC="/usr/local/cuda/bin/nvcc $flags" "$@" ../libsgbm_b200.so
ls -la $dst/pkg/libsgbm_b200.so
