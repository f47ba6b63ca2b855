#!/bin/bash
# tools/build_variant.sh NAME [SRCROOT] [extra nvcc flags...]: a twin of libmde_b200.so whose silog_ss.cu is compiled from
# SRCROOT (default: this tree) with the extra flags -> tools/variants/NAME.so (selected with MDE_B200_LIB; A/B runs of
# several variants on ONE box: box-to-box spread is +-3 %).
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
SRC=${1:-.}; shift || true
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p build/variants tools/variants
nvcc $FLAGS "$@" -I include -c $SRC/mono_depth_estimation_b200/csrc/silog_ss.cu -o build/variants/$NAME.o
nvcc -shared -o tools/variants/$NAME.so $(ls build/mde_b200/*.o | grep -v silog_ss.o) build/variants/$NAME.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart=static
echo built tools/variants/$NAME.so
