#!/bin/bash
# Recompile only csrc/silog_ss.cu and relink libmde_b200.so from the objects of the last full build
# (iteration aid: the full build takes ~3 min because of the masked-loss template instantiations).
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
nvcc $FLAGS $EXTRA -c mono_depth_estimation_b200/csrc/silog_ss.cu -o build/mde_b200/silog_ss.o
nvcc -shared -o mono_depth_estimation_b200/libmde_b200.so build/mde_b200/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart=static
# instrumented twin (per-warp wait / compute cycles in the trace): tools/libmde_dbg.so, selected with MDE_B200_LIB
mkdir -p build/dbg
nvcc $FLAGS -DMDE_SS_TIMING -c mono_depth_estimation_b200/csrc/silog_ss.cu -o build/dbg/silog_ss.o
nvcc -shared -o tools/libmde_dbg.so $(ls build/mde_b200/*.o | grep -v silog_ss.o) build/dbg/silog_ss.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart=static
echo relinked
