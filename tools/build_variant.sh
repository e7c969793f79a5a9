#!/bin/bash
# Build an A/B variant of the library: only ONE source (default spa_qc_spec.cu, the specialised resident kernels; SRC=...
# for another) is recompiled with the given -D flags and linked with the objects of the last full build.
#   [SRC=spa_generic.cu] tools/build_variant.sh NAME -DLDPC_X [-DLDPC_Y ...]   ->  ldpc-simulator_b200/lib/libldpc_NAME.so
set -e
cd "$(dirname "$0")/../ldpc-simulator_b200"
name=$1; shift
src=${SRC:-spa_qc_spec.cu}; obj=${src%.cu}.o
mkdir -p build_var/$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
     -Xcudafe --diag_suppress=177 "$@" -c csrc/$src -o build_var/$name/$obj 2> build_var/$name/ptxas.log
objs=$(ls build/*.o | grep -v "/$obj")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libldpc_$name.so $objs build_var/$name/$obj -cudart static -ldl
echo "lib/libldpc_$name.so"
