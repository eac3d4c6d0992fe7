#!/bin/bash
# Bring-up aid: compile-time variants of the stream kernel as separate .so files under video-graph-ssl_b200/build/variants/
# usage: tools/build_variants.sh name1 "-DFLAG=.. -DFLAG2" name2 "..." ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
P=$ROOT/video-graph-ssl_b200
mkdir -p $P/build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --use_fast_math -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
       -c $P/csrc/infonce_tcx.cu -o $P/build/variants/tcx_$name.o
  objs=$(ls $P/build/*.o | grep -v infonce_tcx.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $P/build/variants/lib_$name.so $objs $P/build/variants/tcx_$name.o -lcudart
  echo built $name
done
