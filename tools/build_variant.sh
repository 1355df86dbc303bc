#!/bin/bash
# Builds light_garden_b200/_lib/variants/lib_<name>.so with extra -D flags for the trace kernels (tuning experiments,
# selected at run time with LG_LIB_PATH).  Usage: tools/build_variant.sh <name> -DLG_TRACE_GROUP=8 ...
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=light_garden_b200/_lib/variants; mkdir -p $out/obj_$name
F="-ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall"
for s in lg_trace_f32 lg_trace_f32_dup lg_trace_f64 lg_trace_f64_large lg_trace_grid lg_capi; do
  /usr/local/cuda/bin/nvcc $F "$@" -c light_garden_b200/csrc/$s.cu -o $out/obj_$name/$s.o &
done
wait
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -shared -o $out/lib_$name.so $out/obj_$name/*.o -ldl
echo $out/lib_$name.so
