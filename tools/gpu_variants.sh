#!/bin/bash
# A/B of library variants (tools/build_variant.sh) on the C5 bench workload and C2/C3; prints one line per variant.
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset LG_LIB_PATH; else export LG_LIB_PATH=$PWD/light_garden_b200/_lib/variants/lib_$v.so; fi
  timeout 300 python bench.py --rays-per-gpu 16000000 --steps 2 --no-cpu-baseline > gpurun_out/variant_$v.log 2>&1
  timeout 300 python tools/bench_configs.py --repeat 2 --only C2,C3 > gpurun_out/variant_${v}_configs.log 2>&1
done
