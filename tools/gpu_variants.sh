#!/bin/bash
# A/B of library variants (tools/build_variant.sh) on C1..C3 (small scenes) and the bench workload.
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset LG_LIB_PATH; else export LG_LIB_PATH=$PWD/light_garden_b200/_lib/variants/lib_$v.so; fi
  timeout 300 python tools/bench_configs.py --repeat 2 --only C1,C2,C3 > gpurun_out/variant_${v}_configs.log 2>&1
  grep '^{' gpurun_out/variant_${v}_configs.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$v', d['config'], 'trace', round(d['trace_ms'],2), 'acc', round(d['accumulate_ms'],1))"
done
