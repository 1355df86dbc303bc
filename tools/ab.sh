#!/bin/bash
# A/B on one GPU box: for each library given (path, or "new" for the in-tree build) run the short bench and print
# trace / accumulate ms per step.  Usage: tools/ab.sh new light_garden_b200/_lib/variants/lib_base.so ...
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = new ]; then unset LG_LIB_PATH; else export LG_LIB_PATH=$PWD/$v; fi
  n=$(basename "$v" .so)
  timeout 300 python bench.py --rays-per-gpu ${AB_RAYS:-16000000} --steps 3 --no-cpu-baseline > gpurun_out/ab_$n.log 2>&1
  python - "$n" <<'PY'
import json, sys
for line in open(f"gpurun_out/ab_{sys.argv[1]}.log"):
    if line.startswith("{"):
        d = json.loads(line)
        g = d["tile_map_enabled"]
        print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "trace", round(d["phase_ms_per_step"]["trace"], 2), "acc",
              round(d["phase_ms_per_step"]["accumulate"], 2), "| grid: trace", round(g["phase_ms_per_step"]["trace"], 2), "acc",
              round(g["phase_ms_per_step"]["accumulate"], 2), "| segments", d["segments_per_ray"])
PY
done
