#!/bin/bash
# A/B of the grid walk's knobs (cells per object, ray slots per thread) on the bench workload.
mkdir -p gpurun_out; : > gpurun_out/grid_variants.txt
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --rays-per-gpu 16000000 --steps 2 --no-cpu-baseline > gpurun_out/gridvar_$name.log 2>&1
  python - "$name" >> gpurun_out/grid_variants.txt <<'PY'
import json, sys
for line in open(f"gpurun_out/gridvar_{sys.argv[1]}.log"):
    if line.startswith("{"):
        t = json.loads(line)["tile_map_enabled"]
        print(sys.argv[1], "grid trace ms", round(t["phase_ms_per_step"]["trace"], 3), "rays/s", round(t["value"] / 1e6, 1))
PY
}
run d1s1 LG_GRID_DENSITY=1 LG_GRID_SLOTS=1
run d1s2 LG_GRID_DENSITY=1 LG_GRID_SLOTS=2
run d2s1 LG_GRID_DENSITY=2 LG_GRID_SLOTS=1
run d4s1 LG_GRID_DENSITY=4 LG_GRID_SLOTS=1
run d05s1 LG_GRID_DENSITY=0.5 LG_GRID_SLOTS=1
run d2s2 LG_GRID_DENSITY=2 LG_GRID_SLOTS=2
cat gpurun_out/grid_variants.txt
