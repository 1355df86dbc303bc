#!/bin/bash
# A/B of the accumulate resolves on the bench workload: LG_ACCUM_MODE 1 (direct), 2 (tile bins), 3 (hybrid) and
# hybrid variants.  One line per run in gpurun_out/accum_modes.txt.
mkdir -p gpurun_out; : > gpurun_out/accum_modes.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --rays-per-gpu 16000000 --steps 2 --no-cpu-baseline > gpurun_out/accum_$name.log 2>&1
  python - "$name" >> gpurun_out/accum_modes.txt <<'PY'
import json, sys
name = sys.argv[1]
for line in open(f"gpurun_out/accum_{name}.log"):
    if line.startswith("{"):
        d = json.loads(line)
        t = d["tile_map_enabled"]
        print(name, "all-objects", d["phase_ms_per_step"], "| grid", t["phase_ms_per_step"], "| grid rays/s", round(t["value"] / 1e6, 1))
PY
}
run direct LG_ACCUM_MODE=1
run tiled LG_ACCUM_MODE=2
run hybrid4 LG_ACCUM_MODE=3 LG_HYBRID_CTAS=4
run hybrid2 LG_ACCUM_MODE=3 LG_HYBRID_CTAS=2
run hybrid6 LG_ACCUM_MODE=3 LG_HYBRID_CTAS=6
run hybrid4y LG_ACCUM_MODE=3 LG_HYBRID_CTAS=4 LG_HYBRID_AXIS=2
cat gpurun_out/accum_modes.txt
