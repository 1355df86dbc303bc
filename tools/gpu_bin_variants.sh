#!/bin/bash
# A/B of the count / fill passes' launch shape on the bench workload (accumulate ms with the grid trace).
mkdir -p gpurun_out; : > gpurun_out/bin_variants.txt
run() {
  name=$1; shift
  env "$@" LG_ACCUM_MODE=2 timeout 300 python bench.py --rays-per-gpu 16000000 --steps 3 --no-cpu-baseline > gpurun_out/binvar_$name.log 2>&1
  python - "$name" >> gpurun_out/bin_variants.txt <<'PY'
import json, sys
for line in open(f"gpurun_out/binvar_{sys.argv[1]}.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1], "accumulate ms", round(d["phase_ms_per_step"]["accumulate"], 2), round(d["tile_map_enabled"]["phase_ms_per_step"]["accumulate"], 2))
PY
}
run c4t256 LG_BIN_CTAS=4 LG_BIN_THREADS=256
run c3t256 LG_BIN_CTAS=3 LG_BIN_THREADS=256
run c3t512 LG_BIN_CTAS=3 LG_BIN_THREADS=512
run c6t512 LG_BIN_CTAS=6 LG_BIN_THREADS=512
run c8t256 LG_BIN_CTAS=8 LG_BIN_THREADS=256
run c3t1024 LG_BIN_CTAS=3 LG_BIN_THREADS=1024
cat gpurun_out/bin_variants.txt
