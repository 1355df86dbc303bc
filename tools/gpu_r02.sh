#!/bin/bash
# Round-2 GPU round trips.  Usage (under gpurun): bash tools/gpu_r02.sh <step>...
#   tests     GPU parity suite + smoke
#   abcfg     BASELINE configs C1..C4, in-tree library against light_garden_b200/_lib/variants/lib_<name>.so (AB_LIBS)
#   abbench   short bench.py A/B (16 M rays)
#   bench     full bench line (+ f64)
#   profraster / proftrace   ncu --set full of the tile raster / the all-objects trace kernel
#   launches  ncu launch list of the short bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.log 2>&1
AB_LIBS="${AB_LIBS:-new light_garden_b200/_lib/variants/lib_r01.so}"
for w in "$@"; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -q --durations=10 --timeout=300 --timeout-method=thread \
        > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
      timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
      echo "smoke rc=$?" >> gpurun_out/smoke.log
      ;;
    abcfg)
      for v in $AB_LIBS; do
        if [ "$v" = new ]; then unset LG_LIB_PATH; else export LG_LIB_PATH=$PWD/$v; fi
        n=$(basename "$v" .so)
        timeout 600 python tools/bench_configs.py --repeat 3 > gpurun_out/cfg_$n.jsonl 2> gpurun_out/cfg_$n.err
      done
      unset LG_LIB_PATH
      ;;
    abbench)
      bash tools/ab.sh $AB_LIBS > gpurun_out/ab.txt 2>&1
      ;;
    abenv)
      # A/B over environment knobs of the in-tree library: AB_ENVS="NAME=V NAME2=V2;NAME=W;..." ("-" = defaults)
      : > gpurun_out/abenv.txt
      IFS=';' read -ra ENVS <<< "${AB_ENVS:--}"
      i=0
      for e in "${ENVS[@]}"; do
        i=$((i+1))
        if [ "$e" = "-" ]; then e=""; fi
        env $e timeout 300 python bench.py --rays-per-gpu ${AB_RAYS:-16000000} --steps 3 --no-cpu-baseline ${AB_ARGS:-} \
          > gpurun_out/abenv_$i.log 2>&1
        python - "$i" "$e" >> gpurun_out/abenv.txt <<'PY'
import json, sys
for line in open(f"gpurun_out/abenv_{sys.argv[1]}.log"):
    if line.startswith("{"):
        d = json.loads(line)
        g = d.get("tile_map_enabled") or {}
        gp = g.get("phase_ms_per_step", {})
        print(f"[{sys.argv[2] or 'defaults'}] ms/step {d['ms_per_step']:.2f} e2e {d['e2e']['ms_per_step']:.2f} trace {d['phase_ms_per_step']['trace']:.2f} acc "
              f"{d['phase_ms_per_step']['accumulate']:.2f} | grid: ms/step {g.get('ms_per_step', 0):.2f} trace {gp.get('trace', 0):.2f} acc {gp.get('accumulate', 0):.2f}")
PY
      done
      ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
      ;;
    benchf64)
      timeout 600 python bench.py --precision f64 --rays-per-gpu 8000000 --no-cpu-baseline > gpurun_out/bench_f64.log 2>&1
      ;;
    cfglaunches)
      # per-kernel times of C1 / C2 (quarter size) under ncu, this tree and the round-1 tree (_r01/, if present)
      for tree in ${CFG_TREES:-. _r01}; do
        [ -d $tree/tools ] || continue
        n=$(echo $tree | tr -d './_'); n=${n:-new}
        (cd $tree && LG_ACCUM_MODE=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file $OLDPWD/gpurun_out/launches_cfg_$n.csv python tools/bench_configs.py --scale 0.25 --repeat 1 --only ${CFG_ONLY:-C1,C2} \
          > $OLDPWD/gpurun_out/cfg_under_ncu_$n.log 2>&1)
      done
      ;;
    launches)
      LG_ACCUM_MODE=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
        --log-file gpurun_out/launches.csv python bench.py --rays-per-gpu 4000000 --steps 2 --warmup 3 --no-cpu-baseline --no-extras \
        > gpurun_out/bench_under_ncu.log 2>&1
      ;;
    profraster)
      LG_ACCUM_MODE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_raster -s 1 -c 1 \
        -f -o gpurun_out/prof_tile_raster python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras \
        > gpurun_out/prof_tile_raster.log 2>&1
      ;;
    proftrace)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 \
        -f -o gpurun_out/prof_trace python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras \
        > gpurun_out/prof_trace.log 2>&1
      ;;
    profbins)
      LG_ACCUM_MODE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fill -s 1 -c 1 \
        -f -o gpurun_out/prof_tile_fill python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras \
        > gpurun_out/prof_tile_fill.log 2>&1
      ;;
    sanitize)
      timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 3 python -m pytest tests -m gpu -q -x \
        -k "not full_size and not bench_contract" > gpurun_out/sanitize_memcheck.log 2>&1
      echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
      timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize_racecheck.log 2>&1
      echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
      ;;
    ubench)
      (cd tools/ubench && for b in pipes coissue mix; do [ -x ./$b ] && { echo "== tools/ubench/$b"; ./$b; }; done) > gpurun_out/ubench.txt 2>&1
      ;;
  esac
done
tail -3 gpurun_out/pytest_gpu.log 2>/dev/null
tail -2 gpurun_out/smoke.log 2>/dev/null
cat gpurun_out/ab.txt 2>/dev/null
cat gpurun_out/abenv.txt 2>/dev/null
exit 0
