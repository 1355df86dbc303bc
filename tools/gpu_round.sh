#!/bin/bash
# One GPU-box round trip: parity tests, smoke, bench, ncu launch list + full captures of the two hot kernels.
# Usage (under gpurun): bash tools/gpu_round.sh [tests|bench|prof]...   (default: all)
set -u
mkdir -p gpurun_out
what="${*:-tests bench prof}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.log 2>&1
for w in $what; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -q --durations=15 --timeout=240 --timeout-method=thread \
        > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
      timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
      ;;
    bench)
      timeout 600 python bench.py > gpurun_out/bench.log 2>&1
      timeout 300 python bench.py --precision f64 --rays-per-gpu 8000000 --no-cpu-baseline > gpurun_out/bench_f64.log 2>&1
      ;;
    configs)
      timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
      timeout 900 python tools/bench_configs.py --precision f64 > gpurun_out/configs_f64.jsonl 2>> gpurun_out/configs.err
      ;;
    sanitize)
      # the whole GPU suite under memcheck (found the nested-call miscompile of round 1), then racecheck on smoke
      timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 3 python -m pytest tests -m gpu -q -x \
        -k "not full_size" > gpurun_out/sanitize_memcheck.log 2>&1
      echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
      timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize_racecheck.log 2>&1
      echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
      ;;
    proftiles)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv \
        --log-file gpurun_out/launches_tiles.csv python tools/bench_configs.py --scale 0.25 --repeat 1 \
        > gpurun_out/configs_under_ncu.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_raster -s 2 -c 1 \
        -f -o gpurun_out/prof_tile_raster python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline \
        > gpurun_out/prof_tile_raster.log 2>&1
      ;;
    slots)
      for s in 1 2 4; do
        LG_TRACE_SLOTS=$s timeout 300 python bench.py --rays-per-gpu 8000000 --steps 2 --no-cpu-baseline > gpurun_out/bench_slots$s.log 2>&1
      done
      LG_TRACE_SLOTS=2 timeout 300 python bench.py --precision f64 --rays-per-gpu 4000000 --steps 2 --no-cpu-baseline > gpurun_out/bench_f64_slots2.log 2>&1
      ;;
    prof)
      # launch list: the accumulate resolve is pinned to the one the un-profiled bench settles on (tile bins) -- under
      # ncu every launch is serialised, which distorts the auto mode's own timing of its two candidates
      LG_ACCUM_MODE=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
        --log-file gpurun_out/launches.csv python bench.py --rays-per-gpu 4000000 --steps 2 --warmup 3 --no-cpu-baseline \
        > gpurun_out/bench_under_ncu.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 12 -c 1 \
        -f -o gpurun_out/prof_trace_grid python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline \
        > gpurun_out/prof_trace_grid.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 \
        -f -o gpurun_out/prof_trace python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline \
        > gpurun_out/prof_trace.log 2>&1
      LG_ACCUM_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_segments -s 1 -c 1 \
        -f -o gpurun_out/prof_accum python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline \
        > gpurun_out/prof_accum.log 2>&1
      LG_ACCUM_MODE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_raster -s 1 -c 1 \
        -f -o gpurun_out/prof_tile_raster python bench.py --rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline \
        > gpurun_out/prof_tile_raster.log 2>&1
      ;;
  esac
done
tail -3 gpurun_out/pytest_gpu.log 2>/dev/null
tail -c 600 gpurun_out/bench.log 2>/dev/null
exit 0
