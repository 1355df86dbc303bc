#!/usr/bin/env python
"""bench.py -- the headline benchmark of the Light Garden hot path on B200.

Workload (BASELINE.json configs[4], the one the north-star target is quoted on): the synthetic 4096-object scene
(2048 circles, 1024 straight mirrors, 1024 rects on a jittered 64x64 lattice) with ALL EIGHT point lights of C5 at
every N; every light emits N x 4 M rays and rank r of N traces the rays r, r + N, r + 2N, ... of every light
(lg_shard_set), i.e. 32 M primary rays per GPU from the same eight light positions whatever N is: weak scaling with
identical per-GPU work (N = 8 is exactly C5 = 256 M rays).  max_bounce 5, all-objects ray x object tests,
3840x2160 accumulation, image reduce over NVLink peer memory.

One step = one frame of the reference (framework.rs:200-234): clear, trace every ray of the shard, accumulate every
segment, sum the partial images onto rank 0.
  value : primary rays/s, whole job, scene + lights already resident on the device
  e2e   : the same through the C ABI with host buffers: lg_scene_set + lg_lights_set (H2D) ... lg_image_read of the
          Rgba16Float frame (D2H) inside the timed region
  roofline : the trace kernel.  frac = EXECUTED FP32 work / the FP32 FMA peak measured in this run: the broad phase
          decides a ray x object test with 3 fused multiply-adds (6 flop).  `contract` keeps SURVEY.md 8d's algorithmic
          figure (16.5 flop per test for this mix), which is NOT a fraction of anything (the kernel does not execute it).
          `traffic` comes from the committed ncu capture named in `traffic_source` (profiles/r02_traffic.json)
  roofline_accumulate : the tile raster.  frac = fragments/s / the shared-memory read-modify-write ceiling measured in
          this run (lg_measure_tile_rmw_peak); DRAM traffic against the algorithmic bytes from the same ncu file
  configs : BASELINE configs C1..C4 at full size (N = 1 only), precision_f64 : the reference-width mode (N = 1 only)
  reduce_check (N > 1) : untimed, after the timed regions -- the reduced frame on rank 0 equals the rank-ordered sum of
          the ranks' partial frames bit for bit, and the ranks' fragment counts add up to a one-rank render's
  cpu_baseline : the oracle (restated reference, f64, chunks of 100 rays over all host cores) on a bounded sample, in
          the reference's default configuration (its TileMap culling enabled, tile_map.rs:61); the all-objects loop's
          rate -- the loop the GPU headline runs -- is reported beside it.  Host threads = len(os.sched_getaffinity(0)),
          not OMP_NUM_THREADS (torchrun exports 1).

`--impl reference` times that CPU restatement alone (the real rayon binary cannot be built offline: no Rust
toolchain, collision2d not vendored).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_TEST = {"circle": 14.0, "segment": 14.0, "rect": 24.0}  # SURVEY.md 8d contract figures
EXECUTED_FLOP_PER_TEST = 6.0   # broad phase: 3 fused multiply-adds per ray x object (lg_trace.cuh: Broad<T>::test4)
RAYS_PER_GPU = 32_000_000
N_LIGHTS = 8                   # C5's eight point lights, at every N
WIDTH, HEIGHT = 3840, 2160


def host_cores():
    """Threads the CPU legs use: the cores this process may run on, whatever OMP_NUM_THREADS says."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bench_spec(scenes, world, rays_per_gpu):
    """The C5 scene with all eight lights; N x (rays_per_gpu / 8) rays per light."""
    return scenes.c5_large(n_lights=N_LIGHTS, rays_per_light=rays_per_gpu * world // N_LIGHTS)


def algorithmic_flops_per_test(objects):
    kinds = {"Circle": "circle", "StraightMirror": "segment", "Rect": "rect"}
    tot = sum(FLOPS_PER_TEST[kinds[o.kind]] for o in objects)
    return tot / len(objects)


class ClockSampler:
    """SM clock + throttle reasons DURING the timed regions (B200_PROFILING.md recipe).  NVML is polled from a thread of
    this process every 20 ms (the timed calls are ctypes calls that release the GIL), which gives tens of samples for a
    half-second region; if NVML cannot be loaded the recipe's `nvidia-smi -lms 200` child process is used instead."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits (nvml.h)
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.proc = self.th = self.nv = self.h = None
        self.lines, self.sm, self.power, self.reasons = [], [], [], set()
        self.sm_max = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode())
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                h = nv.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv, self.h = nv, h
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for name, b in self.BITS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.th.join(timeout=2)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                    "sm_min_mhz": sm[0] if sm else None, "power_w_max": max(self.power) if self.power else None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 20 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 200"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_reference_sample(oracle, osc, spec, abi, seconds, tile_map=True, with_image=True, repeats=1, threads=0):
    """Times the restated reference on a bounded sample of the workload: every stride-th primary ray, chunks of 100
    rays over all host cores, f64 -- with its TileMap on (the app's default, tile_map.rs:61) or off (the all-objects
    loop the GPU headline runs).  Returns (rays/s, segments/s, stride, rays per repeat, seconds per repeat)."""
    osc.enable_tile_map(tile_map)
    stride = max(1, spec.total_rays() // 4000)
    t0 = time.perf_counter()
    probe = osc.trace_all(spec.lights, abi.LG_PRECISION_F64, stride=stride, store=False, threads=threads)
    rps = probe.primary_rays / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(spec.total_rays(), max(20_000, rps * seconds)))
    stride = max(1, spec.total_rays() // sample)
    img = oracle.new_image(WIDTH, HEIGHT) if with_image else None
    total, rays, segs = 0.0, 0, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = osc.trace_all(spec.lights, abi.LG_PRECISION_F64, stride=stride, store=with_image, threads=threads)
        if with_image:
            img[...] = 0
            img[..., 3] = 1
            oracle.accumulate_segments(img, res.seg, threads=threads)
        total += time.perf_counter() - t0
        rays += res.primary_rays
        segs += res.segments_emitted
    return rays / total, segs / total, stride, rays // repeats, total / repeats


def run_reference(args):
    """The reference arm: the CPU restatement of Tracer::trace_all + the line pass on the host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lg_oracle as oracle
    from light_garden_b200 import abi, scenes
    oracle.build()
    n = args.gpus
    spec = bench_spec(scenes, n, args.rays_per_gpu)
    osc = oracle.OracleScene.from_spec(spec)
    cores = host_cores()
    t0 = time.perf_counter()
    entries = osc.enable_tile_map(True)      # TileMap::new(w, h, 100, 100, 8), tracer.rs:27; built once per scene
    build_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        cpu_reference_sample(oracle, osc, spec, abi, args.ref_seconds, tile_map=True, threads=cores)
    value, segs_per_s, stride, rays_step, sec_step = cpu_reference_sample(oracle, osc, spec, abi, args.ref_seconds,
                                                                          tile_map=True, repeats=args.steps, threads=cores)
    brute, _, bstride, brays, bsec = cpu_reference_sample(oracle, osc, spec, abi, min(4.0, args.ref_seconds),
                                                          tile_map=False, with_image=False, threads=cores)
    out = {
        "impl": "reference", "metric": "rays_per_sec_traced_and_accumulated", "value": value, "unit": "rays/s",
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, args.rays_per_gpu),
        "segments_per_s": segs_per_s,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "sample": f"every {stride}-th primary ray of the workload per step ({rays_step} rays/step), f64 "
                                   "restated oracle, rayon-style chunks of 100, TileMap 100x100x8 enabled (the "
                                   f"reference's default, tile_map.rs:61; {entries} list entries built in {build_s:.1f} s, "
                                   "not timed) + host accumulate",
                         "all_objects_loop": {"value": brute, "unit": "rays/s",
                                              "sample": f"every {bstride}-th ray ({brays} rays, {bsec:.1f} s), TileMap off, "
                                                        "trace only"}},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference (oracle/): the real rayon binary cannot be built offline (no Rust, collision2d unvendored)",
    }
    print(json.dumps(out), flush=True)
    return 0


def workload_config(n, rays_per_gpu=RAYS_PER_GPU):
    return {"workload": "C5: synthetic 4096-object scene (2048 circles, 1024 mirrors, 1024 rects), all 8 point lights x "
                        f"{rays_per_gpu * n // N_LIGHTS} rays = {rays_per_gpu} primary rays per GPU x {n} GPU(s) (rank r "
                        "traces rays r, r + N, ... of every light: the same per-GPU work at every N), max_bounce 5, "
                        f"all-objects loop, {WIDTH}x{HEIGHT} RGBA accumulation, image reduce over NVLink peer memory "
                        "(NCCL for the handle exchange)",
            "objects": 4096, "lights": N_LIGHTS, "rays_per_gpu": rays_per_gpu, "max_bounce": 5, "width": WIDTH,
            "height": HEIGHT, "parallelism": f"ray-shard x{n}",
            "l2": "inputs larger than L2: GB-scale segment stream and a 132.7 MB image per step"}


def load_traffic():
    """DRAM traffic of the two hot kernels from the committed ncu captures (profiles/r02_traffic.json, written by
    tools/ncu_traffic.py from the `ncu --set full` reports of this code); None when the file is absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None


def time_configs(abi, scenes, Context, Renderer, Tracer, precision, repeat=3):
    """BASELINE configs C1..C4 at their full sizes on this GPU (C5 is the headline): best of `repeat` frames after the
    accumulate auto mode has settled (it samples each resolve twice)."""
    out = {}
    ctx = Context(0, precision)
    try:
        ctx.call("lg_segment_capacity_set", 512 << 20)
        specs = [("C1", scenes.c1_default(total_rays=1_000_000, width=1920, height=1080)),
                 ("C2", scenes.c2_cavity(total_rays=4_000_000, max_bounce=64, width=1920, height=1080)),
                 ("C3", scenes.c3_refraction(total_rays=16_000_000, width=1920, height=1080))]
        for name, spec in specs:
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
            r = Renderer(ctx, spec.width, spec.height)
            best = None
            for it in range(5 + repeat):
                r.clear()
                t0 = time.perf_counter()
                st = r.render(t)
                dt = time.perf_counter() - t0
                if it >= 5 and (best is None or dt < best[0]):
                    best = (dt, st.as_dict())
            dt, st = best
            out[name] = {"scene": spec.name, "objects": len(spec.objects), "primary_rays": st["primary_rays"],
                         "max_bounce": spec.max_bounce, "image": [spec.width, spec.height], "ms": dt * 1e3,
                         "trace_ms": st["trace_ms"], "accumulate_ms": st["accumulate_ms"],
                         "rays_per_s": st["primary_rays"] / dt, "segments": st["segments"],
                         "segments_per_s_in_kernel": st["segments"] / max(1e-9, st["accumulate_ms"] * 1e-3),
                         "ray_object_tests_per_s_in_kernel": st["object_tests"] / max(1e-9, st["trace_ms"] * 1e-3),
                         "pixel_updates": st["pixel_updates"],
                         "pixel_updates_per_s_in_kernel": st["pixel_updates"] / max(1e-9, st["accumulate_ms"] * 1e-3),
                         "resolve": "tile bins" if st["accumulate_launches"] > st["trace_launches"] else "direct"}
        for key, num in (("C4", 2), ("C4_num7919", 7919)):
            sm = scenes.c4_string_mod(modulo=10_000_000, num=num)
            r = Renderer(ctx, 4096, 4096)
            best = None
            for it in range(5 + repeat):
                r.clear()
                t0 = time.perf_counter()
                st = r.render_string_mod(sm)
                dt = time.perf_counter() - t0
                if it >= 5 and (best is None or dt < best[0]):
                    best = (dt, st.as_dict())
            dt, st = best
            out[key] = {"scene": f"string mod modulo={sm.modulo} num={num} Mul", "image": [4096, 4096], "chords": st["segments"],
                        "ms": dt * 1e3, "accumulate_ms": st["accumulate_ms"], "chords_per_s": st["segments"] / dt,
                        "pixel_updates": st["pixel_updates"],
                        "pixel_updates_per_s_in_kernel": st["pixel_updates"] / max(1e-9, st["accumulate_ms"] * 1e-3)}
    finally:
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays-per-gpu", type=int, default=RAYS_PER_GPU)
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--ref-seconds", type=float, default=8.0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline regions only: no grid pass, no C1..C4, no f64 pass (profiling and A/B runs)")
    args = ap.parse_args()
    if args.rays_per_gpu % N_LIGHTS:
        ap.error(f"--rays-per-gpu must be a multiple of {N_LIGHTS}")
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    import numpy as np
    import torch
    import torch.distributed as dist

    from light_garden_b200 import abi, scenes
    from light_garden_b200._lib import check, load
    from light_garden_b200.scene import flatten_objects, lights_to_array, trace_params
    from light_garden_b200.tracer import Context, Renderer, Tracer, pinned_array

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rays_per_gpu = args.rays_per_gpu
    prec = abi.LG_PRECISION_F64 if args.precision == "f64" else abi.LG_PRECISION_F32
    spec = bench_spec(scenes, world, rays_per_gpu)
    ctx = Context(local, prec)
    lib = load()
    tracer = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
    rend = Renderer(ctx, WIDTH, HEIGHT)
    ctx.call("lg_segment_capacity_set", 512 << 20)          # 16 GB of the 180 GB: one wave per step
    tracer.sync_scene()
    tracer.set_shard(rank, world)
    ctx.call("lg_tags_enable", 0)

    # communicator for the image reduce: the unique id travels over torch.distributed
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            check(None, lib.lg_comm_unique_id(raw))
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        ctx.call("lg_comm_init_rank", raw, rank, world)

    sh = C.c_uint64()
    ctx.call("lg_stream_handle", C.byref(sh))
    stream = torch.cuda.ExternalStream(sh.value, device=torch.device("cuda", local))

    objs, n_obj, nodes, n_nodes = flatten_objects(spec.objects)
    prm = trace_params(spec.max_bounce, spec.cutoff_color, spec.canvas_bounds)
    larr = lights_to_array(spec.lights)
    h2d = C.sizeof(objs) + C.sizeof(nodes) + C.sizeof(larr) + C.sizeof(prm)
    frame16 = pinned_array((HEIGHT, WIDTH, 4), np.float16)   # page-locked: the D2H runs at PCIe speed
    d2h = frame16.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e):
        st = abi.LgTraceStats()
        red = C.c_float()
        if e2e:  # the frame as a host application drives it: scene in, Rgba16Float frame out
            ctx.call("lg_scene_set", C.cast(objs, C.c_void_p), n_obj, C.cast(nodes, C.c_void_p), n_nodes, C.byref(prm))
            ctx.call("lg_lights_set", C.cast(larr, C.c_void_p), len(spec.lights))
        ctx.call("lg_image_clear", C.c_float(1.0 if rank == 0 else 0.0))
        ctx.call("lg_render", C.byref(st))
        ctx.call("lg_image_reduce", 0, C.byref(red))
        if e2e and rank == 0:
            ctx.call("lg_image_read", abi.LG_RGBA16F, abi.array_ptr(frame16), 0)
        return st, red.value

    SUMS = ("ray_steps", "segments", "pixel_updates", "trace_launches", "accumulate_launches")
    TIMES = ("trace_ms", "accumulate_ms", "reduce_ms")

    def timed(steps, e2e):
        """EXACTLY `steps` steps between barrier+sync, CUDA events on the library's stream, max over ranks.  Counters
        are summed over ranks; the phase times come back as max, min and per rank."""
        mine = {k: 0.0 for k in SUMS + TIMES}
        l0 = ctx.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                st, red = step(e2e)
                for k in SUMS + TIMES[:2]:
                    mine[k] += getattr(st, k)
                mine["reduce_ms"] += red
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        mine["step_ms"] = ms
        keys = list(SUMS + TIMES) + ["step_ms"]
        per_rank = [[float(mine[k]) for k in keys]]
        if world > 1:
            t = torch.tensor(per_rank[0], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank = [x.cpu().tolist() for x in allt]
        agg = {}
        for i, k in enumerate(keys):
            col = [r[i] for r in per_rank]
            if k in SUMS:
                agg[k] = sum(col)
            else:
                agg[k] = max(col)
                agg[k + "_min"] = min(col)
                agg[k + "_per_rank"] = [c / steps for c in col]
        agg["launches"] = ctx.launch_count() - l0
        return agg["step_ms"], agg

    def phases(a, steps):
        p = {"trace": a["trace_ms"] / steps, "accumulate": a["accumulate_ms"] / steps, "image_reduce": a["reduce_ms"] / steps}
        if world > 1:   # max over ranks above; the skew between ranks is what the scaling loss is made of
            p["trace_min_over_ranks"] = a["trace_ms_min"] / steps
            p["accumulate_min_over_ranks"] = a["accumulate_ms_min"] / steps
            p["step_min_over_ranks"] = a["step_ms_min"] / steps
            p["trace_per_rank"] = a["trace_ms_per_rank"]
            p["accumulate_per_rank"] = a["accumulate_ms_per_rank"]
        return p

    # measured denominators (same device, same run)
    fma, rmw = C.c_double(), C.c_double()
    ctx.call("lg_measure_fma_peak", prec, 3, C.byref(fma))
    ctx.call("lg_measure_tile_rmw_peak", 3, C.byref(rmw))
    red_coal, red_rand = C.c_double(), C.c_double()
    ctx.call("lg_measure_red_peak", WIDTH * HEIGHT, 0, 2, C.byref(red_coal))
    ctx.call("lg_measure_red_peak", WIDTH * HEIGHT, 1, 2, C.byref(red_rand))

    # untimed set-up, not warm-up: the accumulate auto mode (lg_accumulate_mode_set 0) samples each resolve twice
    # (the first call of a mode pays cudaMalloc for its buffers) before it settles on the cheaper one
    for _ in range(4):
        step(False)
    step(True)   # first use of the end-to-end-only pieces (Rgba16Float buffer, finalize kernel, pinned read-back)
    for _ in range(max(3, args.warmup)):
        step(False)
    sampler = None
    if rank == 0:
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
        except Exception:
            uuid = None
        sampler = ClockSampler(local, uuid)
        sampler.start()
    ms, agg = timed(args.steps, False)          # the two headline regions (value, e2e) are sampled together
    ms_e2e, agg_e2e = timed(args.steps, True)
    clocks = sampler.stop() if sampler else None

    # the same workload with Tracer::enable_tile_map (device grid): identical segments, fewer
    # exact tests.  Reported next to the headline, which stays the all-objects loop the north star names.
    grid_line = None
    if not args.no_extras:
        ctx.call("lg_tile_map_enable", 1)
        for _ in range(6):       # auto-mode samples of this workload + warm-up
            step(False)
        ms_grid, agg_grid = timed(args.steps, False)
        ms_grid_e2e, _ = timed(args.steps, True)
        ctx.call("lg_tile_map_enable", 0)
        grid_line = (ms_grid, agg_grid, ms_grid_e2e)

    # round 1's workload (ONE light x 32 M rays per GPU: a different light per rank, so not the same work at every N --
    # why it was replaced) once more, for continuity with BENCH_r01.json: N = 1 only
    r01_line = None
    if world == 1 and not args.no_extras:
        old = scenes.c5_large(n_lights=1, rays_per_light=rays_per_gpu)
        t_old = old.apply(Tracer(old.canvas_bounds, ctx=ctx))
        t_old.sync_scene()
        t_old.set_shard(0, 1)
        for _ in range(6):       # auto-mode samples of this workload + warm-up
            step(False)
        ms_old, agg_old = timed(args.steps, False)
        r01_line = {"value": rays_per_gpu * args.steps / (ms_old * 1e-3), "unit": "rays/s", "ms_per_step": ms_old / args.steps,
                    "phase_ms_per_step": phases(agg_old, args.steps), "segments_per_ray": agg_old["segments"] / (rays_per_gpu * args.steps),
                    "ray_object_tests_per_s_in_kernel": agg_old["ray_steps"] * len(old.objects) / (agg_old["trace_ms"] * 1e-3),
                    "note": "round 1's bench workload: one C5 light x 32 M rays (BENCH_r01.json: 146.2 ms/step, trace 111.2, "
                            "accumulate 35.1); the headline now traces all eight C5 lights x 4 M rays: more segments per ray "
                            "and more ray steps, identical at every N"}
        tracer.sync_scene(force=True)
        tracer.set_shard(rank, world)

    # multi-GPU correctness, untimed: the reduced frame is the rank-ordered sum of the partial frames, bit for bit,
    # and the ranks' fragment counts add up to a one-rank render's
    reduce_check = None
    if world > 1:
        small = bench_spec(scenes, world, 65536 * N_LIGHTS)          # 65 536 rays per light and rank
        t2 = small.apply(Tracer(small.canvas_bounds, ctx=ctx))
        t2.sync_scene()
        t2.set_shard(rank, world)
        st = abi.LgTraceStats()
        red = C.c_float()
        ctx.call("lg_image_clear", C.c_float(1.0 if rank == 0 else 0.0))
        ctx.call("lg_render", C.byref(st))
        part = torch.from_numpy(rend.read_rgba32f()).cuda()
        parts = [torch.zeros_like(part) for _ in range(world)] if rank == 0 else None
        dist.gather(part, parts, dst=0)
        cnt = torch.tensor([st.pixel_updates, st.segments, st.ray_steps], device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt)
        ctx.call("lg_image_reduce", 0, C.byref(red))
        if rank == 0:
            total = torch.from_numpy(rend.read_rgba32f()).cuda()
            acc = parts[0].clone()
            for q in parts[1:]:
                acc += q                                         # fp32, rank order: what the peer kernel does
            half = torch.from_numpy(rend.read_rgba16f().view(np.int16)).cuda()
            t2.set_shard(0, 1)                                   # the same lights and rays on one rank
            ctx.call("lg_image_clear", C.c_float(1.0))
            st1 = abi.LgTraceStats()
            ctx.call("lg_render", C.byref(st1))
            one = torch.from_numpy(rend.read_rgba32f()).cuda()
            rel = ((total - one).abs() / one.abs().clamp(min=1.0)).max().item()
            reduce_check = {
                "rays_per_rank": int(small.total_rays() // world), "ranks": world,
                "reduced_frame_equals_rank_ordered_sum_bit_for_bit": bool(torch.equal(total, acc)),
                "rgba16f_frame_equals_rounded_sum": bool(torch.equal(half, acc.to(torch.float16).view(torch.int16))),
                "sum_of_rank_pixel_updates": int(cnt[0].item()), "one_rank_pixel_updates": int(st1.pixel_updates),
                "sum_of_rank_segments": int(cnt[1].item()), "one_rank_segments": int(st1.segments),
                "counters_equal": bool(cnt[0].item() == st1.pixel_updates and cnt[1].item() == st1.segments
                                       and cnt[2].item() == st1.ray_steps),
                "max_rel_diff_vs_one_rank_frame": rel,
                "note": "fp32 sums in a different order (partial frames + reduce vs one frame): stated bound 1e-4 of the pixel value"}
            reduce_check["ok"] = bool(reduce_check["reduced_frame_equals_rank_ordered_sum_bit_for_bit"] and
                                      reduce_check["rgba16f_frame_equals_rounded_sum"] and reduce_check["counters_equal"]
                                      and rel < 1e-4)
        tracer.sync_scene(force=True)
        tracer.set_shard(rank, world)
        barrier()

    total_rays = rays_per_gpu * world * args.steps
    value = total_rays / (ms * 1e-3)
    e2e_value = total_rays / (ms_e2e * 1e-3)
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        fpt = algorithmic_flops_per_test(spec.objects)
        tests = agg["ray_steps"] * len(spec.objects)
        # per-launch figures of the dominant kernel (trace), summed over ranks / launches
        tr_launches = max(1.0, agg["trace_launches"])
        tr_ms_per_launch = agg["trace_ms"] / (tr_launches / world)    # trace_ms is max over ranks of per-rank sums
        tests_per_launch = tests / tr_launches
        executed_tflops = tests_per_launch * EXECUTED_FLOP_PER_TEST / (tr_ms_per_launch * 1e-3) / 1e12
        contract_tflops = tests_per_launch * fpt / (tr_ms_per_launch * 1e-3) / 1e12
        acc_ms_per_step_rank = agg["accumulate_ms"] / args.steps
        frag_per_s = agg["pixel_updates"] / world / max(1e-9, agg["accumulate_ms"] * 1e-3)
        traffic = load_traffic() or {}
        tt, tr = traffic.get("trace"), traffic.get("tile_raster")
        rays_per_launch = rays_per_gpu * world * args.steps / tr_launches
        segs_per_step_rank = agg["segments"] / world / args.steps
        out = {
            "metric": "rays_per_sec_traced_and_accumulated", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic", "config": workload_config(world, rays_per_gpu),
            "segments_per_s": agg["segments"] / (ms * 1e-3),
            "ray_object_tests_per_s": tests / (ms * 1e-3),
            "ray_object_tests_per_s_per_gpu_in_kernel": tests / world / (agg["trace_ms"] * 1e-3),
            "pixel_updates_per_s_in_kernel": frag_per_s,
            "segments_per_ray": agg["segments"] / total_rays,
            "phase_ms_per_step": phases(agg, args.steps),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "phase_ms_per_step": phases(agg_e2e, args.steps)},
            "gpu_launches": int(agg["launches"]),
            "clocks": clocks,
            "roofline": {
                "bound": "fp32", "kernel": "lg::trace_kernel (all-objects sweep)", "achieved": executed_tflops,
                "peak": fma.value, "unit": "TFLOP/s", "frac": executed_tflops / fma.value if fma.value else None,
                "flop_per_test_executed": EXECUTED_FLOP_PER_TEST,
                "contract": {"flop_per_test": fpt, "tflops": contract_tflops,
                             "note": "SURVEY.md 8d's algorithmic figure x tests; not a fraction of the peak: the kernel "
                                     "decides ~99 % of the tests with the 3-FMA bounding-circle line test and runs the "
                                     "exact ORACLE.md test on the survivors only"},
                "traffic": (tt["dram_bytes_per_ray"] * rays_per_launch) if tt and args.precision == "f32" else None,
                "traffic_unit": "DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum of the ncu capture, "
                                "per ray of that capture, x rays of one launch here",
                "traffic_source": ({"file": tt["file"], "capture_rays": tt["rays"], "dram_bytes": tt["dram_bytes"],
                                    "algorithmic_bytes_per_ray": 32.0 * agg["segments"] / total_rays} if tt else None),
                "note": "achieved = 6 executed flop (3 FMA) per ray x object test x tests per launch / CUDA-event launch "
                        f"time; peak = FMA microbenchmark of this run ({'FP64' if args.precision == 'f64' else 'FP32'} pipe; "
                        "MEASURED_PEAKS.json has none).  The sweep also issues 2 funnel shifts per object pair for the "
                        "mask: by the pipe model of profiles/r01f_ubench.txt 3 FFMA2 + 2 SHF cost 8.76 issue cycles per "
                        "pair, of which the FMAs are 6 -- the kernel measures 8.7-8.9 (profiles/r02_trace_blocks.txt); "
                        "the contract's hbm/tensor bounds do not apply: the table lives in shared memory and the kernel "
                        "writes 32 B per segment"},
            "roofline_accumulate": {
                "bound": "shared-memory read-modify-write", "kernel": "lg::tile_raster_kernel (+ count / scan / fill passes)",
                "achieved": frag_per_s / 1e9, "peak": rmw.value, "unit": "G fragments/s",
                "frac": frag_per_s / 1e9 / rmw.value if rmw.value else None,
                "phase": "whole accumulate phase of a step (count + scan + fill + raster), per GPU",
                "dram": ({"file": tr["file"], "capture_segments": tr["segments"],
                          "dram_bytes_per_segment": tr["dram_bytes_per_segment"], "algorithmic_bytes_per_segment": 32.0,
                          "ratio": tr["dram_bytes_per_segment"] / 32.0,
                          "traffic_per_step": tr["dram_bytes_per_segment"] * segs_per_step_rank,
                          "hbm_frac": tr["dram_bytes_per_segment"] * segs_per_step_rank / max(1e-9, acc_ms_per_step_rank * 1e-3) / 1e9 / hbm_peak,
                          "hbm_peak_GBps": hbm_peak, "hbm_peak_source": hbm_src} if tr else None),
                "red_v4_peak_gred_per_s": {"coalesced": red_coal.value, "random": red_rand.value},
                "note": "peak = lg_measure_tile_rmw_peak of this run: LDS.128 + 4 FADD + STS.128 on a private tile per "
                        "warp with every lane active, the raster's launch shape.  The gap is lane fill (a (segment, tile) "
                        "pair covers ~17 of 32 lanes on this workload), the parked records (2.5 of ~11 shared-memory "
                        "wavefronts per pair) and the binning passes; ncu: profiles/r02_tile_raster_full.txt.  Fragments "
                        "never reach DRAM: they are summed in shared memory and leave as one red.v4 per touched pixel"},
        }
        if grid_line:
            ms_grid, agg_grid, ms_grid_e2e = grid_line
            out["tile_map_enabled"] = {
                "value": total_rays / (ms_grid * 1e-3), "unit": "rays/s", "ms_per_step": ms_grid / args.steps,
                "e2e": total_rays / (ms_grid_e2e * 1e-3), "phase_ms_per_step": phases(agg_grid, args.steps),
                "segments": agg_grid["segments"], "segments_all_objects_loop": agg["segments"],
                "note": "lg_tile_map_enable(1): nearest hit through the device-side uniform grid (the reference's "
                        "TileMap option, tracer.rs:395-411); bit-identical segments (tests/test_gpu_trace.py), not the "
                        "headline because the north star's metric counts the all-objects loop"}
        if reduce_check is not None:
            out["reduce_check"] = reduce_check
        if r01_line is not None:
            out["r01_workload"] = r01_line
    # N = 1 only, after the timed regions: the other BASELINE configs and the reference-width mode
    if rank == 0 and world == 1 and not args.no_extras:
        ctx.close()
        ctx = None
        try:
            out["configs"] = time_configs(abi, scenes, Context, Renderer, Tracer, prec)
        except Exception as e:
            out["configs"] = {"error": f"{type(e).__name__}: {e}"}
        if args.precision == "f32":
            try:
                n64 = min(rays_per_gpu, 8_000_000)
                c64 = Context(local, abi.LG_PRECISION_F64)
                s64 = bench_spec(scenes, 1, n64)
                t64 = s64.apply(Tracer(s64.canvas_bounds, ctx=c64))
                r64 = Renderer(c64, WIDTH, HEIGHT)
                c64.call("lg_segment_capacity_set", 128 << 20)
                best = None
                for it in range(6):
                    r64.clear()
                    t0 = time.perf_counter()
                    st = r64.render(t64)
                    dt = time.perf_counter() - t0
                    if it >= 4 and (best is None or dt < best[0]):
                        best = (dt, st.as_dict())
                f64peak = C.c_double()
                c64.call("lg_measure_fma_peak", abi.LG_PRECISION_F64, 2, C.byref(f64peak))
                c64.close()
                dt, st = best
                out["precision_f64"] = {
                    "value": n64 / dt, "unit": "rays/s", "rays": n64, "ms": dt * 1e3, "trace_ms": st["trace_ms"],
                    "accumulate_ms": st["accumulate_ms"],
                    "ray_object_tests_per_s_in_kernel": st["object_tests"] / max(1e-9, st["trace_ms"] * 1e-3),
                    "fp64_fma_peak_tflops": f64peak.value,
                    "frac_of_fp64_peak": st["object_tests"] * EXECUTED_FLOP_PER_TEST / max(1e-9, st["trace_ms"] * 1e-3) / 1e12 / f64peak.value
                    if f64peak.value else None,
                    "note": "LG_PRECISION_F64: the reference's own arithmetic width (collision2d Float = f64), same scene, "
                            "same lights, fewer rays; the f32 headline is the throughput mode the north star names"}
            except Exception as e:
                out["precision_f64"] = {"error": f"{type(e).__name__}: {e}"}
    # CPU baseline: rank 0, N = 1 only, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import lg_oracle as oracle
            oracle.build()
            cores = host_cores()
            osc = oracle.OracleScene.from_spec(spec)
            osc.enable_tile_map(True)            # the reference's default (tile_map.rs:61); the build is scene set-up
            v, _, stride, nrays, dt = cpu_reference_sample(oracle, osc, spec, abi, args.cpu_seconds, tile_map=True, threads=cores)
            b, _, bstride, brays, bdt = cpu_reference_sample(oracle, osc, spec, abi, min(4.0, args.cpu_seconds),
                                                             tile_map=False, with_image=False, threads=cores)
            del osc
            out["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                                   "sample": f"every {stride}-th primary ray ({nrays} rays, {dt:.1f} s): f64 restated oracle "
                                             "trace (chunks of 100 rays over all cores, TileMap 100x100x8 enabled as in the "
                                             "reference's default) + host accumulate",
                                   "all_objects_loop": {"value": b, "unit": "rays/s",
                                                        "sample": f"every {bstride}-th ray ({brays} rays, {bdt:.1f} s), TileMap "
                                                                  "off (the loop the GPU headline runs), trace only"}}
        except Exception as e:   # the GPU line must not be lost to a failure of the CPU leg; the error is reported
            out["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        ctx.call("lg_comm_destroy")
        dist.barrier()
        dist.destroy_process_group()
    if ctx is not None:
        ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
