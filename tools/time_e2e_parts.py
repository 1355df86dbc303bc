#!/usr/bin/env python
"""Times the host-facing pieces of one end-to-end frame of the bench workload (C5 / 8): lg_scene_set, lg_lights_set,
lg_image_read(RGBA16F) -- the part of bench.py's `e2e` that is not kernels.  Needs a GPU."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_garden_b200 import abi, scenes  # noqa: E402
from light_garden_b200.scene import flatten_objects, lights_to_array, trace_params  # noqa: E402
from light_garden_b200.tracer import Context, pinned_array  # noqa: E402


def main():
    spec = scenes.c5_large(n_lights=1, rays_per_light=1_000_000)
    ctx = Context(0, abi.LG_PRECISION_F32)
    objs, n_obj, nodes, n_nodes = flatten_objects(spec.objects)
    prm = trace_params(spec.max_bounce, spec.cutoff_color, spec.canvas_bounds)
    larr = lights_to_array(spec.lights)
    ctx.call("lg_image_configure", spec.width, spec.height)
    frame = pinned_array((spec.height, spec.width, 4), np.float16)
    st = abi.LgTraceStats()

    def timed(name, f, reps=10):
        f()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        print(f"{name:28s} {1e3 * (time.perf_counter() - t0) / reps:8.3f} ms")

    timed("lg_scene_set", lambda: ctx.call("lg_scene_set", C.cast(objs, C.c_void_p), n_obj, C.cast(nodes, C.c_void_p),
                                           n_nodes, C.byref(prm)))
    timed("lg_lights_set", lambda: ctx.call("lg_lights_set", C.cast(larr, C.c_void_p), len(spec.lights)))
    timed("lg_image_clear", lambda: ctx.call("lg_image_clear", C.c_float(1.0)))

    def render():
        ctx.call("lg_scene_set", C.cast(objs, C.c_void_p), n_obj, C.cast(nodes, C.c_void_p), n_nodes, C.byref(prm))
        ctx.call("lg_render", C.byref(st))
    timed("lg_scene_set + lg_render 1M", render, 5)
    timed("lg_render 1M", lambda: ctx.call("lg_render", C.byref(st)), 5)
    timed("lg_image_read RGBA16F", lambda: ctx.call("lg_image_read", abi.LG_RGBA16F, abi.array_ptr(frame), 0))


if __name__ == "__main__":
    main()
