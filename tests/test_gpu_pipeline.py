"""lg_render's wave pipeline (trace of wave k + 1 on one stream while the line pass of wave k runs on another, no host
round trip in between) against the one-wave-after-the-other path and against the oracle: same fragments, same counters,
and the two repairs it has -- a wave that overflows its half of the segment buffer, a pair list that turns out too
small -- produce the same frame."""
import numpy as np
import pytest

from light_garden_b200 import abi, scenes
from util import have_cuda, small_specs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="no CUDA device")]

TILED = 2


def _frame(ctx, spec, overlap, waves=0, clear_alpha=1.0):
    from light_garden_b200.tracer import Renderer, Tracer
    ctx.call("lg_render_overlap_set", overlap, waves)
    t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
    r = Renderer(ctx, spec.width, spec.height)
    r.clear(clear_alpha)
    st = r.render(t)
    return st, r.read_rgba32f(), r, t


@pytest.mark.parametrize("precision", [abi.LG_PRECISION_F32, abi.LG_PRECISION_F64])
@pytest.mark.parametrize("name", ["C1", "C3", "C5-16"])
def test_pipelined_frame_equals_sequential_frame(oracle, name, precision):
    from light_garden_b200.tracer import Context
    spec = small_specs()[name]
    ctx = Context(0, precision)
    try:
        ctx.call("lg_accumulate_mode_set", TILED)
        st0, img0, _, _ = _frame(ctx, spec, 0)
        for waves in (2, 3, 7):
            st1, img1, r, t = _frame(ctx, spec, 2, waves)
            assert st1.trace_launches == waves
            for k in ("primary_rays", "ray_steps", "object_tests", "segments", "pixel_updates"):
                assert getattr(st1, k) == getattr(st0, k), k
            assert np.array_equal(img1[..., 3] > 1, img0[..., 3] > 1)
            # the same fragments, summed per tile in a different grouping (waves): fp32 association only
            assert (np.abs(img1 - img0) <= 2e-5 * np.maximum(1.0, np.abs(img0))).all()
        # and against the oracle's f64 sums of the device's own segments
        seg = t.trace_all(ordered=False, control_lines=False)
        exact = np.zeros((spec.height, spec.width, 4), dtype=np.float64)
        exact[..., 3] = 1.0
        assert oracle.accumulate_segments_f64(exact, seg) == st1.pixel_updates
        assert (np.abs(img1 - exact) <= 2e-5 * np.maximum(1.0, np.abs(exact))).all()
    finally:
        ctx.close()


def test_pipeline_repairs_overflowing_waves_and_short_pair_lists(oracle):
    """A fresh context learns 1 segment per ray and ~1 pair per segment from an empty scene on a tiny image; the cavity
    frame that follows (64 bounces, 480x270) overflows the halves of a small segment buffer and the pair list.  The
    frame must still be the sequential one."""
    from light_garden_b200.scene import PointLight, Rect
    from light_garden_b200.tracer import Context
    cav = small_specs()["C2"]
    empty = scenes.SceneSpec("empty", [], [PointLight((0.0, 0.0), 3000, (0.01, 0.01, 0.01, 0.02))], 5, 64, 64)
    ref_ctx = Context(0, abi.LG_PRECISION_F32)
    ctx = Context(0, abi.LG_PRECISION_F32)
    try:
        for c in (ref_ctx, ctx):
            c.call("lg_accumulate_mode_set", TILED)
        st0, img0, _, _ = _frame(ref_ctx, cav, 0)
        ctx.call("lg_segment_capacity_set", 16384)
        st_e, _, _, _ = _frame(ctx, empty, 2, 2)
        assert st_e.segments == 3000                       # one segment per ray: the estimate the next frame starts from
        st1, img1, _, _ = _frame(ctx, cav, 2, 2)
        assert st1.segments == st0.segments > 8 * 16384 // 2   # many times what one half holds
        for k in ("primary_rays", "ray_steps", "segments", "pixel_updates"):
            assert getattr(st1, k) == getattr(st0, k), k
        assert st1.trace_launches > 2                       # the overflowed waves were traced again in smaller ones
        assert (np.abs(img1 - img0) <= 2e-5 * np.maximum(1.0, np.abs(img0))).all()
        # once more: the estimates have adapted, nothing overflows, same frame
        st2, img2, _, _ = _frame(ctx, cav, 2, 2)
        assert st2.pixel_updates == st0.pixel_updates
        assert (np.abs(img2 - img0) <= 2e-5 * np.maximum(1.0, np.abs(img0))).all()
    finally:
        ctx.close()
        ref_ctx.close()


def test_pipeline_repairs_a_short_pair_list(oracle):
    """Default segment capacity (no wave overflows), but the pair list is sized from what the context has seen: an
    empty scene on a 64x64 image.  The cavity's waves need many times that; the passes of those waves run again."""
    from light_garden_b200.scene import PointLight
    from light_garden_b200.tracer import Context
    cav = small_specs()["C2"]
    empty = scenes.SceneSpec("empty", [], [PointLight((0.0, 0.0), 500, (0.01, 0.01, 0.01, 0.02))], 2, 64, 64)
    ref_ctx = Context(0, abi.LG_PRECISION_F32)
    ctx = Context(0, abi.LG_PRECISION_F32)
    try:
        for c in (ref_ctx, ctx):
            c.call("lg_accumulate_mode_set", TILED)
        st0, img0, _, _ = _frame(ref_ctx, cav, 0)
        _frame(ctx, empty, 2, 2)
        st1, img1, _, _ = _frame(ctx, cav, 2, 2)
        assert st1.trace_launches == 2 and st1.accumulate_launches > 2 * 5      # passes of a wave ran twice
        for k in ("primary_rays", "ray_steps", "segments", "pixel_updates"):
            assert getattr(st1, k) == getattr(st0, k), k
        assert (np.abs(img1 - img0) <= 2e-5 * np.maximum(1.0, np.abs(img0))).all()
        st2, img2, _, _ = _frame(ctx, cav, 2, 2)
        assert st2.accumulate_launches == 2 * 5 and st2.pixel_updates == st0.pixel_updates
    finally:
        ctx.close()
        ref_ctx.close()


def test_pipeline_is_skipped_where_it_cannot_run(oracle):
    """Direct resolve, tags, non-default blend states: lg_render falls back to the sequential waves, silently and with
    the same result."""
    from light_garden_b200.tracer import Context
    spec = small_specs()["C1"]
    ctx = Context(0, abi.LG_PRECISION_F32)
    try:
        ctx.call("lg_accumulate_mode_set", 1)
        st0, img0, _, _ = _frame(ctx, spec, 0)
        st1, img1, _, _ = _frame(ctx, spec, 2, 4)
        assert st1.trace_launches == st0.trace_launches == 1 and st1.pixel_updates == st0.pixel_updates
        assert (np.abs(img1 - img0) <= 1e-5 * np.maximum(1.0, np.abs(img0))).all()
    finally:
        ctx.close()


def test_one_context_through_many_workloads_equals_fresh_contexts(oracle):
    """A context carries estimates and buffers from one call to the next (segments per ray, pairs per segment, the pair
    list, the per-workload choice of resolve, the image).  Very different workloads in a row on ONE context -- a deep
    cavity, a 4096 x 4096 chord pattern, a many-object scene on a small frame, a large frame with few rays, with and
    without the wave pipeline, and the first one again -- must each give what a fresh context gives."""
    from light_garden_b200.scene import StringMod, StringModMode
    from light_garden_b200.tracer import Context, Renderer, Tracer
    specs = small_specs()
    k = 2.0 ** -8
    chords = StringMod(modulo=30000, num=11, mode=StringModMode.Mul, color=(k, k, k, k))
    big = scenes.c1_default(total_rays=3000, width=1920, height=1080)
    plan = [("trace", specs["C2"], 0), ("chords", chords, 0), ("trace", specs["C5-16"], 2), ("trace", big, 0),
            ("chords", chords, 0), ("trace", specs["C3"], 1), ("trace", specs["C2"], 2)]

    def run(ctx, kind, what, overlap):
        ctx.call("lg_render_overlap_set", overlap, 3)
        if kind == "chords":
            r = Renderer(ctx, 4096, 4096)
            r.clear()
            st = r.render_string_mod(what)
            img = r.read_rgba32f()
            return (st.segments, st.pixel_updates), img[::8, ::8].copy()
        t = what.apply(Tracer(what.canvas_bounds, ctx=ctx))
        r = Renderer(ctx, what.width, what.height)
        r.clear()
        st = r.render(t)
        return (st.primary_rays, st.ray_steps, st.segments, st.pixel_updates), r.read_rgba32f()

    shared = Context(0, abi.LG_PRECISION_F32)
    try:
        for step, (kind, what, overlap) in enumerate(plan):
            got_c, got_i = run(shared, kind, what, overlap)
            fresh = Context(0, abi.LG_PRECISION_F32)
            try:
                exp_c, exp_i = run(fresh, kind, what, 0)
            finally:
                fresh.close()
            assert got_c == exp_c, (step, kind, got_c, exp_c)
            # either resolve, waves or not: fp32 association only (the direct resolve rounds hot pixels per fragment)
            assert (np.abs(got_i - exp_i) <= 3e-4 * np.maximum(1.0, np.abs(exp_i))).all(), step
    finally:
        shared.close()


def test_contexts_come_and_go_without_leaking_device_memory(oracle):
    """60 contexts created, used (trace, both resolves, string mod, every read-back format, an exported frame) and
    destroyed: the device's free memory ends where it began (within the allocator's slack), and a context destroyed
    while its exported frame is still held by the importer does not take the process down."""
    import ctypes as C
    import os
    import torch
    from light_garden_b200.scene import StringMod, StringModMode
    from light_garden_b200.tracer import Context, Renderer, Tracer
    spec = small_specs()["C5-16"]
    k = 2.0 ** -8
    sm = StringMod(modulo=2000, num=7, mode=StringModMode.Mul, color=(k, k, k, k))

    def use(ctx, hold_fd):
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        r = Renderer(ctx, 640, 360)
        for mode in (1, 2):
            ctx.call("lg_accumulate_mode_set", mode)
            r.clear()
            r.render(t)
            r.render_string_mod(sm)
        r.read_rgba32f(), r.read_rgba16f(), r.make_screenshot(), r.read_surface_bgra8()
        fd, nbytes = r.export_fd()
        if not hold_fd:
            os.close(fd)
        return fd

    torch.cuda.synchronize()
    warm = Context(0, abi.LG_PRECISION_F32)
    os.close(use(warm, True))
    warm.close()
    free0 = torch.cuda.mem_get_info(0)[0]
    held = []
    for i in range(60):
        c = Context(0, abi.LG_PRECISION_F64 if i % 5 == 0 else abi.LG_PRECISION_F32)
        fd = use(c, hold_fd=i % 10 == 0)
        if i % 10 == 0:
            held.append(fd)
        c.close()
    for fd in held:
        os.close(fd)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 64 << 20, (free0, free1)


@pytest.mark.parametrize("seed", range(6))
def test_random_call_sequences_leave_a_usable_context(oracle, seed):
    """The ABI's calls in random order with random (valid and invalid) arguments on one context: errors may only be the
    documented ones (state, invalid, overflow, unsupported), nothing crashes, and a canonical frame rendered afterwards
    equals the frame of a fresh context."""
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.scene import StringMod, StringModMode
    from light_garden_b200.tracer import Context, Renderer, Tracer
    rng = np.random.default_rng(0xAB1 + seed)
    specs = list(small_specs().values())
    k = 2.0 ** -8
    ok_codes = (abi.LG_ERR_STATE, abi.LG_ERR_INVALID, abi.LG_ERR_OVERFLOW, abi.LG_ERR_UNSUPPORTED)
    ctx = Context(0, abi.LG_PRECISION_F32 if seed % 2 == 0 else abi.LG_PRECISION_F64)
    try:
        tracer, rend = None, None
        for step in range(40):
            op = int(rng.integers(0, 13))
            try:
                if op == 0:
                    spec = specs[int(rng.integers(0, len(specs)))]
                    tracer = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
                elif op == 1:
                    rend = Renderer(ctx, int(rng.choice([1, 33, 160, 641])), int(rng.choice([1, 90, 257])))
                elif op == 2 and rend:
                    rend.clear(float(rng.choice([0.0, 1.0])))
                elif op == 3 and rend and tracer:
                    rend.render(tracer)
                elif op == 4 and tracer:
                    tracer.trace_all(ordered=bool(rng.integers(0, 2)), control_lines=False)
                elif op == 5 and rend:
                    rend.render_traced()
                elif op == 6 and rend:
                    rend.render_string_mod(StringMod(modulo=int(rng.choice([0, 1, 2, 500])), num=int(rng.integers(0, 9)),
                                                     mode=int(rng.integers(0, 4)), color=(k, k, k, k)))
                elif op == 7 and rend:
                    [rend.read_rgba32f, rend.read_rgba16f, rend.make_screenshot, rend.read_surface_bgra8][int(rng.integers(0, 4))]()
                elif op == 8 and tracer:
                    tracer.enable_tile_map(bool(rng.integers(0, 2)))
                elif op == 9:
                    ctx.call("lg_accumulate_mode_set", int(rng.integers(-1, 4)))
                elif op == 10 and tracer:
                    w = int(rng.integers(1, 5))
                    tracer.set_shard(int(rng.integers(0, w + 1)), w)           # rank == world: invalid
                elif op == 11:
                    ctx.call("lg_segment_capacity_set", int(rng.choice([0, 64, 5000, 1 << 20, 64 << 20])))
                elif op == 12:
                    ctx.call("lg_render_overlap_set", int(rng.integers(-1, 4)), int(rng.integers(0, 9)))
            except LightGardenError as e:
                assert e.code in ok_codes, (step, op, e.code, e.message)
        # back to a known state, then the canonical frame
        ctx.call("lg_segment_capacity_set", 64 << 20)
        ctx.call("lg_accumulate_mode_set", 2)
        ctx.call("lg_render_overlap_set", 0, 0)
        spec = small_specs()["C3"]
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        t.set_shard(0, 1)
        t.enable_tile_map(False)
        r = Renderer(ctx, spec.width, spec.height)
        r.clear()
        st = r.render(t)
        got = r.read_rgba32f()
        fresh = Context(0, ctx.precision if hasattr(ctx, "precision") else (abi.LG_PRECISION_F32 if seed % 2 == 0 else abi.LG_PRECISION_F64))
        try:
            fresh.call("lg_accumulate_mode_set", 2)
            t2 = spec.apply(Tracer(spec.canvas_bounds, ctx=fresh))
            r2 = Renderer(fresh, spec.width, spec.height)
            r2.clear()
            st2 = r2.render(t2)
            exp = r2.read_rgba32f()
        finally:
            fresh.close()
        assert (st.primary_rays, st.ray_steps, st.segments, st.pixel_updates) == (st2.primary_rays, st2.ray_steps, st2.segments, st2.pixel_updates)
        assert (np.abs(got - exp) <= 2e-5 * np.maximum(1.0, np.abs(exp))).all()
    finally:
        ctx.close()
