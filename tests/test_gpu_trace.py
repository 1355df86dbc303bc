"""Parity of the CUDA trace path (K1 emit + K2 trace) with the oracle, through the C ABI.

Bar (BASELINE.json north_star): hit-object sequences and segment counts bit-exact; here the device and
the oracle implement the same ORACLE.md formulas with IEEE-exact operations only, so for identical
primary rays EVERYTHING (tags, endpoints, colours) is compared bit for bit, in both precisions.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from light_garden_b200 import abi, scenes
from light_garden_b200.scene import (AND_NOT, Circle, CubicBezier, Logic, Material, Object, PointLight, Rect,
                                     SpotLight, lights_to_array)
from util import assert_same_segments, have_cuda, primary_rays, small_specs, ulp_diff64

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="no CUDA device")]

SPECS = small_specs()


@pytest.fixture(scope="module")
def ctx64():
    from light_garden_b200.tracer import Context
    c = Context(0, abi.LG_PRECISION_F64)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx32():
    from light_garden_b200.tracer import Context
    c = Context(0, abi.LG_PRECISION_F32)
    yield c
    c.close()


def make_tracer(spec, ctx):
    from light_garden_b200.tracer import Tracer
    return spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))


@pytest.mark.parametrize("name", list(SPECS))
def test_trace_f64_bit_exact(oracle, ctx64, name):
    spec = SPECS[name]
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    exp = osc.trace_rays(rays, abi.LG_PRECISION_F64)
    t = make_tracer(spec, ctx64)
    got = t.trace(rays)
    assert exp.segments_emitted > len(rays) // 2
    assert_same_segments(got, exp, f64=True)
    assert t.last_stats.ray_steps == exp.ray_steps
    assert t.last_stats.segments == exp.segments_emitted
    assert t.last_stats.object_tests == exp.ray_steps * len(spec.objects)


@pytest.mark.parametrize("name", list(SPECS))
def test_trace_f32_bit_exact(oracle, ctx32, name):
    spec = SPECS[name]
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    exp = osc.trace_rays(rays, abi.LG_PRECISION_F32)
    t = make_tracer(spec, ctx32)
    got = t.trace(rays)
    assert_same_segments(got, exp)
    assert t.last_stats.ray_steps == exp.ray_steps


@pytest.mark.parametrize("slots,merged", [("1", "1"), ("2", "1"), ("2", "0"), ("4", "1")])
def test_slot_counts_give_identical_results(oracle, slots, merged):
    """R ray slots per thread, and whether they share one narrow-phase instance (small scenes) or have one each
    (large scenes), are scheduling choices: results must not depend on them."""
    from light_garden_b200.tracer import Context
    spec = SPECS["C3"]
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    for prec in (abi.LG_PRECISION_F32, abi.LG_PRECISION_F64):
        os.environ["LG_TRACE_SLOTS"], os.environ["LG_TRACE_MERGED"] = slots, merged
        try:
            c = Context(0, prec)
        finally:
            del os.environ["LG_TRACE_SLOTS"], os.environ["LG_TRACE_MERGED"]
        try:
            got = make_tracer(spec, c).trace(rays)
            assert_same_segments(got, osc.trace_rays(rays, prec), f64=prec == abi.LG_PRECISION_F64)
        finally:
            c.close()


def test_emission_matches_reference_formulas(oracle, ctx64):
    """K1: device sincos vs libm sincos may differ in the last place; everything else is exact."""
    spec = SPECS["C3"]   # point, spot and directional lights
    t = make_tracer(spec, ctx64)
    osc = oracle.OracleScene.from_spec(spec)
    for i, l in enumerate(spec.lights):
        got = t.emit_rays(i)
        exp = oracle.emit_rays(l)
        assert np.array_equal(got["origin"], exp["origin"])
        assert np.array_equal(got["color"], exp["color"])
        assert ulp_diff64(got["direction"], exp["direction"]).max() <= 4
        # near the axes a component is ~1e-17 and 4 ulp of it is nothing: also bound the absolute error
        assert np.abs(got["direction"] - exp["direction"]).max() < 1e-15
        assert np.all(got["refractive_index"] == osc.start_medium(l))
    # ragged sub-range
    got = t.emit_rays(0, first=7, count=33)
    exp = oracle.emit_rays(spec.lights[0], first=7, count=33)
    assert np.abs(got["direction"] - exp["direction"]).max() < 1e-15


@pytest.mark.parametrize("name", ["C1", "C5-16"])
def test_trace_all_device_emission_end_to_end(oracle, ctx64, name):
    """Tracer::trace_all with rays emitted on the device: same segments up to the last-place emission
    difference (segment endpoints within 1e-9 relative, identical hit sequences)."""
    spec = SPECS[name]
    osc = oracle.OracleScene.from_spec(spec)
    exp = osc.trace_all(spec.lights, abi.LG_PRECISION_F64)
    t = make_tracer(spec, ctx64)
    seg, tags, f64 = t.trace_all(control_lines=False, return_tags=True)
    assert t.last_stats.primary_rays == spec.total_rays()
    # rays whose whole tag sequence is identical
    same_len = len(seg) == len(exp.seg)
    if same_len and all(np.array_equal(tags[n], exp.tags[n]) for n in ("ray", "generation", "path", "hit_object")):
        scale = np.maximum(1.0, np.abs(exp.f64["b"]))
        assert (np.abs(f64["b"] - exp.f64["b"]) / scale).max() < 1e-9
        assert np.array_equal(seg["color"], exp.seg["color"])
    else:
        # a last-place direction difference flipped a grazing hit somewhere: bound how many rays differ
        got_keys = set(zip(tags["ray"].tolist(), tags["generation"].tolist(), tags["path"].tolist(),
                           tags["hit_object"].tolist()))
        exp_keys = set(zip(exp.tags["ray"].tolist(), exp.tags["generation"].tolist(), exp.tags["path"].tolist(),
                           exp.tags["hit_object"].tolist()))
        diff_rays = {k[0] for k in got_keys ^ exp_keys}
        assert len(diff_rays) <= max(1, spec.total_rays() // 2000), len(diff_rays)


def test_trace_all_appends_control_lines(ctx32):
    """tracer.rs:342-346: the curved mirror's control polygon (3 red segments) follows the traced lines."""
    spec = SPECS["C1"]
    t = make_tracer(spec, ctx32)
    seg = t.trace_all()
    assert np.all(seg["color"][-3:] == np.float32([1, 0, 0, 1]))
    p = spec.objects[1].geo.points
    np.testing.assert_allclose(seg["a"][-3], p[0], rtol=1e-7)
    np.testing.assert_allclose(seg["b"][-1], p[3], rtol=1e-7)


def test_tracer_bookkeeping_mirrors(ctx32):
    """The rest of the Tracer surface a caller of the reference uses around trace_all: get_trace_time is the mean of the
    last 20 trace times in milliseconds (tracer.rs:206-208, 350-354), the control polygon of a curved mirror that is
    still being drawn is shown as well (tracer.rs:341), obj_changed / update_tile_map / new_tile_map leave the scene to be
    lowered again (tracer.rs:126-160)."""
    import math
    from light_garden_b200.scene import CubicBezier
    spec = SPECS["C1"]
    t = make_tracer(spec, ctx32)
    assert math.isnan(t.get_trace_time())
    for _ in range(23):
        t.trace_all(ordered=False, control_lines=False)
    assert len(t._trace_times) == 20 and 0.0 < t.get_trace_time() < 1e3
    assert abs(t.get_trace_time() - sum(t._trace_times) / 20) < 1e-12
    n0 = len(t.trace_all())
    t.add_drawing_object(Object.new_curved_mirror(CubicBezier(((1.2, 0.6), (1.3, 0.8), (1.5, 0.8), (1.6, 0.6)))))
    seg = t.trace_all()
    assert len(seg) == n0 + 3 and np.all(seg["color"][-6:] == np.float32([1, 0, 0, 1]))
    np.testing.assert_allclose(seg["a"][-3], (1.2, 0.6), rtol=1e-7)
    t.finish_drawing_object(True)
    assert len(t.trace_all()) == n0
    t.enable_tile_map(True)
    try:
        before = t.trace_all(control_lines=False)
        t.index_object(2).geo.width *= 0.5            # edited in place: the caller says so
        t.obj_changed(2)
        after = t.trace_all(control_lines=False)
        assert len(after) != len(before) or not np.array_equal(after["b"], before["b"])
        t.new_tile_map(50, 50, 4)
        t.update_tile_map()
        assert np.array_equal(t.trace_all(control_lines=False)["b"], after["b"])
        with pytest.raises(ValueError):
            t.new_tile_map(0, 10, 8)
    finally:
        t.enable_tile_map(False)


def test_shards_partition_the_rays(oracle, ctx32):
    """rank r of R takes the rays r, r + R, ... of every light; the union is the whole trace."""
    spec = SPECS["C1"]
    t = make_tracer(spec, ctx32)
    osc = oracle.OracleScene.from_spec(spec)
    full = osc.trace_all(spec.lights, abi.LG_PRECISION_F32)
    seen = []
    for r in range(3):
        t.set_shard(r, 3)
        seg, tags, _ = t.trace_all(control_lines=False, return_tags=True)
        part = osc.trace_all(spec.lights, abi.LG_PRECISION_F32, rank=r, world=3)
        assert len(seg) == part.segments_emitted
        assert np.array_equal(np.unique(tags["ray"]), np.unique(part.tags["ray"]))
        seen.append(tags["ray"])
    t.set_shard(0, 1)
    allr = np.sort(np.concatenate(seen))
    assert np.array_equal(allr, np.sort(full.tags["ray"]))


# ---- edge cases ---------------------------------------------------------------------------------------
def test_empty_and_ragged_inputs(oracle, ctx32):
    from light_garden_b200.tracer import Tracer
    t = Tracer(scenes.canvas(16 / 9), ctx=ctx32)
    # empty scene, no lights
    seg = t.trace_all()
    assert len(seg) == 0
    # empty scene, rays leave through the canvas: one segment each; 33 rays = one full warp + 1
    t.push_light(PointLight((0.1, 0.2), 33, (0.5, 0.5, 0.5, 0.5)))
    seg, tags, _ = t.trace_all(return_tags=True)
    assert len(seg) == 33 and np.all(tags["hit_object"] == -1) and np.array_equal(tags["ray"], np.arange(33))
    # zero rays
    seg, _, _ = t.trace(np.zeros(0, dtype=abi.RAY_DTYPE))
    assert len(seg) == 0
    # max_bounce = 0: nothing is traced (tracer.rs:373)
    t.max_bounce = 0
    assert len(t.trace_all()) == 0
    # a light below the cutoff emits nothing (tracer.rs:378-384)
    t.max_bounce = 5
    t.index_light(0).color = (0.0005, 0.0005, 0.0005, 0.5)
    assert len(t.trace_all()) == 0


def test_deep_split_tree_matches_oracle(oracle, ctx64):
    """max_bounce = 12 with refraction and a zero cutoff: the full binary split tree (stack depth 11)."""
    from light_garden_b200.tracer import Tracer
    objs = [Object.new_circle((0.0, 0.0), 0.5).with_index(1.5), Object.new_rect((1.0, 0.1), 0.4, 0.9).with_index(1.3),
            Object.new_mirror((-1.2, -0.9), (-1.0, 0.9))]
    t = Tracer(scenes.canvas(16 / 9), ctx=ctx64)
    for o in objs:
        t.push_object(o)
    t.max_bounce = 12
    t.cutoff_color = [0.0, 0.0, 0.0, 0.0]
    light = PointLight((-0.9, 0.05), 40, (0.5, 0.4, 0.3, 0.2))
    osc = oracle.OracleScene(objs, 12, [0.0] * 4, scenes.canvas(16 / 9))
    rays = oracle.emit_rays(light)
    exp = osc.trace_rays(rays, abi.LG_PRECISION_F64)
    got = t.trace(rays)
    assert exp.segments_emitted > 40 * 30
    assert_same_segments(got, exp, f64=True)


def test_table_larger_than_shared_memory(oracle, ctx32):
    """20 000 objects: the broad-phase table (320 KB) no longer fits the 227 KB of shared memory and is read from
    global memory instead; results must not change."""
    from light_garden_b200.tracer import Tracer
    rng = scenes.SplitMix64(0x4C47B16)
    objs = []
    for k in range(20000):
        cx, cy = rng.uniform(-1.7, 1.7), rng.uniform(-0.95, 0.95)
        r = rng.uniform(0.002, 0.004)
        if k % 3 == 0:
            objs.append(Object.new_mirror((cx - r, cy - r), (cx + r, cy + r)))
        elif k % 3 == 1:
            objs.append(Object.new_circle((cx, cy), r).with_index(rng.uniform(1.1, 1.8)))
        else:
            objs.append(Object.new_rect((cx, cy), 2 * r, 1.5 * r).with_index(rng.uniform(1.1, 1.8)))
    light = PointLight((0.01, 0.02), 300, (0.01, 0.008, 0.006, 0.02))
    t = Tracer(scenes.canvas(16 / 9), ctx=ctx32)
    for o in objs:
        t.push_object(o)
    osc = oracle.OracleScene(objs, 5, [0.001] * 4, scenes.canvas(16 / 9))
    rays = oracle.emit_rays(light)
    rays["refractive_index"] = osc.start_medium(light)
    exp = osc.trace_rays(rays, abi.LG_PRECISION_F32)
    got = t.trace(rays)
    assert exp.segments_emitted > 600
    assert_same_segments(got, exp)


@pytest.mark.parametrize("name", list(SPECS))
def test_tile_map_walk_gives_the_all_objects_result(oracle, ctx32, ctx64, name):
    """with Tracer::enable_tile_map the nearest-hit search walks the device-side uniform grid
    instead of testing every object; segments, tags and colours must stay bit-identical to the oracle's
    all-objects loop, in both precisions."""
    spec = SPECS[name]
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    for ctx, prec in ((ctx32, abi.LG_PRECISION_F32), (ctx64, abi.LG_PRECISION_F64)):
        t = make_tracer(spec, ctx)
        assert not t.tile_map_enabled()
        t.enable_tile_map(True)
        try:
            got = t.trace(rays)
            exp = osc.trace_rays(rays, prec)
            assert_same_segments(got, exp, f64=prec == abi.LG_PRECISION_F64)
            assert t.last_stats.ray_steps == exp.ray_steps
        finally:
            t.enable_tile_map(False)


@pytest.mark.parametrize("density,slots", [("0.25", "1"), ("4", "2")])
def test_tile_map_cell_size_and_slots_do_not_matter(oracle, density, slots):
    from light_garden_b200.tracer import Context
    spec = SPECS["C5"]
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    os.environ["LG_GRID_DENSITY"], os.environ["LG_GRID_SLOTS"] = density, slots
    try:
        c = Context(0, abi.LG_PRECISION_F32)
    finally:
        del os.environ["LG_GRID_DENSITY"], os.environ["LG_GRID_SLOTS"]
    try:
        t = make_tracer(spec, c)
        t.enable_tile_map(True)
        assert_same_segments(t.trace(rays), osc.trace_rays(rays, abi.LG_PRECISION_F32))
    finally:
        c.close()


def test_tile_map_with_twenty_thousand_objects(oracle, ctx32):
    """The scene of test_table_larger_than_shared_memory through the grid, device emission, end to end."""
    from light_garden_b200.tracer import Tracer
    rng = scenes.SplitMix64(0x4C47B16)
    t = Tracer(scenes.canvas(16 / 9), ctx=ctx32)
    for k in range(20000):
        cx, cy = rng.uniform(-1.7, 1.7), rng.uniform(-0.95, 0.95)
        r = rng.uniform(0.002, 0.004)
        if k % 3 == 0:
            t.push_object(Object.new_mirror((cx - r, cy - r), (cx + r, cy + r)))
        elif k % 3 == 1:
            t.push_object(Object.new_circle((cx, cy), r).with_index(rng.uniform(1.1, 1.8)))
        else:
            t.push_object(Object.new_rect((cx, cy), 2 * r, 1.5 * r).with_index(rng.uniform(1.1, 1.8)))
    t.push_light(PointLight((0.01, 0.02), 20000, (0.01, 0.008, 0.006, 0.02)))
    seg_all, tag_all, _ = t.trace_all(control_lines=False, return_tags=True)
    steps = t.last_stats.ray_steps
    t.enable_tile_map(True)
    try:
        seg_grid, tag_grid, _ = t.trace_all(control_lines=False, return_tags=True)
    finally:
        t.enable_tile_map(False)
    assert t.last_stats.ray_steps == steps and len(seg_all) > 40000
    assert np.array_equal(tag_all, tag_grid)
    assert seg_all.tobytes() == seg_grid.tobytes()


def test_errors_are_reported_not_hidden(ctx32):
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Context, Tracer
    c = Context(0, abi.LG_PRECISION_F32)
    try:
        with pytest.raises(LightGardenError) as e:       # trace before any scene
            c.call("lg_trace", None)
        assert e.value.code == abi.LG_ERR_STATE
        objs = (abi.LgObject * 1)()
        objs[0].root = 5                                  # node index out of range
        nodes = (abi.LgGeoNode * 1)()
        prm = abi.LgTraceParams()
        with pytest.raises(LightGardenError) as e:
            c.call("lg_scene_set", C.cast(objs, C.c_void_p), 1, C.cast(nodes, C.c_void_p), 1, C.byref(prm))
        assert e.value.code == abi.LG_ERR_INVALID and "range" in e.value.message
        # hostile node tables: a node that is its own child, and 30 levels that share their children (2^30 tokens unfolded)
        nodes = (abi.LgGeoNode * 31)()
        for k in range(31):
            nodes[k].kind, nodes[k].op = abi.LG_GEO_LOGIC, abi.LG_OP_OR
            nodes[k].child_a = nodes[k].child_b = 0
            nodes[k].rot[0] = nodes[k].rot[3] = 1.0
        objs[0].root = 0
        with pytest.raises(LightGardenError) as e:
            c.call("lg_scene_set", C.cast(objs, C.c_void_p), 1, C.cast(nodes, C.c_void_p), 31, C.byref(prm))
        assert e.value.code == abi.LG_ERR_INVALID and "cycle" in e.value.message
        for k in range(30):
            nodes[k].child_a = nodes[k].child_b = k + 1
        nodes[30].kind = abi.LG_GEO_CIRCLE
        nodes[30].p[2] = 0.5
        with pytest.raises(LightGardenError) as e:
            c.call("lg_scene_set", C.cast(objs, C.c_void_p), 1, C.cast(nodes, C.c_void_p), 31, C.byref(prm))
        assert e.value.code == abi.LG_ERR_INVALID and "tokens" in e.value.message
    finally:
        c.close()
    # segment buffer too small for lg_trace: LG_ERR_OVERFLOW, not silent truncation
    spec = SPECS["C1"]
    t = make_tracer(spec, ctx32)
    ctx32.call("lg_segment_capacity_set", 1000)
    try:
        with pytest.raises(LightGardenError) as e:
            t.trace_all()
        assert e.value.code == abi.LG_ERR_OVERFLOW
    finally:
        ctx32.call("lg_segment_capacity_set", 64 << 20)
    # negative refractive index (reachable from the GUI slider, gui/mod.rs:287-294) is rejected
    t2 = Tracer(scenes.canvas(1.0), ctx=ctx32)
    t2.push_object(Object.new_circle((0, 0), 0.5).with_index(-1.0))
    with pytest.raises(LightGardenError) as e:
        t2.sync_scene()
    assert e.value.code == abi.LG_ERR_UNSUPPORTED


def test_directional_light_conventions(oracle, ctx64):
    """light.rs:111 calls start.eval_at_r(-(i as f64) / n); collision2d's eval_at_r is not in the reference.  Both
    readings exist behind ONE named switch (LG_LIGHT_DIRECTIONAL_NEG_R), read by the device and by the oracle."""
    from light_garden_b200.scene import DirectionalLight, LineSegment
    from light_garden_b200.tracer import Tracer
    spec = SPECS["C1"]
    for neg in (False, True):
        light = DirectionalLight((0.5, 0.25, 0.125, 1.0), 257, LineSegment((-0.7, -0.9), (0.6, -0.8)), neg_r=neg)
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx64))
        t.lights[:] = [light]
        t._lights_dirty = True
        got, exp = t.emit_rays(0), oracle.emit_rays(light)
        assert np.array_equal(got["origin"], exp["origin"]) and np.array_equal(got["direction"], exp["direction"])
        step = np.array([0.6 - -0.7, -0.8 - -0.9]) / 257
        np.testing.assert_allclose(got["origin"][5] - got["origin"][4], -step if neg else step, atol=1e-15)
        assert np.array_equal(got["origin"][0], [-0.7, -0.9])          # ray 0 starts at `a` either way


def test_drawing_object_joins_the_start_medium_scan(oracle, ctx64):
    """tracer.rs:279-287: `self.objects.iter().chain(&self.drawing_object)` -- the object being dragged out decides the
    start medium of a light inside it (last match wins) without being traced against (tracer.rs:412-424)."""
    from light_garden_b200.tracer import Tracer
    spec = SPECS["C1"]                      # the point light sits inside the n = 1.73 rect
    t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx64))
    pos = spec.lights[0].position
    assert np.all(t.emit_rays(0)["refractive_index"] == 1.73)
    t.add_drawing_object(Object.new_circle(pos, 0.05).with_index(2.4))
    assert np.all(t.emit_rays(0)["refractive_index"] == 2.4)           # last in the chain wins
    assert np.all(t.emit_rays(1)["refractive_index"] == 1.0)           # the spot light is outside of it
    # not an obstacle: the traced segments are those of the scene without it, starting in its medium
    rays = t.emit_rays(0, 0, 300)
    osc = oracle.OracleScene.from_spec(spec)
    assert_same_segments(t.trace(rays), osc.trace_rays(rays, abi.LG_PRECISION_F64), f64=True)
    t.add_drawing_object(Object.new_mirror((pos[0] - 1, pos[1]), (pos[0] + 1, pos[1])))   # no material: no medium
    assert np.all(t.emit_rays(0)["refractive_index"] == 1.73)
    t.finish_drawing_object(abort=True)
    assert np.all(t.emit_rays(0)["refractive_index"] == 1.73)
    # the same through the plain override field a host can fill in itself
    larr = lights_to_array(spec.lights)
    larr[0].flags |= abi.LG_LIGHT_START_MEDIUM
    larr[0].start_medium = 1.31
    ctx64.call("lg_lights_set", C.cast(larr, C.c_void_p), len(spec.lights))
    out = np.zeros(4, dtype=abi.RAY_DTYPE)
    ctx64.call("lg_emit_rays", 0, 0, 4, abi.array_ptr(out))
    assert np.all(out["refractive_index"] == 1.31)
    larr[0].start_medium = -1.0
    with pytest.raises(Exception):
        ctx64.call("lg_lights_set", C.cast(larr, C.c_void_p), len(spec.lights))
    t.sync_scene(force=True)
    # a drawing light is traced like the others, after them (tracer.rs:279)
    t.add_drawing_light(PointLight((0.3, 0.3), 64, (0.01, 0.01, 0.01, 0.02)))
    seg = t.trace_all(ordered=False, control_lines=False)
    assert t.last_stats.primary_rays == spec.total_rays() + 64 and len(seg) > 0
    t.finish_drawing_light(abort=False)
    assert len(t.lights) == 3 and t.drawing_light is None


def test_trace_rays_rejects_directions_that_are_not_unit(oracle, ctx64, ctx32):
    """Ray::from_origin normalises (light.rs:172); LgRay.direction must already be unit: the broad phase and the range
    filter rely on it, so anything else is refused instead of silently losing hits."""
    spec = SPECS["C1"]
    for ctx in (ctx64, ctx32):
        t = make_tracer(spec, ctx)
        rays = t.emit_rays(0, 0, 8)
        t.trace(rays)                                     # what the library itself emits passes
        for bad in (1.0 + 1e-3, 0.0, float("nan")):
            r = rays.copy()
            r["direction"][3] *= bad
            with pytest.raises(Exception) as e:
                t.trace(r)
            assert "unit" in str(e.value)
    r = rays.copy()                                       # f32-rounded directions: fine for the f32 context only
    r["direction"] = r["direction"].astype(np.float32).astype(np.float64)
    make_tracer(spec, ctx32).trace(r)
    with pytest.raises(Exception):
        make_tracer(spec, ctx64).trace(r)


def test_many_lights_few_rays_and_a_chain_200_generations_deep(oracle, ctx32):
    """600 lights of one to three rays each (the prefix search over lights, shards that get no ray of a light), and a
    split chain far deeper than any BASELINE config."""
    objs = [Object.new_circle((0.0, 0.0), 0.4).with_index(1.5), Object.new_mirror((-1.2, -0.8), (-1.1, 0.8))]
    lights = [PointLight((-0.9 + 0.003 * k, 0.3 * math.sin(k)), 1 + k % 3, (0.3, 0.2, 0.1, 0.5)) for k in range(600)]
    spec = scenes.SceneSpec("many lights", objs, lights, 4, 320, 180)
    osc = oracle.OracleScene.from_spec(spec)
    t = make_tracer(spec, ctx32)
    exp = osc.trace_all(spec.lights, abi.LG_PRECISION_F32)
    seg, tags, _ = t.trace_all(control_lines=False, return_tags=True)
    assert len(seg) == exp.segments_emitted and np.array_equal(tags["ray"], exp.tags["ray"])
    assert np.array_equal(seg["b"], exp.seg["b"])
    seen = 0
    for r in range(7):                                    # 7 shards over lights of 1..3 rays: most shards skip most lights
        t.set_shard(r, 7)
        part = t.trace_all(control_lines=False)
        assert len(part) == osc.trace_all(spec.lights, abi.LG_PRECISION_F32, rank=r, world=7).segments_emitted
        seen += len(part)
    t.set_shard(0, 1)
    assert seen == len(seg)
    # a light inside a lens, nothing culled, 200 generations: every hit splits, the transmitted ray leaves and the
    # reflected one goes round again -- a chain 200 deep whose colours run down to denormals
    deep = scenes.SceneSpec("deep chain", [Object.new_circle((0.0, 0.0), 0.4).with_index(1.5)],
                            [PointLight((0.1, 0.05), 64, (0.5, 0.5, 0.5, 0.5))], 200, 320, 180, (0.0,) * 4)
    exp = oracle.OracleScene.from_spec(deep).trace_all(deep.lights, abi.LG_PRECISION_F32)
    seg, tags, _ = make_tracer(deep, ctx32).trace_all(control_lines=False, return_tags=True)
    assert len(seg) == exp.segments_emitted > 64 * 300 and int(tags["generation"].max()) == 199
    assert np.array_equal(seg["b"], exp.seg["b"]) and np.array_equal(seg["color"], exp.seg["color"])


def test_split_stack_overflow_is_an_error_not_a_silent_loss(ctx32):
    """A ray through 45 glass slabs refracts 90 times in a row and leaves a live reflected sibling behind at every surface:
    more parked branches than the 64 a slot's stack holds (lg_capi.cu: min(max_bounce - 1, 64)).  The reference's Vec
    grows without bound; here the call must fail with LG_ERR_OVERFLOW rather than return a frame with rays missing."""
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Tracer
    t = Tracer(scenes.canvas(16 / 9), ctx=ctx32)
    for k in range(45):
        t.push_object(Object.new_rect((-0.9 + 0.04 * k, 0.0), 0.02, 1.5).with_index(1.5))
    t.push_light(SpotLight((-1.5, 0.0), 0.01, (1.0, 0.02), 16, (0.5, 0.5, 0.5, 0.5)))
    t.max_bounce = 200
    t.cutoff_color = [1e-5] * 4                              # siblings die after a reflection or two: the work stays bounded
    with pytest.raises(LightGardenError) as e:
        t.trace_all()
    assert e.value.code == abi.LG_ERR_OVERFLOW and "stack" in e.value.message
    t.max_bounce = 60                                        # within the stack: the same scene traces
    seg, tags, _ = t.trace_all(control_lines=False, return_tags=True)
    assert int(tags["generation"].max()) == 59 and len(seg) > 16 * 60
