"""Random scenes against the oracle, bit for bit.

The fixed scenes of test_gpu_trace.py are tidy: objects do not overlap, lights sit outside everything.  Here seeded
random scenes mix every object kind the lowering knows (mirrors, curved mirrors, circles, rotated rects, lenses, ellipses,
convex polygons, CSG trees up to three levels deep with their own frames), let them overlap freely (nested media,
the start-medium scan of tracer.rs:280-287, refraction from one object straight into the next), put lights anywhere
(inside objects too), and vary bounce limit and cutoff.  Every scene is traced in both precisions, with the all-objects
loop and with the grid walk, and tags, end points and colours must equal the oracle's.
"""
import math
import os

import numpy as np
import pytest

from light_garden_b200 import abi, scenes
from light_garden_b200.scene import (AND, AND_NOT, OR, Circle, ConvexPolygon, CubicBezier, DirectionalLight, Ellipse,
                                     LineSegment, Logic, Material, Object, PointLight, Rect, SpotLight, rot2)
from util import assert_same_segments, have_cuda, primary_rays

N_SCENES = int(os.environ.get("LG_FUZZ_SCENES", "64"))   # a one-off run with 4000 seeds passed on a B200
A = 16.0 / 9.0


def _pt(rng, sx=A, sy=1.0):
    return (float(rng.uniform(-sx, sx)), float(rng.uniform(-sy, sy)))


def _leaf(rng, local=False):
    """one primitive; local = centred near the origin of a CSG frame"""
    c = (float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1))) if local else _pt(rng, A * 0.9, 0.9)
    k = rng.integers(0, 4)
    if k == 0:
        return Circle(c, float(rng.uniform(0.05, 0.45)))
    if k == 1:
        return Rect(c, rot2(float(rng.uniform(0, math.tau))), float(rng.uniform(0.08, 0.7)), float(rng.uniform(0.08, 0.7)))
    if k == 2:
        return Ellipse(c, float(rng.uniform(0.08, 0.5)), float(rng.uniform(0.05, 0.3)), rot2(float(rng.uniform(0, math.tau))))
    n = int(rng.integers(3, 9))
    r = float(rng.uniform(0.1, 0.4))
    pts = [(r * math.cos(t) * float(rng.uniform(0.6, 1.0)), r * math.sin(t) * float(rng.uniform(0.6, 1.0)))
           for t in sorted(rng.uniform(0, math.tau, n))]
    hull = ConvexPolygon.new_convex_hull(pts)
    if len(hull.points) < 3:
        return Circle(c, r)
    return ConvexPolygon(hull.points, c, rot2(float(rng.uniform(0, math.tau))))


def _tree(rng, depth, local=False):
    if depth == 0 or rng.random() < 0.3:
        return _leaf(rng, local)
    op = (AND, OR, AND_NOT)[int(rng.integers(0, 3))]
    origin = (float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1))) if local else _pt(rng, A * 0.85, 0.85)
    return Logic(op, _tree(rng, depth - 1, True), _tree(rng, depth - 1, True), origin, rot2(float(rng.uniform(0, math.tau))))


def _object(rng):
    k = rng.integers(0, 8)
    n = float(rng.uniform(1.05, 2.5))
    if k == 0:
        return Object.new_mirror(_pt(rng), _pt(rng))
    if k == 1:
        p0 = _pt(rng)
        ctrl = [p0] + [(p0[0] + float(rng.uniform(-0.6, 0.6)), p0[1] + float(rng.uniform(-0.6, 0.6))) for _ in range(3)]
        return Object.new_curved_mirror(CubicBezier(tuple(ctrl)))
    if k == 2:
        return Object.new_circle(_pt(rng, A * 0.9, 0.9), float(rng.uniform(0.05, 0.5))).with_index(n)
    if k == 3:
        return Object.new_lens(_pt(rng, A * 0.8, 0.8), float(rng.uniform(0.8, 2.0)), float(rng.uniform(0.2, 1.5))).with_index(n)
    if k == 4:
        o = Object.new_geo(_leaf(rng)).with_index(n)
        if rng.random() < 0.2:
            o.material_opt = None          # a closed shape that reflects
        return o
    o = Object.new_geo(_tree(rng, int(rng.integers(1, 4)))).with_index(n)
    if rng.random() < 0.15:
        o.material_opt = None
    return o


def _light(rng, n_rays):
    col = tuple(float(v) for v in rng.uniform(0.2, 1.0, 3)) + (float(rng.uniform(0.3, 1.0)),)
    k = rng.integers(0, 3)
    if k == 0:
        return PointLight(_pt(rng, A * 0.95, 0.95), n_rays, col)
    if k == 1:
        d = float(rng.uniform(0, math.tau))
        return SpotLight(_pt(rng, A * 0.95, 0.95), float(rng.uniform(0.2, 3.0)), (math.cos(d), math.sin(d)), n_rays, col)
    a = _pt(rng, A * 0.9, 0.9)
    return DirectionalLight(col, n_rays, LineSegment(a, (a[0] + float(rng.uniform(-0.8, 0.8)), a[1] + float(rng.uniform(-0.8, 0.8)))))


def random_spec(seed):
    rng = np.random.default_rng(0x4C47F000 + seed)
    objs = [_object(rng) for _ in range(int(rng.integers(1, 20)))]
    lights = [_light(rng, int(rng.integers(120, 260))) for _ in range(int(rng.integers(1, 4)))]
    cutoff = (0.001,) * 4 if seed % 3 else (1e-5,) * 4
    return scenes.SceneSpec(f"fuzz{seed}", objs, lights, int(rng.integers(1, 13)), 320, 180, cutoff)


@pytest.fixture(scope="module")
def ctxs():
    from light_garden_b200.tracer import Context
    c = {abi.LG_PRECISION_F32: Context(0, abi.LG_PRECISION_F32), abi.LG_PRECISION_F64: Context(0, abi.LG_PRECISION_F64)}
    yield c
    for v in c.values():
        v.close()


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("seed", range(N_SCENES))
def test_random_scene_equals_the_oracle(oracle, ctxs, seed):
    from light_garden_b200.tracer import Tracer
    spec = random_spec(seed)
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    for prec, ctx in ctxs.items():
        exp = osc.trace_rays(rays, prec)
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        for grid in (False, True):
            t.enable_tile_map(grid)
            try:
                got = t.trace(rays)
                assert_same_segments(got, exp, f64=prec == abi.LG_PRECISION_F64)
                assert t.last_stats.ray_steps == exp.ray_steps, (seed, prec, grid)
            finally:
                t.enable_tile_map(False)


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("mode", [1, 2], ids=["direct", "tiled"])
@pytest.mark.parametrize("seed", range(0, int(os.environ.get("LG_FUZZ_FRAMES", "16"))))
def test_random_scene_frame_equals_the_oracle_accumulation(oracle, ctxs, seed, mode):
    """The frame of a random scene (device emission, trace, line pass through lg_render) against the oracle
    accumulating the device's own segments: the same covered pixels, the same number of fragments, sums within fp32
    association of the f64 sums.  Segments that leave the frame, lie on pixel boundaries or have zero length come
    with the territory here."""
    from light_garden_b200.tracer import Renderer, Tracer
    ctx = ctxs[abi.LG_PRECISION_F32]
    spec = random_spec(seed)
    ctx.call("lg_accumulate_mode_set", mode)
    try:
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        r = Renderer(ctx, spec.width, spec.height)
        r.clear()
        st = r.render(t)
        got = r.read_rgba32f()
        seg = t.trace_all(ordered=False, control_lines=False)
        assert st.segments == len(seg)
        exact = np.zeros((spec.height, spec.width, 4), dtype=np.float64)
        exact[..., 3] = 1.0
        assert oracle.accumulate_segments_f64(exact, seg) == st.pixel_updates
        assert np.array_equal(got[..., 3] > 1, exact[..., 3] > 1)
        rel = np.abs(got - exact) / np.maximum(1.0, np.abs(exact))
        assert rel.max() < (3e-4 if mode == 1 else 2e-5), rel.max()
    finally:
        ctx.call("lg_accumulate_mode_set", 0)


def degenerate_spec():
    """Ties and degenerate shapes: coincident objects (equal hit distances: the order of tracer.rs:412-424 decides),
    surfaces shared by neighbours, zero-size shapes, CSG of a shape with itself, and rays aimed exactly at centres,
    corners, tangent lines and end points."""
    objs = [
        Object.new_circle((0.0, 0.0), 0.25).with_index(1.5),
        Object.new_circle((0.0, 0.0), 0.25).with_index(1.3),                 # the same circle twice
        Object.new_rect((0.75, 0.0), 0.5, 0.5).with_index(1.4),
        Object.new_rect((1.25, 0.0), 0.5, 0.5).with_index(1.6),              # shares the edge x = 1.0 with the one before
        Object.new_mirror((-1.0, -0.5), (-1.0, 0.5)),
        Object.new_mirror((-1.0, 0.5), (-0.5, 0.5)),                         # meets the one before in a corner
        Object.new_mirror((-0.5, -0.75), (-0.5, -0.75)),                     # zero length
        Object.new_circle((0.5, 0.75), 0.0).with_index(1.5),                 # zero radius
        Object(Rect((-0.25, 0.75), rot2(0.0), 0.0, 0.25), Material(1.5), "Rect"),          # zero width
        Object(Logic(AND, Circle((0.0, 0.0), 0.125), Circle((0.0, 0.0), 0.125), (0.0, -0.75), rot2(0.0)), Material(1.5), "Geo"),
        Object(Logic(AND_NOT, Circle((0.0, 0.0), 0.125), Circle((0.0, 0.0), 0.125), (0.5, -0.75), rot2(0.0)), Material(1.5), "Geo"),
        Object(Logic(OR, Rect((0.0, 0.0), rot2(0.0), 0.25, 0.25), Rect((0.25, 0.0), rot2(0.0), 0.25, 0.25), (1.0, -0.75), rot2(0.0)),
               Material(1.5), "Geo"),                                        # two boxes that share an edge
    ]
    lights = [PointLight((-0.5, 0.0), 256, (0.5, 0.5, 0.5, 0.5)),          # 256 rays: exact multiples of 2 pi / 256, axis-aligned ones included
              PointLight((0.0, 0.0), 64, (0.5, 0.4, 0.3, 0.5)),            # at the centre of the doubled circle
              PointLight((1.0, 0.0), 64, (0.3, 0.4, 0.5, 0.5)),            # on the shared edge
              SpotLight((-1.0, 0.5), 1.0, (1.0, -1.0), 33, (0.5, 0.5, 0.5, 0.5)),           # in the mirrors' corner
              DirectionalLight((0.5, 0.5, 0.5, 0.5), 65, LineSegment((-1.5, 0.25), (1.5, 0.25)))]   # grazes the circles' top (y = 0.25)
    return scenes.SceneSpec("degenerate", objs, lights, 8, 320, 180, (1e-4,) * 4)


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
def test_ties_and_degenerate_shapes_equal_the_oracle(oracle, ctxs):
    from light_garden_b200.tracer import Tracer
    spec = degenerate_spec()
    osc = oracle.OracleScene.from_spec(spec)
    rays = primary_rays(oracle, spec, osc)
    # plus hand-aimed rays: at a corner, along an edge, tangent to the circle, along a mirror, from inside the zero-width box
    extra = np.zeros(6, dtype=abi.RAY_DTYPE)
    aims = [((-1.5, 0.9), (0.5, 0.25)), ((0.5, -0.9), (0.5, 0.25)), ((-1.5, 0.25), (1.0, 0.0)), ((-1.0, -0.9), (0.0, 1.0)),
            ((-0.25, 0.75), (1.0, 0.0)), ((1.0, 0.9), (0.0, -1.0))]
    for k, (o, d) in enumerate(aims):
        n = math.hypot(*d)
        extra[k]["origin"], extra[k]["direction"] = o, (d[0] / n, d[1] / n)
        extra[k]["color"], extra[k]["refractive_index"] = (0.5, 0.5, 0.5, 0.5), 1.0
    rays = np.concatenate([rays, extra])
    # a flat ellipse has no frame to intersect in (ORACLE.md 3.7 divides by the semi axes): refused, not traced
    from light_garden_b200._lib import LightGardenError
    flat = Tracer(spec.canvas_bounds, ctx=ctxs[abi.LG_PRECISION_F32])
    flat.push_object(Object(Ellipse((-1.25, -0.5), 0.25, 0.0, rot2(0.0)), Material(1.5), "Ellipse"))
    with pytest.raises(LightGardenError, match="LG_ERR_INVALID"):
        flat.trace(extra)
    for prec, ctx in ctxs.items():
        exp = osc.trace_rays(rays, prec)
        assert exp.segments_emitted > len(rays)
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        for grid in (False, True):
            t.enable_tile_map(grid)
            try:
                assert_same_segments(t.trace(rays), exp, f64=prec == abi.LG_PRECISION_F64)
            finally:
                t.enable_tile_map(False)


def _scaled_geo(g, k):
    import dataclasses
    sc = lambda p: (p[0] * k, p[1] * k)
    if isinstance(g, Circle):
        return Circle(sc(g.origin), g.radius * k)
    if isinstance(g, Rect):
        return Rect(sc(g.origin), g.rotation, g.width * k, g.height * k)
    if isinstance(g, Ellipse):
        return Ellipse(sc(g.origin), g.a * k, g.b * k, g.rot)
    if isinstance(g, ConvexPolygon):
        return ConvexPolygon(tuple(sc(p) for p in g.points), sc(g.origin), g.rotation)
    if isinstance(g, LineSegment):
        return LineSegment(sc(g.a), sc(g.b))
    if isinstance(g, CubicBezier):
        return CubicBezier(tuple(sc(p) for p in g.points))
    if isinstance(g, Logic):
        return Logic(g.op, _scaled_geo(g.a, k), _scaled_geo(g.b, k), sc(g.origin), g.rotation)
    raise TypeError(type(g))


def scaled_spec(spec, k):
    """The same scene in units k times as large (canvas included)."""
    objs = [Object(_scaled_geo(o.geo, k), o.material_opt, o.kind, o.moved) for o in spec.objects]
    lights = []
    for l in spec.lights:
        if isinstance(l, PointLight):
            lights.append(PointLight((l.position[0] * k, l.position[1] * k), l.num_rays, l.color))
        elif isinstance(l, SpotLight):
            lights.append(SpotLight((l.position[0] * k, l.position[1] * k), l.spot_angle, l.spot_direction, l.num_rays, l.color))
        else:
            lights.append(DirectionalLight(l.color, l.num_rays, _scaled_geo(l.start, k)))
    out = scenes.SceneSpec(f"{spec.name}x{k}", objs, lights, spec.max_bounce, spec.width, spec.height, tuple(spec.cutoff_color))
    cb = spec.canvas_bounds
    out.canvas_bounds = Rect((cb.origin[0] * k, cb.origin[1] * k), cb.rotation, cb.width * k, cb.height * k)
    return out


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("k", [2.0 ** -20, 1000.0, 2.0 ** 20], ids=["micro", "pixel-units", "mega"])
def test_scenes_in_other_units_equal_the_oracle(oracle, ctxs, k):
    """The broad phase's rounding margin, the grid's cell size and the table's padding all scale with the coordinate bound
    of the scene: random scenes in units 2^-20, 1000 (a scene kept in pixels) and 2^20 times the usual ones, both widths,
    both nearest-hit paths, against the oracle bit for bit."""
    from light_garden_b200.tracer import Tracer
    for seed in range(10):
        spec = scaled_spec(random_spec(seed), k)
        osc = oracle.OracleScene.from_spec(spec)
        rays = primary_rays(oracle, spec, osc)
        for prec, ctx in ctxs.items():
            exp = osc.trace_rays(rays, prec)
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
            for grid in (False, True):
                t.enable_tile_map(grid)
                try:
                    assert_same_segments(t.trace(rays), exp, f64=prec == abi.LG_PRECISION_F64)
                finally:
                    t.enable_tile_map(False)


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("index", [1.0, 0.5, 1.0 + 2.0 ** -20, 100.0, 1e-3], ids=["one", "half", "one-plus-eps", "hundred", "milli"])
def test_extreme_refractive_indices_equal_the_oracle(oracle, ctxs, index):
    """Every material of the random scenes replaced by one extreme index: no bending at all (reflectance 0), denser outside
    than inside (total internal reflection from the outside), a hair above 1, and ratios of 100 and 1000 either way
    (ORACLE.md 4: Snell, Fresnel, the critical angle)."""
    from light_garden_b200.tracer import Tracer
    for seed in range(8):
        spec = random_spec(seed)
        for o in spec.objects:
            if o.material_opt is not None:
                o.material_opt = Material(index)
        osc = oracle.OracleScene.from_spec(spec)
        rays = primary_rays(oracle, spec, osc)
        for prec, ctx in ctxs.items():
            exp = osc.trace_rays(rays, prec)
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
            for grid in (False, True):
                t.enable_tile_map(grid)
                try:
                    assert_same_segments(t.trace(rays), exp, f64=prec == abi.LG_PRECISION_F64)
                finally:
                    t.enable_tile_map(False)


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("mode", [1, 2], ids=["direct", "tiled"])
def test_frame_of_the_degenerate_scene(oracle, ctxs, mode):
    """The scene of ties and degenerate shapes through lg_render (device emission, waves, both resolves, both widths,
    with and without the grid): rays whose direction turns NaN on the way, zero-length segments and all.  The frame is the
    oracle's accumulation of the device's own segments; the run is part of the compute-sanitizer pass."""
    from light_garden_b200.tracer import Renderer, Tracer
    spec = degenerate_spec()
    for prec, ctx in ctxs.items():
        ctx.call("lg_accumulate_mode_set", mode)
        try:
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
            counts = []
            for grid in (False, True):
                t.enable_tile_map(grid)
                try:
                    r = Renderer(ctx, spec.width, spec.height)
                    r.clear()
                    st = r.render(t)
                    got = r.read_rgba32f()
                    seg = t.trace_all(ordered=False, control_lines=False)
                finally:
                    t.enable_tile_map(False)
                exact = np.zeros((spec.height, spec.width, 4), dtype=np.float64)
                exact[..., 3] = 1.0
                assert oracle.accumulate_segments_f64(exact, seg) == st.pixel_updates
                assert st.segments == len(seg)
                rel = np.abs(got - exact) / np.maximum(1.0, np.abs(exact))
                assert rel.max() < (3e-4 if mode == 1 else 2e-5), rel.max()
                counts.append((st.segments, st.pixel_updates))
            assert counts[0] == counts[1]
        finally:
            ctx.call("lg_accumulate_mode_set", 0)


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("mode", [1, 2], ids=["direct", "tiled"])
def test_random_string_mod_patterns_equal_the_oracle(oracle, ctxs, mode):
    """string_mod.rs over random parameters: every mode (u64 wrapping powers included), tiny and prime moduli, factors
    far above the modulo, colour rules that overlap (the LAST matching rule wins, string_mod.rs:141-150)."""
    from light_garden_b200.scene import ModRemColor, StringMod, StringModMode
    from light_garden_b200.tracer import Renderer
    ctx = ctxs[abi.LG_PRECISION_F32]
    ctx.call("lg_accumulate_mode_set", mode)
    rng = np.random.default_rng(0x4C475D)
    k = 2.0 ** -8
    W = H = 192
    try:
        r = Renderer(ctx, W, H)
        for case in range(24):
            m = int(rng.choice([1, 2, 3, 64, 997, 1024, 2311, 4096]))
            num = int(rng.choice([0, 1, 2, 3, 7, 255, 65537, 2 ** 31 + 11, 2 ** 63 + 5]))
            md = [StringModMode.Mul, StringModMode.Add, StringModMode.Pow, StringModMode.Base][int(rng.integers(0, 4))]
            rules = [ModRemColor(int(rng.integers(1, 9)), int(rng.integers(0, 4)), tuple(k * float(v) for v in rng.integers(0, 3, 4)))
                     for _ in range(int(rng.integers(0, 4)))]
            sm = StringMod(modulo=m, num=num, mode=md, color=(k, k, k, k), modulo_colors=rules)
            r.clear()
            st = r.render_string_mod(sm)
            got = r.read_rgba32f()
            exp = oracle.new_image(W, H)
            n = oracle.accumulate_pairs(exp, oracle.string_mod(sm))
            assert st.segments == m
            # end points pass through sincos on both sides: a handful of fragments may move (test_gpu_accum.py)
            assert abs(int(st.pixel_updates) - int(n)) <= 8, (case, m, num, md, st.pixel_updates, n)
            # the direct resolve adds fragment by fragment like the oracle's fp32 loop; the tiled one adds per-tile partial
            # sums (closer to the exact sum, test_traced_segments_image): thousands of chords meeting in one point
            # (num^i = 1 mod m) differ by fp32 association there
            tol = 1e-6 if mode == 1 else 3e-4
            diff = np.nonzero((np.abs(got - exp) > tol * np.maximum(1.0, np.abs(exp))).any(axis=2))
            assert len(diff[0]) <= 16, (case, m, num, md, len(diff[0]))
            moved = np.nonzero(((got != (0, 0, 0, 1)).any(axis=2)) != ((exp != (0, 0, 0, 1)).any(axis=2)))
            assert len(moved[0]) <= 16, (case, m, num, md, len(moved[0]))
    finally:
        ctx.call("lg_accumulate_mode_set", 0)


HOSTILE_LIGHTS = {
    "spot with a zero direction": lambda: SpotLight((0.5, 0.5), 1.0, (0.0, 0.0), 50, (0.5,) * 4),
    "spot of angle 0": lambda: SpotLight((0.5, 0.5), 0.0, (1.0, 0.0), 50, (0.5,) * 4),
    "NaN position": lambda: PointLight((float("nan"), 0.0), 50, (0.5,) * 4),
    "infinite position": lambda: PointLight((float("inf"), 0.0), 50, (0.5,) * 4),
    "directional light of zero length": lambda: DirectionalLight((0.5,) * 4, 50, LineSegment((0.3, 0.3), (0.3, 0.3))),
    "outside the canvas": lambda: PointLight((10.0, 10.0), 50, (0.5,) * 4),
    "no rays": lambda: PointLight((0.5, 0.5), 0, (0.5,) * 4),
    "NaN colour": lambda: PointLight((0.5, 0.5), 50, (float("nan"), 0.5, 0.5, 0.5)),
}


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("name", list(HOSTILE_LIGHTS))
def test_hostile_light_parameters_equal_the_oracle(oracle, ctxs, name):
    """Lights the GUI can produce by accident (a spot light dragged out by zero pixels) or a file can contain: whatever the
    reference's arithmetic makes of them -- NaN directions, rays that start nowhere -- the device makes the same of them,
    in both widths and both nearest-hit paths, and the line pass swallows the result (part of the compute-sanitizer pass)."""
    from light_garden_b200.tracer import Renderer, Tracer
    objs = [Object.new_circle((0.0, 0.0), 0.4).with_index(1.5), Object.new_mirror((-1.2, -0.8), (-1.1, 0.8))]
    spec = scenes.SceneSpec(name, objs, [HOSTILE_LIGHTS[name]()], 5, 160, 90)
    osc = oracle.OracleScene.from_spec(spec)
    for prec, ctx in ctxs.items():
        exp = osc.trace_all(spec.lights, prec)
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        for grid in (False, True):
            t.enable_tile_map(grid)
            try:
                seg, tags, _ = t.trace_all(control_lines=False, return_tags=True)
                r = Renderer(ctx, spec.width, spec.height)
                r.clear()
                st = r.render(t)
            finally:
                t.enable_tile_map(False)
            assert len(seg) == exp.segments_emitted == st.segments
            assert np.array_equal(seg["a"], exp.seg["a"], equal_nan=True) and np.array_equal(seg["b"], exp.seg["b"], equal_nan=True)
            assert np.array_equal(seg["color"], exp.seg["color"], equal_nan=True)


NAN, INF = float("nan"), float("inf")
HOSTILE_OBJECTS = {
    "circle with a NaN centre": lambda: Object.new_circle((NAN, 0.0), 0.3),
    "circle of infinite radius": lambda: Object.new_circle((0.0, 0.0), INF),
    "circle of negative radius": lambda: Object.new_circle((0.0, 0.0), -0.3),
    "circle of NaN radius": lambda: Object.new_circle((0.0, 0.0), NAN),
    "rect of negative width": lambda: Object.new_rect((0.2, 0.1), -0.4, 0.3),
    "rect of NaN height": lambda: Object.new_rect((0.2, 0.1), 0.4, NAN),
    "rect larger than everything": lambda: Object.new_rect((0.0, 0.0), 1e30, 1e30),
    "mirror to infinity": lambda: Object.new_mirror((0.0, 0.0), (INF, 1.0)),
    "mirror with a NaN end": lambda: Object.new_mirror((0.0, NAN), (1.0, 1.0)),
    "bezier with a NaN control point": lambda: Object.new_curved_mirror(CubicBezier(((0.0, 0.0), (0.3, NAN), (0.6, 0.5), (0.9, 0.0)))),
    "index 0": lambda: Object.new_circle((0.0, 0.0), 0.3).with_index(0.0),
    "index NaN": lambda: Object.new_circle((0.0, 0.0), 0.3).with_index(NAN),
    "index infinite": lambda: Object.new_circle((0.0, 0.0), 0.3).with_index(INF),
    "rotation that is not a rotation": lambda: Object(Rect((0.1, 0.1), (2.0, 0.5, -3.0, 0.0), 0.4, 0.3), Material(1.5), "Rect"),
    "polygon with a NaN vertex": lambda: Object(ConvexPolygon(((0.0, 0.0), (0.4, NAN), (0.2, 0.3))), Material(1.5), "ConvexPolygon"),
}


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("name", list(HOSTILE_OBJECTS))
def test_hostile_object_parameters_are_refused_or_equal_the_oracle(oracle, ctxs, name):
    """Shapes no constructor of the app would make but a RON file can hold.  Each is either refused when the scene is set
    (LG_ERR_INVALID / LG_ERR_UNSUPPORTED, with a message) or traced exactly as the oracle traces it -- never a crash, a
    hang or a read out of bounds (the run is part of the compute-sanitizer pass)."""
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Renderer, Tracer
    objs = [Object.new_mirror((-1.2, -0.8), (-1.1, 0.8)), HOSTILE_OBJECTS[name](), Object.new_circle((0.9, 0.3), 0.2).with_index(1.4)]
    lights = [PointLight((-0.6, 0.1), 96, (0.5, 0.4, 0.3, 0.5)), SpotLight((0.05, 0.02), 1.0, (1.0, 0.3), 33, (0.3, 0.4, 0.5, 0.5))]
    spec = scenes.SceneSpec(name, objs, lights, 6, 160, 90)
    for prec, ctx in ctxs.items():
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        try:
            t.sync_scene(force=True)
        except LightGardenError as e:
            assert e.code in (abi.LG_ERR_INVALID, abi.LG_ERR_UNSUPPORTED) and e.message
            continue
        osc = oracle.OracleScene.from_spec(spec)
        exp = osc.trace_all(spec.lights, prec)
        for grid in (False, True):
            t.enable_tile_map(grid)
            try:
                seg, tags, _ = t.trace_all(control_lines=False, return_tags=True)
                r = Renderer(ctx, spec.width, spec.height)
                r.clear()
                st = r.render(t)
            finally:
                t.enable_tile_map(False)
            assert len(seg) == exp.segments_emitted == st.segments, (name, prec, grid, len(seg), exp.segments_emitted)
            assert np.array_equal(tags["hit_object"], exp.tags["hit_object"])
            assert np.array_equal(seg["b"], exp.seg["b"], equal_nan=True) and np.array_equal(seg["color"], exp.seg["color"], equal_nan=True)


HOSTILE_PARAMS = {
    "canvas of zero size": dict(canvas=Rect((0.0, 0.0), (1.0, 0.0, 0.0, 1.0), 0.0, 0.0)),
    "canvas of negative size": dict(canvas=Rect((0.0, 0.0), (1.0, 0.0, 0.0, 1.0), -3.0, -2.0)),
    "canvas with a NaN": dict(canvas=Rect((NAN, 0.0), (1.0, 0.0, 0.0, 1.0), 3.0, 2.0)),
    "canvas far from the scene": dict(canvas=Rect((100.0, 100.0), (1.0, 0.0, 0.0, 1.0), 3.0, 2.0)),
    "infinite canvas": dict(canvas=Rect((0.0, 0.0), (1.0, 0.0, 0.0, 1.0), INF, INF)),
    "NaN cutoff": dict(cutoff=(NAN, NAN, NAN, NAN)),
    "negative cutoff": dict(cutoff=(-1.0, -1.0, -1.0, -1.0)),
    "cutoff above every colour": dict(cutoff=(10.0, 10.0, 10.0, 10.0)),
    "one generation": dict(max_bounce=1),
}


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("name", list(HOSTILE_PARAMS))
def test_hostile_trace_parameters_are_refused_or_equal_the_oracle(oracle, ctxs, name):
    """Canvas bounds, cutoff colours and bounce limits nobody would choose: refused with a message, or the oracle's result."""
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Renderer, Tracer
    hp = HOSTILE_PARAMS[name]
    spec = random_spec(5)
    spec.max_bounce = hp.get("max_bounce", 6)
    spec.cutoff_color = list(hp.get("cutoff", (0.001,) * 4))
    if "canvas" in hp:
        spec.canvas_bounds = hp["canvas"]
    for prec, ctx in ctxs.items():
        t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
        try:
            t.sync_scene(force=True)
        except LightGardenError as e:
            assert e.code in (abi.LG_ERR_INVALID, abi.LG_ERR_UNSUPPORTED) and e.message
            continue
        osc = oracle.OracleScene.from_spec(spec)
        exp = osc.trace_all(spec.lights, prec)
        for grid in (False, True):
            t.enable_tile_map(grid)
            try:
                seg, tags, _ = t.trace_all(control_lines=False, return_tags=True)
                r = Renderer(ctx, spec.width, spec.height)
                r.clear()
                st = r.render(t)
            finally:
                t.enable_tile_map(False)
            assert len(seg) == exp.segments_emitted == st.segments, (name, prec, grid, len(seg), exp.segments_emitted)
            assert np.array_equal(tags["hit_object"], exp.tags["hit_object"])
            assert np.array_equal(seg["b"], exp.seg["b"], equal_nan=True) and np.array_equal(seg["color"], exp.seg["color"], equal_nan=True)


def _hostile_string_mods():
    from light_garden_b200.scene import Curve, ModRemColor, StringMod, StringModMode
    k = 2.0 ** -8
    col = (k, k, k, k)
    return {
        "turns 2^63": StringMod(modulo=997, num=2, turns=2 ** 63, mode=StringModMode.Mul, color=col),
        "turns 0": StringMod(modulo=997, num=2, turns=0, mode=StringModMode.Mul, color=col),
        "num 2^64 - 1": StringMod(modulo=997, num=2 ** 64 - 1, mode=StringModMode.Mul, color=col),
        "power with a huge exponent": StringMod(modulo=1024, num=2 ** 40 + 3, mode=StringModMode.Pow, color=col),
        "modulo 2 on the Lissajous curve": StringMod(modulo=2, num=1, mode=StringModMode.Add, color=col, init_curve=Curve.Lissajous(3, 2, 0.5)),
        "NaN Lissajous phase": StringMod(modulo=500, num=3, mode=StringModMode.Mul, color=col, init_curve=Curve.Lissajous(3, 2, NAN)),
        "complex base outside the unit disc": StringMod(modulo=300, num=3, mode=StringModMode.Mul, color=col,
                                                        init_curve=Curve.ComplexExp(complex(1.5, 0.7))),
        "complex base NaN": StringMod(modulo=300, num=3, mode=StringModMode.Mul, color=col, init_curve=Curve.ComplexExp(complex(NAN, 0.1))),
        "hypotrochoid with r = s": StringMod(modulo=700, num=5, mode=StringModMode.Mul, color=col, init_curve=Curve.Hypotrochoid(4, 4, 2)),
        "hypotrochoid with zeros": StringMod(modulo=700, num=5, mode=StringModMode.Mul, color=col, init_curve=Curve.Hypotrochoid(0, 0, 0)),
        "NaN colour": StringMod(modulo=700, num=5, mode=StringModMode.Mul, color=(NAN, k, k, k)),
        "infinite colour": StringMod(modulo=700, num=5, mode=StringModMode.Mul, color=(INF, k, k, k)),
        "rule with modulo 0 and a remainder beyond its modulo": StringMod(modulo=700, num=5, mode=StringModMode.Mul, color=col,
            modulo_colors=[ModRemColor(0, 0, (k, 0, 0, k)), ModRemColor(3, 7, (0, k, 0, k)), ModRemColor(2 ** 63, 1, (0, 0, k, k))]),
    }


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("mode", [1, 2], ids=["direct", "tiled"])
@pytest.mark.parametrize("name", list(_hostile_string_mods()))
def test_hostile_string_mod_parameters_are_refused_or_equal_the_oracle(oracle, ctxs, name, mode):
    """string_mod.rs with numbers no slider reaches: u64 products that wrap, curves that degenerate or turn NaN, colours
    that are not numbers.  Refused with a message, or the oracle's fragments (a handful may move: sincos, see
    test_gpu_accum.py) -- no crash, no hang, nothing out of bounds."""
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Renderer
    sm = _hostile_string_mods()[name]
    ctx = ctxs[abi.LG_PRECISION_F32]
    ctx.call("lg_accumulate_mode_set", mode)
    try:
        r = Renderer(ctx, 160, 160)
        r.clear()
        try:
            st = r.render_string_mod(sm)
        except LightGardenError as e:
            assert e.code in (abi.LG_ERR_INVALID, abi.LG_ERR_UNSUPPORTED) and e.message
            return
        got = r.read_rgba32f()
        exp = oracle.new_image(160, 160)
        n = oracle.accumulate_pairs(exp, oracle.string_mod(sm))
        assert abs(int(st.pixel_updates) - int(n)) <= 8, (name, st.pixel_updates, n)
        fin = np.isfinite(exp).all(axis=2) & np.isfinite(got).all(axis=2)
        assert np.array_equal(np.isfinite(exp).all(axis=2), np.isfinite(got).all(axis=2)) or (~fin).sum() <= 16
        diff = np.abs(np.where(fin[..., None], got - exp, 0.0)) > (1e-6 if mode == 1 else 3e-4) * np.maximum(1.0, np.abs(np.where(fin[..., None], exp, 0.0)))
        assert diff.any(axis=2).sum() <= 16, (name, int(diff.any(axis=2).sum()))
    finally:
        ctx.call("lg_accumulate_mode_set", 0)


def test_the_random_scenes_exercise_what_they_claim(oracle):
    """Guard against a generator that quietly stops producing the hard cases: over the seeds there are lights that start
    inside a medium, rays that cross from one object directly into another (two refractive hits in a row with no
    exit between them), deep split trees and total internal reflection."""
    started_inside = deep = 0
    kinds = set()
    for seed in range(N_SCENES):
        spec = random_spec(seed)
        osc = oracle.OracleScene.from_spec(spec)
        for l in spec.lights:
            started_inside += osc.start_medium(l) != 1.0
        for o in spec.objects:
            kinds.add(type(o.geo).__name__)
        rays = primary_rays(oracle, spec, osc)
        exp = osc.trace_rays(rays, abi.LG_PRECISION_F64)
        deep += int(exp.tags["generation"].max()) >= 6
    assert started_inside >= 3 and deep >= 5
    assert {"LineSegment", "CubicBezier", "Circle", "Logic", "Rect", "Ellipse", "ConvexPolygon"} <= kinds


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
def test_grid_corner_cases_equal_the_all_objects_loop(oracle, ctxs):
    """Scenes the uniform grid has to survive: no object at all, one object larger than the canvas, objects far outside
    the canvas, two hundred coincident circles in one cell, and 60 000 tiny ones (cell lists, slot counts and indices well
    past 16 bits) -- each against the oracle, with and without the grid."""
    from light_garden_b200.tracer import Tracer
    rng = np.random.default_rng(0x6A1D)
    tiny = [Object.new_circle((float(rng.uniform(-1.7, 1.7)), float(rng.uniform(-0.95, 0.95))), 0.0015).with_index(1.4) for _ in range(60000)]
    cases = {
        "empty": [],
        "one huge": [Object.new_circle((0.0, 0.0), 50.0).with_index(1.3)],
        "far outside": [Object.new_circle((500.0, -300.0), 2.0).with_index(1.5), Object.new_mirror((-900.0, 1.0), (-900.0, -1.0)),
                        Object.new_circle((0.2, 0.1), 0.3).with_index(1.5)],
        "coincident": [Object.new_circle((0.3, 0.2), 0.25).with_index(1.2 + 0.001 * k) for k in range(200)],
        "sixty thousand": tiny,
    }
    lights = [PointLight((-0.9, 0.05), 120, (0.5, 0.4, 0.3, 0.5)), SpotLight((1.2, -0.6), 0.8, (-1.0, 0.5), 60, (0.3, 0.4, 0.5, 0.5))]
    for name, objs in cases.items():
        spec = scenes.SceneSpec(name, objs, lights, 4, 160, 90)
        osc = oracle.OracleScene.from_spec(spec)
        rays = primary_rays(oracle, spec, osc)
        for prec, ctx in ctxs.items():
            if name == "sixty thousand" and prec == abi.LG_PRECISION_F64:
                continue                                   # the oracle's all-objects loop over 60 000 objects once is enough
            exp = osc.trace_rays(rays, prec)
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
            for grid in (False, True):
                t.enable_tile_map(grid)
                try:
                    assert_same_segments(t.trace(rays), exp, f64=prec == abi.LG_PRECISION_F64)
                finally:
                    t.enable_tile_map(False)
