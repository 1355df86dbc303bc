"""Helpers shared by the parity tests."""
import numpy as np

from light_garden_b200 import abi, scenes


def have_cuda():
    try:
        import ctypes
        from light_garden_b200 import _lib
        n = ctypes.c_int32(0)
        return _lib.load().lg_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def ellipse_spec(total_rays=3000):
    """Ellipses (object.rs:38-45), alone and inside CSG trees, plus a mirror: SURVEY.md §8f rank 2."""
    import math
    from light_garden_b200.scene import (AND_NOT, OR, Circle, Ellipse, Logic, Material, Object, PointLight, Rect,
                                         SpotLight, rot2, rot2_identity)
    objs = [
        Object.new_ellipse((-0.9, 0.2), 0.5, 0.2).with_index(1.5),
        Object(Ellipse((0.5, -0.3), 0.25, 0.6, rot2(0.6)), Material(1.33), "Ellipse"),
        Object(Logic(AND_NOT, Ellipse((0.0, 0.0), 0.45, 0.3, rot2(0.2)), Circle((0.15, 0.0), 0.18), (0.9, 0.45), rot2(-0.4)),
               Material(1.7), "Geo"),
        Object(Logic(OR, Ellipse((0.0, 0.0), 0.3, 0.12, rot2_identity()), Rect((0.0, 0.0), rot2(0.9), 0.15, 0.5),
                     (-0.2, -0.55), rot2(1.1)), Material(1.2), "Geo"),
        Object.new_mirror((-1.6, -0.9), (-1.2, 0.9)),
    ]
    lights = [PointLight((0.05, 0.1), total_rays // 2, (0.012, 0.01, 0.008, 0.02)),
              SpotLight((-1.5, 0.8), 1.2, (1.0, -0.5), total_rays - total_rays // 2, (0.006, 0.01, 0.014, 0.02))]
    return scenes.SceneSpec("ellipses", objs, lights, 6, 480, 270)


def polygon_spec(total_rays=3000):
    """Convex polygons (object.rs:34-36): refractive, as a mirror-less prism, inside CSG trees, rotated frames, and
    the 32-vertex maximum: SURVEY.md §8f rank 2."""
    import math
    from light_garden_b200.scene import (AND, AND_NOT, Circle, ConvexPolygon, Logic, Material, Object, PointLight,
                                         SpotLight, rot2)
    ring = [(0.3 * math.cos(2 * math.pi * k / 32), 0.22 * math.sin(2 * math.pi * k / 32)) for k in range(32)]
    objs = [
        Object.new_convex_polygon([(-1.2, -0.5), (-0.6, -0.55), (-0.9, 0.1), (-0.9, -0.3)]).with_index(1.5),  # prism
        Object.new_convex_polygon([(0.2, 0.2), (0.8, 0.25), (0.95, 0.6), (0.5, 0.85), (0.15, 0.6)]).with_index(1.33),
        Object(ConvexPolygon(tuple(ring), (0.6, -0.45), rot2(0.5)), Material(1.7), "ConvexPolygon"),
        Object(Logic(AND_NOT, ConvexPolygon(((-0.3, -0.2), (0.3, -0.2), (0.3, 0.2), (-0.3, 0.2))), Circle((0.1, 0.0), 0.15),
                     (-0.4, 0.55), rot2(-0.3)), Material(1.2), "Geo"),
        Object(Logic(AND, ConvexPolygon(((0.0, -0.25), (0.25, 0.2), (-0.25, 0.2))), Circle((0.0, 0.0), 0.2),
                     (-1.3, 0.5), rot2(1.0)), Material(2.0), "Geo"),
        Object.new_mirror((1.4, -0.9), (1.6, 0.9)),
    ]
    lights = [PointLight((0.0, -0.1), total_rays // 2, (0.012, 0.01, 0.008, 0.02)),
              SpotLight((-1.6, -0.8), 1.0, (1.0, 0.6), total_rays - total_rays // 2, (0.006, 0.01, 0.014, 0.02))]
    return scenes.SceneSpec("polygons", objs, lights, 6, 480, 270)


def small_specs():
    """CPU-oracle sized versions of the BASELINE configs (same shapes, fewer rays)."""
    return {
        "ELL": ellipse_spec(),
        "POLY": polygon_spec(),
        "C1": scenes.c1_default(total_rays=6000, width=480, height=270),
        "C2": scenes.c2_cavity(total_rays=1500, max_bounce=64, width=480, height=270),
        "C3": scenes.c3_refraction(total_rays=6000, grid=16, width=480, height=270),
        "C5": scenes.c5_large(n_lights=2, rays_per_light=1500, grid=64, width=480, height=270),
        "C5-16": scenes.c5_large(n_lights=3, rays_per_light=2001, grid=16, width=480, height=270),
    }


def primary_rays(oracle, spec, osc=None):
    """Light::get_rays for every light (oracle emission) with the start medium of tracer.rs:280-287."""
    osc = osc or oracle.OracleScene.from_spec(spec)
    parts = []
    for l in spec.lights:
        r = oracle.emit_rays(l)
        r["refractive_index"] = osc.start_medium(l)
        parts.append(r)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=abi.RAY_DTYPE)


def assert_same_segments(got, exp, f64=False):
    """got = (seg, tags, f64) from the device, sorted; exp = oracle TraceResult (already in reference order)."""
    seg, tags, s64 = got
    assert len(seg) == exp.segments_emitted == len(exp.seg), (len(seg), exp.segments_emitted)
    for name in ("ray", "generation", "path", "hit_object"):
        bad = np.nonzero(tags[name] != exp.tags[name])[0]
        assert bad.size == 0, f"tag {name} differs first at {bad[:5]}: {tags[name][bad[:5]]} vs {exp.tags[name][bad[:5]]}"
    for name in ("a", "b", "color"):
        a, b = seg[name].view(np.uint32), exp.seg[name].view(np.uint32)
        bad = np.nonzero((a != b).any(axis=1))[0]
        assert bad.size == 0, f"segment {name} differs at {bad[:5]}: {seg[name][bad[:3]]} vs {exp.seg[name][bad[:3]]}"
    if f64:
        for name in ("a", "b"):
            a, b = s64[name].view(np.uint64), exp.f64[name].view(np.uint64)
            assert np.array_equal(a, b), f"f64 endpoint {name} differs"


def ulp_diff64(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    a = np.where(a < 0, np.int64(-2 ** 63) - a, a)
    b = np.where(b < 0, np.int64(-2 ** 63) - b, b)
    return np.abs(a - b)
