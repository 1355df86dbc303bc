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


def small_specs():
    """CPU-oracle sized versions of the BASELINE configs (same shapes, fewer rays)."""
    return {
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
