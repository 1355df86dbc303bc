"""ctypes binding of the CPU oracle (oracle/_build/liblg_oracle.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from light_garden_b200/.
PARITY UNPINNED: see oracle/ORACLE.md.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from light_garden_b200 import abi  # noqa: E402  (struct layouts only)
from light_garden_b200.scene import flatten_objects, lights_to_array, trace_params  # noqa: E402

LIB_PATH = os.path.join(HERE, "_build", "liblg_oracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("lg_oracle_capi.cpp", "lg_oracle.hpp", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.lgo_scene_create.restype = vp
        L.lgo_scene_create.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.POINTER(abi.LgTraceParams)]
        L.lgo_scene_destroy.argtypes = [vp]
        L.lgo_scene_tokens.restype = C.c_uint32
        L.lgo_scene_tokens.argtypes = [vp, vp, C.c_uint32]
        L.lgo_emit_rays.argtypes = [C.POINTER(abi.LgLight), C.c_uint64, C.c_uint64, vp]
        L.lgo_start_medium.restype = C.c_double
        L.lgo_start_medium.argtypes = [vp, C.POINTER(abi.LgLight)]
        L.lgo_contains.restype = C.c_int32
        L.lgo_contains.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_double]
        L.lgo_intersect.restype = C.c_int32
        L.lgo_intersect.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32]
        L.lgo_refract.restype = C.c_int32
        L.lgo_refract.argtypes = [C.c_int32, vp, vp, C.c_double, C.c_double, vp]
        L.lgo_reflect.argtypes = [vp, vp, vp]
        L.lgo_trace_rays.restype = vp
        L.lgo_trace_rays.argtypes = [vp, C.c_int32, vp, C.c_uint64, C.c_int32, C.c_int32, C.c_int32]
        L.lgo_trace_all.restype = vp
        L.lgo_trace_all.argtypes = [vp, C.c_int32, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int32,
                                    C.c_int32, C.c_int32]
        L.lgo_tile_map_enable.restype = C.c_uint64
        L.lgo_tile_map_enable.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.lgo_tile_map_tile.restype = C.c_int32
        L.lgo_tile_map_tile.argtypes = [vp, C.c_double, C.c_double]
        L.lgo_tile_map_slab.restype = C.c_int32
        L.lgo_tile_map_slab.argtypes = [vp, C.c_double, C.c_double]
        L.lgo_tile_map_candidates.restype = C.c_uint32
        L.lgo_tile_map_candidates.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_uint32]
        for f in ("lgo_result_count", "lgo_result_stored", "lgo_result_ray_steps", "lgo_result_primary_rays",
                  "lgo_result_object_tests"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [vp]
        L.lgo_result_seconds.restype = C.c_double
        L.lgo_result_seconds.argtypes = [vp]
        L.lgo_result_copy.argtypes = [vp, vp, vp, vp]
        L.lgo_result_free.argtypes = [vp]
        L.lgo_string_mod.argtypes = [C.POINTER(abi.LgStringMod), vp, C.c_uint32, C.c_uint64, C.c_uint64, vp]
        L.lgo_accumulate_pairs_blend.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_uint64, C.POINTER(abi.LgBlendState)]
        L.lgo_accumulate_pairs_blend.restype = C.c_uint64
        L.lgo_line_crossings.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
        L.lgo_line_crossings.restype = C.c_uint64
        L.lgo_nested_chords.argtypes = [C.POINTER(abi.LgStringMod), vp, C.c_uint32, vp, C.c_uint64, vp]
        L.lgo_image_clear.argtypes = [vp, C.c_int32, C.c_int32, C.c_float]
        L.lgo_accumulate_segments.restype = C.c_uint64
        L.lgo_accumulate_segments.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_uint64, C.c_int32]
        L.lgo_accumulate_segments_f64.restype = C.c_uint64
        L.lgo_accumulate_segments_f64.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_uint64]
        L.lgo_accumulate_pairs.restype = C.c_uint64
        L.lgo_accumulate_pairs.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_uint64, C.c_int32]
        L.lgo_image_to_f16.argtypes = [vp, C.c_uint64, vp]
        L.lgo_image_to_bgra8.argtypes = [vp, C.c_uint64, vp]
        L.lgo_image_to_bgra8_srgb.argtypes = [vp, C.c_uint64, vp]
        L.lgo_num_threads.restype = C.c_int32
        _lib = L
    return _lib


def _vec2(v):
    return (C.c_double * 2)(float(v[0]), float(v[1]))


class TraceResult:
    def __init__(self, h):
        L = lib()
        self.segments_emitted = L.lgo_result_count(h)
        self.ray_steps = L.lgo_result_ray_steps(h)
        self.object_tests = L.lgo_result_object_tests(h)   # Ray::intersect calls of the nearest-hit search
        self.primary_rays = L.lgo_result_primary_rays(h)
        self.seconds = L.lgo_result_seconds(h)
        n = L.lgo_result_stored(h)
        self.seg = np.zeros(n, dtype=abi.SEGMENT_DTYPE)
        self.tags = np.zeros(n, dtype=abi.SEGMENT_TAG_DTYPE)
        self.f64 = np.zeros(n, dtype=abi.SEGMENT_F64_DTYPE)
        if n:
            L.lgo_result_copy(h, abi.array_ptr(self.seg), abi.array_ptr(self.tags), abi.array_ptr(self.f64))
        L.lgo_result_free(h)


class OracleScene:
    """Scene as Tracer holds it (objects + max_bounce + cutoff_color + canvas_bounds), lowered by the oracle."""

    def __init__(self, objects, max_bounce, cutoff_color, canvas_bounds):
        self.L = lib()
        objs, n_obj, nodes, n_nodes = flatten_objects(objects)
        prm = trace_params(max_bounce, cutoff_color, canvas_bounds)
        self.n_obj = n_obj
        self.h = self.L.lgo_scene_create(C.cast(objs, C.c_void_p), n_obj, C.cast(nodes, C.c_void_p), n_nodes,
                                         C.byref(prm))
        if not self.h:
            raise ValueError("oracle: scene rejected")

    @staticmethod
    def from_spec(spec):
        return OracleScene(spec.objects, spec.max_bounce, spec.cutoff_color, spec.canvas_bounds)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.lgo_scene_destroy(self.h)
            self.h = None

    def enable_tile_map(self, enable=True, tiles_x=100, tiles_y=100, slabs=8):
        """Tracer::enable_tile_map (tracer.rs:137-146); the defaults are TileMap::new's in Tracer::new (tracer.rs:27).
        The reference starts with it enabled (tile_map.rs:61).  Returns the number of candidate entries."""
        return self.L.lgo_tile_map_enable(self.h, 1 if enable else 0, tiles_x, tiles_y, slabs)

    def tile_of(self, x, y):
        """TileMap::get_tile (tile_map.rs:134-146): tile index ixx + ixy * tiles_x, or -1 outside the window."""
        return self.L.lgo_tile_map_tile(self.h, float(x), float(y))

    def slab_of(self, dx, dy):
        """Tile::get_index (tile_map.rs:229-235) for a unit direction."""
        return self.L.lgo_tile_map_slab(self.h, float(dx), float(dy))

    def tile_candidates(self, tile, slab):
        """Object indices a ray starting in `tile` with a direction in `slab` is tested against (tracer.rs:395-411)."""
        n = self.L.lgo_tile_map_candidates(self.h, tile, slab, None, 0)
        out = np.zeros(n, dtype=np.int32)
        if n:
            self.L.lgo_tile_map_candidates(self.h, tile, slab, abi.array_ptr(out), n)
        return out

    def tokens(self):
        buf = np.zeros((4096 * 8, 12), dtype=np.float64)
        n = self.L.lgo_scene_tokens(self.h, abi.array_ptr(buf), len(buf))
        return buf[:n]

    def contains(self, obj, p, precision=abi.LG_PRECISION_F64):
        return bool(self.L.lgo_contains(self.h, precision, obj, float(p[0]), float(p[1])))

    def intersect(self, obj, origin, direction, precision=abi.LG_PRECISION_F64):
        """Ray::intersect(&obj.get_geometry()) -> rows (px, py, nx, ny, t)."""
        out = np.zeros((128, 5), dtype=np.float64)
        n = self.L.lgo_intersect(self.h, precision, obj, _vec2(origin), _vec2(direction), abi.array_ptr(out), 128)
        return out[:n]

    def start_medium(self, light):
        from light_garden_b200.scene import light_to_pod
        pod = light_to_pod(light)
        return self.L.lgo_start_medium(self.h, C.byref(pod))

    def trace_rays(self, rays, precision=abi.LG_PRECISION_F64, threads=0, chunk=100, store=True):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        h = self.L.lgo_trace_rays(self.h, precision, abi.array_ptr(rays), len(rays), chunk, threads, 1 if store else 0)
        return TraceResult(h)

    def trace_all(self, lights, precision=abi.LG_PRECISION_F64, rank=0, world=1, stride=1, threads=0, chunk=100,
                  store=True):
        arr = lights_to_array(lights)
        h = self.L.lgo_trace_all(self.h, precision, C.cast(arr, C.c_void_p), len(lights), rank, world, stride, chunk,
                                 threads, 1 if store else 0)
        return TraceResult(h)


def emit_rays(light, first=0, count=None):
    from light_garden_b200.scene import light_to_pod
    pod = light_to_pod(light)
    if count is None:
        count = int(light.num_rays) - first
    out = np.zeros(count, dtype=abi.RAY_DTYPE)
    lib().lgo_emit_rays(C.byref(pod), first, count, abi.array_ptr(out))
    return out


def refract(d, n, n1, n2, precision=abi.LG_PRECISION_F64):
    out = (C.c_double * 5)()
    has = lib().lgo_refract(precision, _vec2(d), _vec2(n), n1, n2, out)
    return (out[0], out[1]), ((out[2], out[3]) if has else None), out[4]


def reflect(d, n):
    out = (C.c_double * 2)()
    lib().lgo_reflect(_vec2(d), _vec2(n), out)
    return (out[0], out[1])


def string_mod(sm, first=0, count=None):
    pod, rules, n = sm.to_pod()
    if count is None:
        count = sm.modulo - first
    out = np.zeros(count, dtype=abi.VERTEX_PAIR_DTYPE)
    lib().lgo_string_mod(C.byref(pod), C.cast(rules, C.c_void_p), n, first, count, abi.array_ptr(out))
    return out


def accumulate_pairs_blend(img, pairs, color, alpha, constant=(0.0, 0.0, 0.0, 0.0)):
    """The line pass under a blend state: color / alpha = (src_factor, dst_factor, operation) in abi.LG_BF_* / LG_BO_*."""
    pairs = np.ascontiguousarray(pairs, dtype=abi.VERTEX_PAIR_DTYPE)
    st = abi.LgBlendState()
    st.color.src_factor, st.color.dst_factor, st.color.operation = color
    st.alpha.src_factor, st.alpha.dst_factor, st.alpha.operation = alpha
    st.constant[:] = [float(v) for v in constant]
    h, w = img.shape[:2]
    return lib().lgo_accumulate_pairs_blend(abi.array_ptr(img), w, h, abi.array_ptr(pairs), len(pairs), C.byref(st))


def line_crossings(lines):
    """StringMod::line_crossings_as_points (string_mod.rs:87-101) over chords given as vertex pairs -> n x 2 f64."""
    lines = np.ascontiguousarray(lines, dtype=abi.VERTEX_PAIR_DTYPE)
    n = lib().lgo_line_crossings(abi.array_ptr(lines), len(lines), None, 0)
    out = np.zeros((n, 2), dtype=np.float64)
    if n:
        lib().lgo_line_crossings(abi.array_ptr(lines), len(lines), abi.array_ptr(out), n)
    return out


def nested_chords(inner, points):
    """inner.draw_init_points(points) (string_mod.rs:103-122) -> vertex pairs."""
    pod, rules, n = inner.to_pod()
    points = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros(inner.modulo if len(points) else 0, dtype=abi.VERTEX_PAIR_DTYPE)
    if len(points):
        lib().lgo_nested_chords(C.byref(pod), C.cast(rules, C.c_void_p), n, abi.array_ptr(points), len(points),
                                abi.array_ptr(out))
    return out


def string_mod_draw(sm):
    """StringMod::draw (string_mod.rs:152-158), nested or not -> vertex pairs."""
    if sm.nested is None:
        return string_mod(sm)
    return nested_chords(sm.nested, line_crossings(string_mod(sm)))


def new_image(width, height, clear_alpha=1.0):
    img = np.zeros((height, width, 4), dtype=np.float32)
    lib().lgo_image_clear(abi.array_ptr(img), width, height, C.c_float(clear_alpha))
    return img


def accumulate_segments(img, seg, threads=0):
    seg = np.ascontiguousarray(seg, dtype=abi.SEGMENT_DTYPE)
    h, w = img.shape[:2]
    return lib().lgo_accumulate_segments(abi.array_ptr(img), w, h, abi.array_ptr(seg), len(seg), threads)


def accumulate_segments_f64(img, seg):
    """Same fragments, f64 sums (img: float64 H x W x 4)."""
    seg = np.ascontiguousarray(seg, dtype=abi.SEGMENT_DTYPE)
    h, w = img.shape[:2]
    return lib().lgo_accumulate_segments_f64(abi.array_ptr(img), w, h, abi.array_ptr(seg), len(seg))


def accumulate_pairs(img, pairs, threads=0):
    pairs = np.ascontiguousarray(pairs, dtype=abi.VERTEX_PAIR_DTYPE)
    h, w = img.shape[:2]
    return lib().lgo_accumulate_pairs(abi.array_ptr(img), w, h, abi.array_ptr(pairs), len(pairs), threads)


def to_f16(img):
    out = np.zeros(img.shape, dtype=np.float16)
    lib().lgo_image_to_f16(abi.array_ptr(np.ascontiguousarray(img)), img.size, abi.array_ptr(out))
    return out


def to_bgra8(img):
    """Renderer::make_screenshot's conversion of the Rgba16Float frame (renderer.rs:313-328)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    out = np.zeros(img.shape[:2] + (4,), dtype=np.uint8)
    lib().lgo_image_to_bgra8(abi.array_ptr(img), img.shape[0] * img.shape[1], abi.array_ptr(out))
    return out


def to_bgra8_srgb(img):
    """The 8-bit surface target (render_to_texture off, sub_render_pass.rs:59-63; Bgra8UnormSrgb,
    renderer.rs:207-209): saturate, sRGB-encode the colour, round to nearest (ORACLE.md 8.7)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    out = np.zeros(img.shape[:2] + (4,), dtype=np.uint8)
    lib().lgo_image_to_bgra8_srgb(abi.array_ptr(img), img.shape[0] * img.shape[1], abi.array_ptr(out))
    return out


def num_threads():
    return lib().lgo_num_threads()
