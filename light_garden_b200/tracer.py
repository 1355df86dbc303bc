"""Host-side mirror of the reference's Tracer (src/light_garden/tracer.rs) and of
the LineList half of its Renderer (src/renderer.rs, src/sub_render_pass.rs) on
top of the C ABI.  Same method names and argument meaning as the Rust code so
that tests read like tests of the reference would; every computation happens in
liblight_garden_b200.so on the GPU.
"""
import ctypes as C
import weakref
from typing import List, Optional

import numpy as np

from . import abi
from ._lib import LightGardenError, check, load
from .scene import (CubicBezier, DirectionalLight, Object, PointLight, Rect, SpotLight, StringMod, flatten_objects,
                    lights_to_array, trace_params)


class Context:
    """Owns one lg_ctx (one device, one stream)."""

    def __init__(self, device: int = 0, precision: int = abi.LG_PRECISION_F32):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.lg_create(int(device), int(precision), C.byref(self.h))
        if rc != 0:
            raise LightGardenError(rc, "lg_create failed (no CUDA device? this library has no CPU fallback)")
        self.device = device
        self.precision = precision

    def close(self):
        if self.h:
            self.lib.lg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        check(self.h, getattr(self.lib, name)(self.h, *args))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self.call("lg_launch_count", C.byref(n))
        return n.value


def pinned_array(shape, dtype):
    """numpy array over page-locked host memory (lg_host_alloc).  The allocation lives exactly as long as the array
    (and its views, which keep it alive through .base): a finalizer on the owning array calls lg_host_free."""
    lib = load()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = C.c_void_p()
    check(None, lib.lg_host_alloc(n, C.byref(p)))
    buf = (C.c_ubyte * n).from_address(p.value)
    owner = np.frombuffer(buf, dtype=dt)        # every view derived from it has .base == owner
    weakref.finalize(owner, lib.lg_host_free, C.c_void_p(p.value))
    return owner.reshape(shape)


def sort_segments(seg, tags, f64=None):
    """Device order -> the reference's order: light -> ray -> generation -> queue order."""
    order = np.lexsort((tags["path"], tags["generation"], tags["ray"]))
    return seg[order], tags[order], (f64[order] if f64 is not None else None)


class Tracer:
    """tracer.rs:4-17.  Scene-edit methods keep the reference's names; trace_all/trace run on the device."""

    def __init__(self, canvas_bounds: Rect, device: int = 0, precision: int = abi.LG_PRECISION_F32,
                 ctx: Optional[Context] = None):
        self.ctx = ctx or Context(device, precision)
        self.objects: List[Object] = []
        self.lights: list = []
        self.max_bounce = 5                              # tracer.rs:37
        self.cutoff_color = [0.001, 0.001, 0.001, 0.001]  # tracer.rs:38
        self.chunk_size = 100                            # tracer.rs:39 (rayon chunking; unused on the device)
        self.canvas_bounds = canvas_bounds
        self.last_stats = abi.LgTraceStats()
        self.drawing_object: Optional[Object] = None     # tracer.rs:8-9
        self.drawing_light = None
        self._scene_dirty = True
        self._lights_dirty = True
        self._sent_params = None
        self._rank, self._world = 0, 1

    # -- scene editing: tracer.rs:48-181 ------------------------------------------------------
    def clear(self):
        self.drawing_object = self.drawing_light = None
        self.objects.clear()
        self.lights.clear()
        self._scene_dirty = self._lights_dirty = True

    def clear_objects(self):
        self.drawing_object = None
        self.objects.clear()
        self._scene_dirty = self._lights_dirty = True

    def add_drawing_object(self, obj: Object):
        """tracer.rs:61-63: the object being dragged out takes part in the start-medium scan only (tracer.rs:281)."""
        self.drawing_object = obj
        self._lights_dirty = True

    def finish_drawing_object(self, abort: bool):
        """tracer.rs:65-72"""
        obj, self.drawing_object = self.drawing_object, None
        self._lights_dirty = True
        if not abort and obj is not None:
            self.push_object(obj)

    def add_drawing_light(self, light):
        """tracer.rs:74-76: traced like any other light, after them (tracer.rs:279)."""
        self.drawing_light = light
        self._lights_dirty = True

    def finish_drawing_light(self, abort: bool):
        """tracer.rs:78-84"""
        light, self.drawing_light = self.drawing_light, None
        self._lights_dirty = True
        if not abort and light is not None:
            self.push_light(light)

    def push_object(self, obj: Object):
        self.objects.append(obj)
        self._scene_dirty = True

    def push_light(self, light):
        self.lights.append(light)
        self._lights_dirty = True

    def index_object(self, ix):
        self._scene_dirty = True
        return self.objects[ix]

    def index_light(self, ix):
        self._lights_dirty = True
        return self.lights[ix]

    def replace_object(self, ix, obj):
        self.objects[ix] = obj
        self._scene_dirty = True

    def remove_object(self, ix):
        del self.objects[ix]
        self._scene_dirty = True

    def remove_light(self, ix):
        del self.lights[ix]
        self._lights_dirty = True

    def object_iterator(self):
        return iter(self.objects)

    def light_iterator(self):
        return iter(self.lights)

    def resize(self, bounds: Rect):
        self.canvas_bounds = bounds
        self._scene_dirty = True

    def load(self, data: str):
        """Tracer::load (tracer.rs:190-204): RON text of (Vec<Object>, Vec<Light>)."""
        from .ron import load_scene
        objects, lights = load_scene(data)
        self.clear()
        self.objects, self.lights = objects, lights

    def serialize(self) -> str:
        """Tracer::serialize (tracer.rs:183-188): RON text of (Vec<Object>, Vec<Light>) that load() reads back."""
        from .ron import serialize_scene
        return serialize_scene(self.objects, self.lights)

    def enable_tile_map(self, enable: bool):
        """Tracer::enable_tile_map (tracer.rs:137-146): nearest-hit search through the device grid (same segments)."""
        self._tile_map_enabled = bool(enable)
        self.ctx.call("lg_tile_map_enable", 1 if enable else 0)

    def tile_map_enabled(self) -> bool:
        """Tracer::tile_map_enabled (tracer.rs:126-129)."""
        return getattr(self, "_tile_map_enabled", False)

    def new_tile_map(self, num_tiles_x: int, num_tiles_y: int, num_slabs: int):
        """Tracer::new_tile_map (tracer.rs:148-160).  The device grid sizes itself from the scene (one cell per ~object,
        LG_GRID_DENSITY) and has no angular slabs; the arguments are accepted for the caller's sake and the state they
        leave behind is the reference's: a fresh map, enabled as before."""
        if num_tiles_x <= 0 or num_tiles_y <= 0 or num_slabs <= 0:
            raise ValueError("new_tile_map: tile and slab counts must be positive")
        self._scene_dirty = True

    def update_tile_map(self):
        """Tracer::update_tile_map (tracer.rs:131-135): the device grid is rebuilt with the next trace."""
        self._scene_dirty = True

    def obj_changed(self, obj_index: int):
        """Tracer::obj_changed (tracer.rs:126-129): an object was edited in place through index_object."""
        self.objects[obj_index].moved = True
        self._scene_dirty = True

    def get_trace_time(self) -> float:
        """Tracer::get_trace_time (tracer.rs:206-208): mean of the last (up to 20) trace_all times in milliseconds
        (tracer.rs:350-354); NaN before the first one, like the reference's 0 / 0.  Device time of the trace kernels."""
        t = getattr(self, "_trace_times", [])
        return sum(t) / len(t) if t else float("nan")

    # -- device state -----------------------------------------------------------------------------
    def set_shard(self, rank: int, world: int):
        self._rank, self._world = rank, world
        self.ctx.call("lg_shard_set", rank, world)

    def sync_scene(self, force: bool = False):
        """lg_scene_set / lg_lights_set only when the host copy changed (the reference's `moved` dirty bit,
        object.rs:54): every edit method sets a flag; the scalar knobs (max_bounce, cutoff_color, canvas_bounds) are
        public fields in the reference, so their values are compared with what was last sent.  Code that mutates an
        Object or Light it kept a reference to must go through index_object / index_light again (or pass force)."""
        owner = getattr(self.ctx, "_scene_owner", None)
        if owner is None or owner() is not self:      # another Tracer drove this context in between
            force = True
            self.ctx._scene_owner = weakref.ref(self)
        cb = self.canvas_bounds
        params = (int(self.max_bounce), tuple(float(v) for v in self.cutoff_color),
                  (tuple(cb.origin), tuple(cb.rotation), float(cb.width), float(cb.height)), id(self.ctx))
        if force or self._scene_dirty or params != self._sent_params:
            objs, n_obj, nodes, n_nodes = flatten_objects(self.objects)
            prm = trace_params(self.max_bounce, self.cutoff_color, self.canvas_bounds)
            self.ctx.call("lg_scene_set", C.cast(objs, C.c_void_p), n_obj, C.cast(nodes, C.c_void_p), n_nodes,
                          C.byref(prm))
            self._sent_params = params
            self._scene_dirty = False
        if force or self._lights_dirty:
            if self.drawing_object is not None:   # before the lights: their start media depend on it
                o, _, nd, nn = flatten_objects([self.drawing_object])
                self.ctx.call("lg_drawing_object_set", C.cast(o, C.c_void_p), C.cast(nd, C.c_void_p), nn)
            else:
                self.ctx.call("lg_drawing_object_set", None, None, 0)
            lights = list(self.lights) + ([self.drawing_light] if self.drawing_light is not None else [])
            arr = lights_to_array(lights)
            self.ctx.call("lg_lights_set", C.cast(arr, C.c_void_p), len(lights))
            self._lights_dirty = False

    # -- tracing -----------------------------------------------------------------------------------
    def emit_rays(self, light_index: int, first: int = 0, count: Optional[int] = None):
        """Light::get_rays (light.rs:17-23) for one light, computed on the device."""
        self.sync_scene()
        if count is None:
            count = int(self.lights[light_index].num_rays) - first
        out = np.zeros(count, dtype=abi.RAY_DTYPE)
        self.ctx.call("lg_emit_rays", light_index, first, count, abi.array_ptr(out))
        return out

    def _read_segments(self, tags: bool):
        n = C.c_uint64()
        self.ctx.call("lg_segments_count", C.byref(n))
        seg = np.zeros(n.value, dtype=abi.SEGMENT_DTYPE)
        tg = np.zeros(n.value, dtype=abi.SEGMENT_TAG_DTYPE) if tags else None
        f64 = np.zeros(n.value, dtype=abi.SEGMENT_F64_DTYPE) if (
            tags and self.ctx.precision == abi.LG_PRECISION_F64) else None
        got = C.c_uint64()
        self.ctx.call("lg_segments_read", abi.array_ptr(seg), abi.array_ptr(tg) if tags else None,
                      abi.array_ptr(f64) if f64 is not None else None, n.value, C.byref(got))
        return seg, tg, f64

    def trace_all(self, ordered: bool = True, control_lines: bool = True, return_tags: bool = False):
        """Tracer::trace_all (tracer.rs:276-358): all lights' rays -> segments, in the reference's order."""
        self.ctx.call("lg_tags_enable", 1 if (ordered or return_tags) else 0)
        self.sync_scene()
        st = abi.LgTraceStats()
        self.ctx.call("lg_trace", C.byref(st))
        self.last_stats = st
        self._trace_times = (getattr(self, "_trace_times", []) + [float(st.trace_ms)])[-20:]  # trace_time_vd, tracer.rs:350-354
        seg, tg, f64 = self._read_segments(ordered or return_tags)
        if ordered:
            seg, tg, f64 = sort_segments(seg, tg, f64)
        if control_lines:  # tracer.rs:342-346
            shown = self.objects + ([self.drawing_object] if self.drawing_object is not None else [])
            extra = [cl for ob in shown for cl in ob.get_control_lines()]
            if extra:
                add = np.zeros(len(extra), dtype=abi.SEGMENT_DTYPE)
                for i, (a, b, col) in enumerate(extra):
                    add[i]["a"], add[i]["b"], add[i]["color"] = a, b, col
                seg = np.concatenate([seg, add])
        return (seg, tg, f64) if return_tags else seg

    def trace(self, rays: np.ndarray, ordered: bool = True):
        """Tracer::trace (tracer.rs:360-493) over caller supplied primary rays (abi.RAY_DTYPE)."""
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        self.ctx.call("lg_tags_enable", 1)
        self.sync_scene()
        st = abi.LgTraceStats()
        self.ctx.call("lg_trace_rays", abi.array_ptr(rays), len(rays), C.byref(st))
        self.last_stats = st
        seg, tg, f64 = self._read_segments(True)
        if ordered:
            seg, tg, f64 = sort_segments(seg, tg, f64)
        return seg, tg, f64


class Renderer:
    """The LineList pass of renderer.rs into the Rgba16Float target (texture_renderer.rs:5), on the device."""

    def __init__(self, ctx: Context, width: int, height: int):
        self.ctx = ctx
        self.width, self.height = int(width), int(height)
        ctx.call("lg_image_configure", self.width, self.height)
        self.last_stats = abi.LgTraceStats()

    def clear(self, clear_alpha: float = 1.0):
        """LoadOp::Clear(BLACK) (renderer.rs:174-177)."""
        self.ctx.call("lg_image_clear", C.c_float(clear_alpha))

    def render_traced(self):
        """sub_rpass_lines.render for the segments of the last trace (sub_render_pass.rs:205-212)."""
        st = abi.LgTraceStats()
        self.ctx.call("lg_accumulate_traced", C.byref(st))
        self.last_stats = st
        return st

    def render_lines(self, pairs: np.ndarray):
        """update_vertex_buffer + render for host vertex pairs (abi.VERTEX_PAIR_DTYPE)."""
        pairs = np.ascontiguousarray(pairs, dtype=abi.VERTEX_PAIR_DTYPE)
        st = abi.LgTraceStats()
        self.ctx.call("lg_accumulate_segments", abi.array_ptr(pairs), len(pairs), C.byref(st))
        self.last_stats = st
        return st

    def render_string_mod(self, sm: StringMod, first: int = 0, count: int = 0):
        """LightGarden::draw in Mode::StringMod (mod.rs:681-689) + the line pass."""
        pod, rules, n = sm.to_pod()
        st = abi.LgTraceStats()
        if sm.nested is not None:   # StringMod::draw, string_mod.rs:152-158
            if first or count:
                raise ValueError("a nested string mod is drawn as a whole")
            ipod, irules, ni = sm.nested.to_pod()
            self.ctx.call("lg_string_mod_nested", C.byref(pod), C.byref(ipod), C.cast(irules, C.c_void_p), ni, C.byref(st))
        else:
            self.ctx.call("lg_string_mod", C.byref(pod), C.cast(rules, C.c_void_p), n, first, count, C.byref(st))
        self.last_stats = st
        return st

    def nested_crossings(self):
        """The outer chords (abi.VERTEX_PAIR_DTYPE) and crossing points (n x 2 f64) of the last nested string mod."""
        nl, npts = C.c_uint64(), C.c_uint64()
        self.ctx.call("lg_string_mod_nested_read", None, 0, None, 0, C.byref(nl), C.byref(npts))
        lines = np.zeros(nl.value, dtype=abi.VERTEX_PAIR_DTYPE)
        pts = np.zeros((npts.value, 2), dtype=np.float64)
        self.ctx.call("lg_string_mod_nested_read", abi.array_ptr(lines) if nl.value else None, nl.value,
                      abi.array_ptr(pts) if npts.value else None, npts.value, C.byref(nl), C.byref(npts))
        return lines, pts

    def render(self, tracer: Tracer):
        """Renderer::render's trace + line pass fused (waves through the bounded segment buffer)."""
        tracer.ctx.call("lg_tags_enable", 0)
        tracer.sync_scene()
        st = abi.LgTraceStats()
        self.ctx.call("lg_render", C.byref(st))
        self.last_stats = st
        return st

    def set_blend(self, color=None, alpha=None, constant=(0.0, 0.0, 0.0, 0.0)):
        """LightGarden.color_state_descriptor.blend (mod.rs:57-73, gui/settings.rs:59-127): each component is
        (src_factor, dst_factor, operation) in abi.LG_BF_* / abi.LG_BO_*; None restores the default state."""
        if color is None and alpha is None:
            self.ctx.call("lg_blend_set", None)
            return
        st = abi.LgBlendState()
        c = color if color is not None else (abi.LG_BF_ONE, abi.LG_BF_ONE, abi.LG_BO_ADD)
        a = alpha if alpha is not None else (abi.LG_BF_SRC_ALPHA, abi.LG_BF_ONE, abi.LG_BO_ADD)
        st.color.src_factor, st.color.dst_factor, st.color.operation = c
        st.alpha.src_factor, st.alpha.dst_factor, st.alpha.operation = a
        st.constant[:] = [float(v) for v in constant]
        self.ctx.call("lg_blend_set", C.byref(st))

    def read_rgba32f(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        self.ctx.call("lg_image_read", abi.LG_RGBA32F, abi.array_ptr(out), 0)
        return out

    def make_screenshot(self, pitch: int = 0):
        """Renderer::make_screenshot's pixel conversion (renderer.rs:294-328): H x W x 4 uint8 in [b, g, r, a] order.
        `pitch` > 0 pads every row (wgpu's COPY_BYTES_PER_ROW_ALIGNMENT layout, renderer.rs:250-255)."""
        row = pitch if pitch else self.width * 4
        out = np.zeros((self.height, row), dtype=np.uint8)
        self.ctx.call("lg_image_read", abi.LG_BGRA8_GAMMA, abi.array_ptr(out), row)
        return out[:, : self.width * 4].reshape(self.height, self.width, 4)

    def read_surface_bgra8(self, pitch: int = 0):
        """The frame as the 8-bit surface target holds it when `render_to_texture` is off (sub_render_pass.rs:59-63;
        the screenshot's Bgra8UnormSrgb, renderer.rs:207-209): saturated, colour sRGB-encoded, [b, g, r, a]."""
        row = pitch if pitch else self.width * 4
        out = np.zeros((self.height, row), dtype=np.uint8)
        self.ctx.call("lg_image_read", abi.LG_BGRA8_SRGB, abi.array_ptr(out), row)
        return out[:, : self.width * 4].reshape(self.height, self.width, 4)

    def export_fd(self, fmt=abi.LG_RGBA16F):
        """(fd, allocation bytes) of the frame in `fmt` as device memory another API can import (Vulkan OPAQUE_FD /
        cuMemImportFromShareableHandle): the display hand-off without renderer.rs's host copy.  The caller closes fd."""
        fd, n = C.c_int32(-1), C.c_uint64(0)
        self.ctx.call("lg_image_export_fd", fmt, C.byref(fd), C.byref(n))
        return fd.value, n.value

    def export_refresh(self, fmt=abi.LG_RGBA16F):
        """Converts the current image into the exported frame; importers may read when this returns."""
        self.ctx.call("lg_image_export_refresh", fmt)

    def read_rgba16f(self, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 4), dtype=np.float16)
        self.ctx.call("lg_image_read", abi.LG_RGBA16F, abi.array_ptr(out), 0)
        return out
