"""ctypes mirror of include/light_garden_b200.h (struct layouts + prototypes).

Shared by the product binding (light_garden_b200._lib) and, for the struct
definitions only, by the oracle's test binding (oracle/lg_oracle.py): both
sides of a parity test are fed the very same buffers.
"""
import ctypes as C

import numpy as np

LG_ABI_VERSION = 2

LG_OK, LG_ERR_INVALID, LG_ERR_CUDA, LG_ERR_NOMEM = 0, -1, -2, -3
LG_ERR_UNSUPPORTED, LG_ERR_OVERFLOW, LG_ERR_NCCL, LG_ERR_STATE = -4, -5, -6, -7
ERROR_NAMES = {
    0: "LG_OK", -1: "LG_ERR_INVALID", -2: "LG_ERR_CUDA", -3: "LG_ERR_NOMEM", -4: "LG_ERR_UNSUPPORTED",
    -5: "LG_ERR_OVERFLOW", -6: "LG_ERR_NCCL", -7: "LG_ERR_STATE",
}

LG_PRECISION_F32, LG_PRECISION_F64 = 0, 1
LG_GEO_CIRCLE, LG_GEO_RECT, LG_GEO_SEGMENT, LG_GEO_BEZIER, LG_GEO_LOGIC, LG_GEO_ELLIPSE = 0, 1, 2, 3, 4, 5
LG_GEO_POLYGON, LG_GEO_POINTS, LG_POLYGON_MAX_VERTICES = 6, 7, 32
# wgpu::BlendFactor / BlendOperation (gui/settings.rs:59-105)
(LG_BF_ZERO, LG_BF_ONE, LG_BF_SRC, LG_BF_ONE_MINUS_SRC, LG_BF_SRC_ALPHA, LG_BF_ONE_MINUS_SRC_ALPHA, LG_BF_DST,
 LG_BF_ONE_MINUS_DST, LG_BF_DST_ALPHA, LG_BF_ONE_MINUS_DST_ALPHA, LG_BF_SRC_ALPHA_SATURATED, LG_BF_CONSTANT,
 LG_BF_ONE_MINUS_CONSTANT) = range(13)
LG_BO_ADD, LG_BO_SUBTRACT, LG_BO_REVERSE_SUBTRACT, LG_BO_MIN, LG_BO_MAX = range(5)
LG_OP_AND, LG_OP_OR, LG_OP_ANDNOT = 0, 1, 2
LG_LIGHT_POINT, LG_LIGHT_DIRECTIONAL, LG_LIGHT_SPOT = 0, 1, 2
LG_LIGHT_DIRECTIONAL_NEG_R, LG_LIGHT_START_MEDIUM = 1, 2   # LgLight.flags
LG_SM_ADD, LG_SM_MUL, LG_SM_POW, LG_SM_BASE = 0, 1, 2, 3
LG_CURVE_CIRCLE, LG_CURVE_COMPLEX_EXP, LG_CURVE_HYPOTROCHOID, LG_CURVE_LISSAJOUS = 0, 1, 2, 3
LG_RGBA32F, LG_RGBA16F, LG_BGRA8_GAMMA, LG_BGRA8_SRGB = 0, 1, 2, 3


class LgGeoNode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("child_a", C.c_int32), ("child_b", C.c_int32),
                ("p", C.c_double * 8), ("rot", C.c_double * 4)]


class LgObject(C.Structure):
    _fields_ = [("root", C.c_int32), ("has_material", C.c_int32), ("refractive_index", C.c_double)]


class LgTraceParams(C.Structure):
    _fields_ = [("max_bounce", C.c_uint32), ("cutoff_color", C.c_float * 4), ("_pad", C.c_uint32),
                ("canvas_tlbr", C.c_double * 4)]


class LgLight(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("num_rays", C.c_uint64), ("color", C.c_float * 4),
                ("position", C.c_double * 2), ("b", C.c_double * 2), ("spot_angle", C.c_double),
                ("spot_direction", C.c_double * 2), ("start_medium", C.c_double)]


class LgTraceStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("ray_steps", C.c_uint64), ("object_tests", C.c_uint64),
                ("segments", C.c_uint64), ("pixel_updates", C.c_uint64), ("trace_ms", C.c_float),
                ("accumulate_ms", C.c_float), ("trace_launches", C.c_uint32), ("accumulate_launches", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class LgStringMod(C.Structure):
    _fields_ = [("modulo", C.c_uint64), ("num", C.c_uint64), ("turns", C.c_uint64), ("mode", C.c_int32),
                ("curve", C.c_int32), ("color", C.c_float * 4), ("curve_p", C.c_double * 4)]


class LgModRemColor(C.Structure):
    _fields_ = [("modulo", C.c_uint64), ("rem", C.c_uint64), ("color", C.c_float * 4)]


class LgBlendComponent(C.Structure):
    _fields_ = [("src_factor", C.c_int32), ("dst_factor", C.c_int32), ("operation", C.c_int32)]


class LgBlendState(C.Structure):
    _fields_ = [("color", LgBlendComponent), ("alpha", LgBlendComponent), ("constant", C.c_float * 4)]


# bulk data travels as numpy structured arrays with the same layout
RAY_DTYPE = np.dtype([("origin", "<f8", 2), ("direction", "<f8", 2), ("color", "<f4", 4),
                      ("refractive_index", "<f8")], align=True)
SEGMENT_DTYPE = np.dtype([("a", "<f4", 2), ("b", "<f4", 2), ("color", "<f4", 4)], align=True)
VERTEX_PAIR_DTYPE = np.dtype([("a", "<f8", 2), ("b", "<f8", 2), ("color_a", "<f4", 4), ("color_b", "<f4", 4)],
                             align=True)
SEGMENT_TAG_DTYPE = np.dtype([("ray", "<u8"), ("path", "<u8"), ("generation", "<u4"), ("hit_object", "<i4")],
                             align=True)
SEGMENT_F64_DTYPE = np.dtype([("a", "<f8", 2), ("b", "<f8", 2)], align=True)

SIZES = {
    "LgGeoNode": (C.sizeof(LgGeoNode), 112), "LgObject": (C.sizeof(LgObject), 16),
    "LgTraceParams": (C.sizeof(LgTraceParams), 56), "LgLight": (C.sizeof(LgLight), 96),
    "LgRay": (RAY_DTYPE.itemsize, 56), "LgSegment": (SEGMENT_DTYPE.itemsize, 32),
    "LgVertexPair": (VERTEX_PAIR_DTYPE.itemsize, 64), "LgSegmentTag": (SEGMENT_TAG_DTYPE.itemsize, 24),
    "LgSegmentF64": (SEGMENT_F64_DTYPE.itemsize, 32), "LgModRemColor": (C.sizeof(LgModRemColor), 32),
    "LgStringMod": (C.sizeof(LgStringMod), 80), "LgTraceStats": (C.sizeof(LgTraceStats), 56),
    "LgBlendState": (C.sizeof(LgBlendState), 40),
}

_ctx = C.c_void_p
_p = C.c_void_p
# name -> argtypes; every entry point returns int32 except lg_last_error
PROTOTYPES = {
    "lg_abi_version": [],
    "lg_device_count": [C.POINTER(C.c_int32)],
    "lg_create": [C.c_int32, C.c_int32, C.POINTER(_ctx)],
    "lg_destroy": [_ctx],
    "lg_last_error": [_ctx],
    "lg_scene_set": [_ctx, _p, C.c_uint32, _p, C.c_uint32, C.POINTER(LgTraceParams)],
    "lg_drawing_object_set": [_ctx, _p, _p, C.c_uint32],
    "lg_lights_set": [_ctx, _p, C.c_uint32],
    "lg_shard_set": [_ctx, C.c_uint32, C.c_uint32],
    "lg_segment_capacity_set": [_ctx, C.c_uint64],
    "lg_accumulate_mode_set": [_ctx, C.c_int32],
    "lg_tile_map_enable": [_ctx, C.c_int32],
    "lg_blend_set": [_ctx, C.POINTER(LgBlendState)],
    "lg_tags_enable": [_ctx, C.c_int32],
    "lg_emit_rays": [_ctx, C.c_uint32, C.c_uint64, C.c_uint64, _p],
    "lg_trace": [_ctx, C.POINTER(LgTraceStats)],
    "lg_trace_rays": [_ctx, _p, C.c_uint64, C.POINTER(LgTraceStats)],
    "lg_segments_count": [_ctx, C.POINTER(C.c_uint64)],
    "lg_segments_read": [_ctx, _p, _p, _p, C.c_uint64, C.POINTER(C.c_uint64)],
    "lg_image_configure": [_ctx, C.c_uint32, C.c_uint32],
    "lg_image_clear": [_ctx, C.c_float],
    "lg_accumulate_traced": [_ctx, C.POINTER(LgTraceStats)],
    "lg_accumulate_segments": [_ctx, _p, C.c_uint64, C.POINTER(LgTraceStats)],
    "lg_string_mod": [_ctx, C.POINTER(LgStringMod), _p, C.c_uint32, C.c_uint64, C.c_uint64,
                      C.POINTER(LgTraceStats)],
    "lg_string_mod_nested": [_ctx, C.POINTER(LgStringMod), C.POINTER(LgStringMod), _p, C.c_uint32,
                             C.POINTER(LgTraceStats)],
    "lg_string_mod_nested_read": [_ctx, _p, C.c_uint64, _p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    "lg_render": [_ctx, C.POINTER(LgTraceStats)],
    "lg_render_overlap_set": [_ctx, C.c_int32, C.c_uint32],
    "lg_image_read": [_ctx, C.c_int32, _p, C.c_size_t],
    "lg_image_export_fd": [_ctx, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)],
    "lg_image_export_refresh": [_ctx, C.c_int32],
    "lg_import_fd_read": [C.c_int32, C.c_int32, C.c_uint64, _p, C.c_uint64],
    "lg_comm_unique_id": [_p],
    "lg_comm_init_rank": [_ctx, _p, C.c_int32, C.c_int32],
    "lg_comm_init_all": [C.POINTER(_ctx), C.c_int32],
    "lg_image_reduce": [_ctx, C.c_int32, C.POINTER(C.c_float)],
    "lg_reduce_mode_set": [_ctx, C.c_int32],
    "lg_comm_destroy": [_ctx],
    "lg_stream_handle": [_ctx, C.POINTER(C.c_uint64)],
    "lg_image_device_ptr": [_ctx, C.POINTER(C.c_uint64)],
    "lg_launch_count": [_ctx, C.POINTER(C.c_uint64)],
    "lg_host_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "lg_host_free": [C.c_void_p],
    "lg_measure_fma_peak": [_ctx, C.c_int32, C.c_int32, C.POINTER(C.c_double)],
    "lg_measure_red_peak": [_ctx, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_double)],
    "lg_measure_tile_rmw_peak": [_ctx, C.c_int32, C.POINTER(C.c_double)],
}


def bind(lib):
    """Attach argtypes/restype for every symbol the header declares."""
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.argtypes = args
        fn.restype = C.c_char_p if name == "lg_last_error" else C.c_int32
    return lib


def array_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None
