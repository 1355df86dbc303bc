/*
 * light_garden_b200 — C ABI of the B200-native trace + line-accumulation path.
 *
 * This is the drop-in boundary for sphereflow/light_garden's two hot paths.
 * The reference has no FFI of its own (it is one Rust binary); the boundary is
 * the two Rust call sites that a patched app redirects here:
 *
 *   B1  LightGarden::draw -> Tracer::trace_all() -> Vec<(P2, Color)>
 *       (src/light_garden/mod.rs:691, src/light_garden/tracer.rs:276-358)
 *   B2  Renderer::render -> SubRenderPass::update_vertex_buffer + render into
 *       the Rgba16Float target (src/renderer.rs:431,164-188,
 *       src/sub_render_pass.rs:188-212, src/texture_renderer.rs:5)
 *
 * Conventions
 *   - every function returns int32_t: 0 = LG_OK, negative = error; nothing
 *     unwinds across the boundary (the reference panics instead:
 *     src/light_garden/tracer.rs:192, src/framework.rs:44,51,84).
 *   - lg_last_error(ctx) returns a human readable message for the last failure.
 *   - the caller owns every host buffer it passes in; the library copies.
 *   - the library owns device buffers, its stream and its NCCL communicator.
 *   - one lg_ctx drives ONE device; a context is thread-compatible, not
 *     thread-safe (the reference mutates the Tracer only from the winit event
 *     loop thread, src/framework.rs:180-258).
 *   - calls are blocking (return after the stream drained) unless stated.
 *   - all structs are plain C, naturally aligned, no padding surprises:
 *     sizes are asserted in csrc/lg_capi.cu and in tests/test_abi.py.
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device lg_create
 * fails with LG_ERR_CUDA.
 */
#ifndef LIGHT_GARDEN_B200_H
#define LIGHT_GARDEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------- */
enum {
  LG_OK = 0,
  LG_ERR_INVALID = -1,     /* bad argument / malformed scene                  */
  LG_ERR_CUDA = -2,        /* CUDA runtime error or no device                 */
  LG_ERR_NOMEM = -3,       /* device or host allocation failed                */
  LG_ERR_UNSUPPORTED = -4, /* legal in the reference, not supported here      */
  LG_ERR_OVERFLOW = -5,    /* segment buffer / split stack capacity exceeded  */
  LG_ERR_NCCL = -6,        /* NCCL missing or failed                          */
  LG_ERR_STATE = -7        /* call order violated (e.g. trace before scene)   */
};

/* ---- precision of the device trace ------------------------------------- */
/* The reference computes geometry in f64 (collision2d Float = f64) and colour
 * in f32 (src/light_garden/light.rs:7). LG_PRECISION_F32 is the throughput
 * path named by BASELINE.json's north_star; LG_PRECISION_F64 is the
 * reference-width path used for exact parity runs. */
enum { LG_PRECISION_F32 = 0, LG_PRECISION_F64 = 1 };

/* ---- geometry (collision2d `Geo`, exhaustively matched at
 *      src/light_garden/drawer.rs:23-99) ----------------------------------- */
enum {
  LG_GEO_CIRCLE = 0,  /* Geo::GeoCircle       p = {ox, oy, radius}            */
  LG_GEO_RECT = 1,    /* Geo::GeoRect         p = {ox, oy, width, height}, rot */
  LG_GEO_SEGMENT = 2, /* Geo::GeoLineSegment  p = {ax, ay, bx, by}            */
  LG_GEO_BEZIER = 3,  /* Geo::GeoCubicBezier  p = {x0,y0,x1,y1,x2,y2,x3,y3}   */
  LG_GEO_LOGIC = 4,   /* Geo::GeoLogic        p = {ox, oy}, rot, op, a, b     */
  LG_GEO_ELLIPSE = 5, /* Geo::GeoEllipse      p = {ox, oy, a, b}, rot         */
  /* Geo::GeoConvexPolygon (object.rs:34-36, ConvexPolygon::new_convex_hull): p = {ox, oy}, rot,
   * op = number of hull vertices (3..32), child_a = first LG_GEO_POINTS node. The vertices are in
   * the polygon's local frame (world = origin + rot * local), already in hull order. */
  LG_GEO_POLYGON = 6,
  /* continuation node of a polygon: up to 4 vertices p = {x0,y0,..,x3,y3} (op = how many, 1..4),
   * child_a = next LG_GEO_POINTS node or -1 */
  LG_GEO_POINTS = 7
};
#define LG_POLYGON_MAX_VERTICES 32
/* collision2d LogicOp (src/light_garden/mod.rs:368-372) */
enum { LG_OP_AND = 0, LG_OP_OR = 1, LG_OP_ANDNOT = 2 };

/* One node of a geometry tree. `rot` is a nalgebra Rotation2 exactly as serde
 * writes it, column-major [m11, m21, m12, m22] = [cos, sin, -sin, cos]
 * (default.ron:24-29). Children of a LOGIC node live in that node's local
 * frame: world = origin + rot * local (src/light_garden/object.rs:393-410).
 * lg_scene_set / lg_drawing_object_set answer LG_ERR_INVALID for a node whose
 * parameters are not finite or exceed 1e12 in magnitude, whose radius, width or
 * height is negative, or whose `rot` is not orthonormal within 1e-6: the
 * bounding circles and the grid are built from these numbers, and a shape they
 * cannot bound would be traced differently from the reference. */
typedef struct LgGeoNode {
  int32_t kind;    /* LG_GEO_*                                               */
  int32_t op;      /* LG_OP_* (LOGIC only)                                   */
  int32_t child_a; /* node index (LOGIC only), else -1                       */
  int32_t child_b; /* node index (LOGIC only), else -1                       */
  double p[8];
  double rot[4];
} LgGeoNode; /* 112 bytes */

/* Object{object_enum, material_opt} (src/light_garden/object.rs:51-55).
 * StraightMirror -> SEGMENT root without material, CurvedMirror -> BEZIER root
 * without material, Lens -> LOGIC(And, circle, circle) with material. */
typedef struct LgObject {
  int32_t root;         /* index into the LgGeoNode array                    */
  int32_t has_material; /* material_opt.is_some()                            */
  double refractive_index;
} LgObject; /* 16 bytes */

/* Tracer{max_bounce, cutoff_color, canvas_bounds}
 * (src/light_garden/tracer.rs:4-17, defaults 37-39). canvas is given the way
 * the app builds it: Rect::from_tlbr(top, left, bottom, right)
 * (src/sub_render_pass.rs:156). */
typedef struct LgTraceParams {
  uint32_t max_bounce;
  float cutoff_color[4];
  uint32_t _pad;
  double canvas_tlbr[4];
} LgTraceParams; /* 56 bytes */

/* ---- lights (src/light_garden/light.rs:10-14) --------------------------- */
enum { LG_LIGHT_POINT = 0, LG_LIGHT_DIRECTIONAL = 1, LG_LIGHT_SPOT = 2 };
/* LgLight.flags */
enum {
  /* DirectionalLight::set_num_rays calls start.eval_at_r(-(i as f64) / n) (light.rs:111); what collision2d's
   * LineSegment::eval_at_r does with a negative parameter is not in the reference (the crate is not vendored).
   * Default: the origins walk the drawn segment, a + (i/n)(b - a) (ORACLE.md 6.3).  With this flag the call is
   * taken literally on the natural parametrisation eval_at_r(r) = a + r (b - a): origins a - (i/n)(b - a).
   * One named switch, read by the device code and by the oracle; directional-light parity is UNVERIFIED either way. */
  LG_LIGHT_DIRECTIONAL_NEG_R = 1,
  /* start_medium holds the refractive index of the medium the light's rays start in and replaces the scan over
   * the scene's objects (tracer.rs:279-287).  That scan chains `self.drawing_object` -- the object the user is
   * dragging out, which is not in the scene yet -- behind the objects, so it wins when it contains the light: the
   * host evaluates that one `contains` itself and passes the index here. */
  LG_LIGHT_START_MEDIUM = 2
};
typedef struct LgLight {
  int32_t kind;
  int32_t flags;            /* LG_LIGHT_* flags above                        */
  uint64_t num_rays;
  float color[4];
  double position[2];       /* Point/Spot position; Directional: start.a     */
  double b[2];              /* Directional: start.b                          */
  double spot_angle;        /* Spot only                                     */
  double spot_direction[2]; /* Spot only                                     */
  double start_medium;      /* read when flags & LG_LIGHT_START_MEDIUM       */
} LgLight; /* 96 bytes */

/* One primary ray as Tracer::trace receives it (tracer.rs:360-367). */
typedef struct LgRay {
  double origin[2];
  double direction[2]; /* unit: | |d|^2 - 1 | <= 16 eps of the context's precision, else LG_ERR_INVALID
                        * (Ray::from_origin normalises, light.rs:172; the broad phase relies on it) */
  float color[4];
  double refractive_index; /* medium the ray starts in                        */
} LgRay; /* 56 bytes */

/* ---- segments ------------------------------------------------------------ */
/* Compact device segment: what SubRenderPass::update_vertex_buffer would
 * upload for one vertex pair of a traced ray (positions cast to f32, one colour
 * for both ends; sub_render_pass.rs:189-196, tracer.rs:451-452,474-475). */
typedef struct LgSegment {
  float a[2];
  float b[2];
  float color[4];
} LgSegment; /* 32 bytes */

/* General vertex pair (P2, Color),(P2, Color) for host supplied lines: control
 * polygons, grid, drawer overlays (tracer.rs:342-349, mod.rs:692). */
typedef struct LgVertexPair {
  double a[2];
  double b[2];
  float color_a[4];
  float color_b[4];
} LgVertexPair; /* 64 bytes */

/* Order / provenance tag of a traced segment. The reference's output order is
 * light -> ray -> generation -> queue order; the device emits
 * unordered, and (ray, generation, path) is the sort key that restores it.
 * `path` holds one bit per ancestor generation, most significant = first
 * split: 0 = reflected child (pushed first, tracer.rs:456-458), 1 = refracted
 * child (tracer.rs:460-472). */
typedef struct LgSegmentTag {
  uint64_t ray;        /* global primary-ray index (lights concatenated)      */
  uint64_t path;
  uint32_t generation; /* 0 = primary ray                                    */
  int32_t hit_object;  /* object index, or -1 when the ray left via canvas   */
} LgSegmentTag; /* 24 bytes */

typedef struct LgSegmentF64 {
  double a[2];
  double b[2];
} LgSegmentF64; /* 32 bytes: unrounded endpoints, F64 contexts with tags only */

/* ---- string mod (src/light_garden/string_mod.rs:4-15) -------------------- */
enum { LG_SM_ADD = 0, LG_SM_MUL = 1, LG_SM_POW = 2, LG_SM_BASE = 3 };
/* Curve (string_mod.rs:182-188); curve_p: COMPLEX_EXP {re, im}; HYPOTROCHOID {r, s, d};
 * LISSAJOUS {a, b, delta} */
enum { LG_CURVE_CIRCLE = 0, LG_CURVE_COMPLEX_EXP = 1, LG_CURVE_HYPOTROCHOID = 2, LG_CURVE_LISSAJOUS = 3 };
typedef struct LgModRemColor {
  uint64_t modulo;
  uint64_t rem;
  float color[4];
} LgModRemColor; /* 32 bytes */
typedef struct LgStringMod {
  uint64_t modulo;
  uint64_t num;
  uint64_t turns;
  int32_t mode;  /* LG_SM_*                                                  */
  int32_t curve; /* LG_CURVE_*                                               */
  float color[4];
  double curve_p[4];
} LgStringMod; /* 80 bytes */

/* ---- blend state (wgpu::BlendState as the GUI edits it, src/gui/settings.rs:59-105;
 *      default src/light_garden/mod.rs:57-73) --------------------------------- */
enum { /* wgpu::BlendFactor (discriminants of wgpu-types = the order of gui/settings.rs:59-73) */
  LG_BF_ZERO = 0, LG_BF_ONE, LG_BF_SRC, LG_BF_ONE_MINUS_SRC, LG_BF_SRC_ALPHA, LG_BF_ONE_MINUS_SRC_ALPHA,
  LG_BF_DST, LG_BF_ONE_MINUS_DST, LG_BF_DST_ALPHA, LG_BF_ONE_MINUS_DST_ALPHA, LG_BF_SRC_ALPHA_SATURATED,
  LG_BF_CONSTANT, LG_BF_ONE_MINUS_CONSTANT
};
enum { /* wgpu::BlendOperation (discriminants of wgpu-types; the GUI lists them at gui/settings.rs:99-105) */
  LG_BO_ADD = 0, LG_BO_SUBTRACT, LG_BO_REVERSE_SUBTRACT, LG_BO_MIN, LG_BO_MAX
};
typedef struct LgBlendComponent {
  int32_t src_factor, dst_factor, operation;
} LgBlendComponent;
typedef struct LgBlendState {
  LgBlendComponent color, alpha;
  float constant[4]; /* wgpu blend constant (the app never sets it: transparent black) */
} LgBlendState;      /* 40 bytes */

/* ---- accumulation target -------------------------------------------------- */
/* LG_BGRA8_GAMMA is the reference's screenshot conversion (src/renderer.rs:313-328):
 * every Rgba16Float channel -> (f.powf(1/2.2) * 255) as u8, stored [b, g, r, a].
 * LG_BGRA8_SRGB is the 8-bit target of the path with `render_to_texture` off
 * (the app's default, src/light_garden/mod.rs:86): the pipeline then draws
 * into the surface format (src/sub_render_pass.rs:59-63) and the screenshot
 * into Bgra8UnormSrgb (src/renderer.rs:207-209).  Every channel saturates at
 * 1; colour is stored sRGB-encoded, alpha linear, both rounded to nearest
 * (ORACLE.md 8.7: the order-free limit of the ROP's saturating blend). */
enum { LG_RGBA32F = 0, LG_RGBA16F = 1, LG_BGRA8_GAMMA = 2, LG_BGRA8_SRGB = 3 };

typedef struct LgTraceStats {
  uint64_t primary_rays;    /* rays this context traced (its shard)          */
  uint64_t ray_steps;       /* popped rays that passed the cutoff test       */
  uint64_t object_tests;    /* ray_steps * n_objects (brute force)           */
  uint64_t segments;        /* segments emitted                              */
  uint64_t pixel_updates;   /* fragments blended by accumulate               */
  float trace_ms;           /* CUDA-event time of the trace kernels          */
  float accumulate_ms;      /* CUDA-event time of the accumulate kernels     */
  uint32_t trace_launches;  /* kernels launched by the last trace/render     */
  uint32_t accumulate_launches;
} LgTraceStats; /* 56 bytes */

typedef struct lg_ctx lg_ctx;

/* ---- lifecycle ------------------------------------------------------------ */
int32_t lg_abi_version(void);
int32_t lg_device_count(int32_t *count);
int32_t lg_create(int32_t device, int32_t precision, lg_ctx **out);
int32_t lg_destroy(lg_ctx *ctx);
const char *lg_last_error(const lg_ctx *ctx);

/* ---- B1: scene + trace ---------------------------------------------------- */
/* Replaces the state Tracer::trace_all reads: objects + params
 * (tracer.rs:4-17). Lowers every Geo tree to world-space primitives. */
int32_t lg_scene_set(lg_ctx *ctx, const LgObject *objects, uint32_t n_objects,
                     const LgGeoNode *nodes, uint32_t n_nodes,
                     const LgTraceParams *params);
/* Replaces Tracer.lights; the per-light start medium of tracer.rs:280-287 is
 * evaluated here. */
int32_t lg_lights_set(lg_ctx *ctx, const LgLight *lights, uint32_t n_lights);
/* Tracer::add_drawing_object / finish_drawing_object (tracer.rs:61-72): the object the user is dragging out.  It
 * is not traced against (tracer.rs:412-424 walks self.objects only) but the start-medium scan chains it behind the
 * scene's objects (tracer.rs:281), so a light inside it starts in its material.  object = NULL clears it
 * (finish_drawing_object: the host then pushes the object into the scene it sends with lg_scene_set, or drops it).
 * `object->root` indexes `nodes`.  The reference's drawing_light needs no entry point: it is chained behind
 * self.lights (tracer.rs:279), i.e. the host appends it to the array it passes to lg_lights_set. */
int32_t lg_drawing_object_set(lg_ctx *ctx, const LgObject *object, const LgGeoNode *nodes, uint32_t n_nodes);
/* Data-parallel shard of the primary rays: rank r of `world` takes the rays
 * r, r + world, r + 2*world, ... of every light (interleaved, so that every rank
 * sees every direction of every light and the ranks' work is balanced;
 * SURVEY.md §8e). Default 0 of 1. */
int32_t lg_shard_set(lg_ctx *ctx, uint32_t rank, uint32_t world);
/* Capacity of the device segment buffer in segments (default 64 Mi). */
int32_t lg_segment_capacity_set(lg_ctx *ctx, uint64_t n_segments);
/* Where the additive blend is resolved: 0 = automatic (by segment count), 1 = direct
 * (one red.global.add.v4.f32 per fragment), 2 = tile-binned (fragments are summed in
 * shared-memory tiles first). Same image either way up to fp32 summation order. */
int32_t lg_accumulate_mode_set(lg_ctx *ctx, int32_t mode);
/* Tracer::enable_tile_map (tracer.rs:137-146): when enabled, the nearest-hit
 * search of tracer.rs:395-411 walks a device-side uniform grid over the objects'
 * bounding circles instead of testing every object (the reference's TileMap,
 * tile_map.rs, is the same idea with angular slabs per tile). The segments are
 * bit-identical to the all-objects loop; only the number of exact tests changes.
 * Default 0: every object is tested (tracer.rs:412-424). */
int32_t lg_tile_map_enable(lg_ctx *ctx, int32_t enable);
/* Keep LgSegmentTag (and LgSegmentF64 on F64 contexts) per segment. */
int32_t lg_tags_enable(lg_ctx *ctx, int32_t enable);

/* Ray emission alone: Light::set_num_rays (light.rs:103-115,163-174,225-249)
 * for rays [first, first+count) of light `light`. */
int32_t lg_emit_rays(lg_ctx *ctx, uint32_t light, uint64_t first,
                     uint64_t count, LgRay *dst);

/* Tracer::trace_all (tracer.rs:276-330) for this context's shard: emits the
 * rays on the device and traces them; segments stay on the device.
 * LG_ERR_OVERFLOW if they do not fit the segment buffer (use lg_render). */
int32_t lg_trace(lg_ctx *ctx, LgTraceStats *stats);
/* Tracer::trace (tracer.rs:360-493) on caller supplied primary rays. */
int32_t lg_trace_rays(lg_ctx *ctx, const LgRay *rays, uint64_t n_rays,
                      LgTraceStats *stats);
int32_t lg_segments_count(lg_ctx *ctx, uint64_t *n);
/* Copies up to `cap` segments out, in device (unordered) order. tags/f64 may
 * be NULL. */
int32_t lg_segments_read(lg_ctx *ctx, LgSegment *dst, LgSegmentTag *tags,
                         LgSegmentF64 *f64, uint64_t cap, uint64_t *n);

/* ---- B2: accumulation ----------------------------------------------------- */
/* Rgba16Float render target of width x height (texture_renderer.rs:69-80);
 * the working image is RGBA fp32. */
int32_t lg_image_configure(lg_ctx *ctx, uint32_t width, uint32_t height);
/* LoadOp::Clear(BLACK) (renderer.rs:174-177): rgb = 0, alpha =
 * clear_alpha (1 on the rank that owns the clear, 0 on the others). */
int32_t lg_image_clear(lg_ctx *ctx, float clear_alpha);
/* LightGarden.color_state_descriptor.blend (mod.rs:57-73, edited by gui/settings.rs:59-127); NULL
 * restores the default (rgb: One/One/Add, alpha: SrcAlpha/One/Add). Supported are the states whose
 * result does not depend on the order of the fragments, which is what a parallel line pass can
 * reproduce: Add or ReverseSubtract with dst_factor One and a source-only src_factor (Zero, One,
 * Src, OneMinusSrc, SrcAlpha, OneMinusSrcAlpha, Constant, OneMinusConstant), and Min / Max (wgpu
 * ignores their factors). Anything else is LG_ERR_UNSUPPORTED, here and never silently later.
 * Non-default states use the direct resolve; Min / Max images cannot be summed by lg_image_reduce. */
int32_t lg_blend_set(lg_ctx *ctx, const LgBlendState *state);
/* SubRenderPass::render for the device segment buffer. */
int32_t lg_accumulate_traced(lg_ctx *ctx, LgTraceStats *stats);
/* update_vertex_buffer + render for host lines. */
int32_t lg_accumulate_segments(lg_ctx *ctx, const LgVertexPair *pairs,
                               uint64_t n, LgTraceStats *stats);
/* StringMod::draw (string_mod.rs:152-158) + render; chords [first, first+count)
 * of the pattern (count = 0 means this context's shard of all chords). */
int32_t lg_string_mod(lg_ctx *ctx, const LgStringMod *sm,
                      const LgModRemColor *rules, uint32_t n_rules,
                      uint64_t first, uint64_t count, LgTraceStats *stats);
/* StringMod::draw with `nested: Some(inner)` (string_mod.rs:87-101,152-158): the chords of
 * `outer` are intersected pairwise in the reference's order (diff = 1..L-1, ixa = 0..L-1,
 * partner (ixa + diff) % L; ORACLE.md 7.2 defines the segment intersection), the crossing
 * points become the point set of `inner`, whose chords are then drawn with its own colours. */
int32_t lg_string_mod_nested(lg_ctx *ctx, const LgStringMod *outer, const LgStringMod *inner,
                             const LgModRemColor *inner_rules, uint32_t n_inner_rules,
                             LgTraceStats *stats);
/* The outer chords (f64 end points) and the crossing points (x, y pairs) of the last
 * lg_string_mod_nested call; either destination may be NULL. */
int32_t lg_string_mod_nested_read(lg_ctx *ctx, LgVertexPair *outer_chords, uint64_t chord_cap,
                                  double *crossings_xy, uint64_t crossing_cap,
                                  uint64_t *n_chords, uint64_t *n_crossings);
/* trace_all + render fused: rays are traced in waves through the bounded
 * segment buffer and each wave is accumulated before the next is traced.
 * lg_render_overlap_set can turn the waves into a two-stage pipeline instead
 * (tile-binned resolve only): the line pass of wave k runs on a second stream
 * while wave k + 1 is being traced into the other half of the buffer, with no
 * host round trip in between.  Same fragments either way.  trace_ms / accumulate_ms are the sums of the
 * waves' kernel times: in the pipeline they overlap and add up to more than
 * the frame. */
int32_t lg_render(lg_ctx *ctx, LgTraceStats *stats);
/* The wave pipeline of lg_render: mode 0 = off (default: on B200 the two kernels
 * compete for the same issue slots and shared memory, and the trace kernel needs
 * its full occupancy -- measured 73 ms sequential against 83-95 ms pipelined per
 * 16 M rays of C5), 1 = automatic (frames of >= 2^20 rays whose
 * resolve is the tile bins), 2 = always; waves = how many waves a frame is cut
 * into (0 keeps the current value, default 8). */
int32_t lg_render_overlap_set(lg_ctx *ctx, int32_t mode, uint32_t waves);
/* Image out: LG_RGBA32F (16 B/px), LG_RGBA16F (8 B/px, round to nearest even,
 * what the ROP would have stored), LG_BGRA8_GAMMA (4 B/px, the screenshot
 * path) or LG_BGRA8_SRGB (4 B/px, the 8-bit surface target). pitch in
 * bytes, 0 = tight (wgpu's readback pads rows to 256 bytes,
 * renderer.rs:250-255: pass that pitch to get the same layout). */
int32_t lg_image_read(lg_ctx *ctx, int32_t format, void *dst, size_t pitch);
/* Display hand-off without the host bounce (replaces the
 * queue.write_texture of the frame, src/renderer.rs:356-417 keeps its blit):
 * the frame in `format` lives in device memory that another API can import.
 * lg_image_export_fd returns a POSIX file descriptor for it (an OPAQUE_FD
 * handle: Vulkan VkImportMemoryFdInfoKHR, which wgpu-hal exposes, or
 * cuMemImportFromShareableHandle) and the size of the allocation (the tight
 * W x H frame at offset 0, rounded up to the allocation granularity); every
 * call returns a new descriptor of the same memory and the caller closes it.
 * lg_image_export_refresh converts the current image into that memory and
 * returns when the stream has drained, so the importer may read right after.
 * The memory stays valid until lg_image_configure changes the size or the
 * context is destroyed.  LG_ERR_UNSUPPORTED when the driver cannot export. */
int32_t lg_image_export_fd(lg_ctx *ctx, int32_t format, int32_t *fd, uint64_t *bytes);
int32_t lg_image_export_refresh(lg_ctx *ctx, int32_t format);
/* What a consumer does, for tests and for CUDA-side consumers: imports `fd`
 * (allocation size `bytes`) on `device` and copies its first dst_bytes out. */
int32_t lg_import_fd_read(int32_t device, int32_t fd, uint64_t bytes, void *dst, uint64_t dst_bytes);

/* ---- multi-GPU: one context per device ------------------------------------ */
/* 128-byte ncclUniqueId produced on one rank and handed to all of them. */
int32_t lg_comm_unique_id(void *id128);
int32_t lg_comm_init_rank(lg_ctx *ctx, const void *id128, int32_t rank,
                          int32_t world);
/* Single process driving several devices (the Rust app's shape). */
int32_t lg_comm_init_all(lg_ctx **ctxs, int32_t n);
/* Sum of the partial fp32 images onto `root` (collective: every rank calls it).
 * Default: one kernel per rank over NVLink peer memory that sums its band of rows
 * of every rank's image in rank order and stores the fp32 sum and the Rgba16Float
 * frame into the root's buffers (reduce-scatter + finalize + gather fused); falls
 * back to ncclReduce when a peer cannot be mapped. */
int32_t lg_image_reduce(lg_ctx *ctx, int32_t root, float *reduce_ms);
/* 0 = automatic, 1 = ncclReduce, 2 = peer-memory fused (error if unreachable). */
int32_t lg_reduce_mode_set(lg_ctx *ctx, int32_t mode);
int32_t lg_comm_destroy(lg_ctx *ctx);

/* ---- plumbing for hosts that own the process (bench, tests) --------------- */
/* cudaStream_t the context launches on, as an integer handle. */
int32_t lg_stream_handle(lg_ctx *ctx, uint64_t *stream);
/* Device pointer of the fp32 RGBA working image (width*height*16 bytes). */
int32_t lg_image_device_ptr(lg_ctx *ctx, uint64_t *ptr);
/* Kernels launched by this context since creation. */
int32_t lg_launch_count(lg_ctx *ctx, uint64_t *n);
/* Page-locked host memory for frames / ray / segment buffers: lg_image_read and the
 * other copies run at full PCIe speed into it (any host pointer is accepted, pageable
 * ones go through the driver's staging copy). */
int32_t lg_host_alloc(size_t bytes, void **out);
int32_t lg_host_free(void *p);

/* ---- measurement: the roofline denominators BASELINE.md leaves to the builder -- */
/* FP32 (precision F32) or FP64 (F64) FMA throughput of the device in TFLOP/s,
 * best of `reps` launches of a register-only FMA kernel. */
int32_t lg_measure_fma_peak(lg_ctx *ctx, int32_t precision, int32_t reps, double *tflops);
/* red.global.add.v4.f32 throughput in 1e9 reductions/s over an image of
 * `span_px` RGBA fp32 pixels; pattern 0 = coalesced sweep, 1 = random pixels. */
int32_t lg_measure_red_peak(lg_ctx *ctx, uint64_t span_px, int32_t pattern, int32_t reps, double *gred_per_s);
/* Ceiling of the tile-binned resolve: 16-byte shared-memory read-modify-writes (LDS.128, 4 FADD, STS.128 on a
 * private 32x32 RGBA tile per warp, every lane active, conflict-free, the raster kernel's launch shape) in 1e9
 * fragments/s. */
int32_t lg_measure_tile_rmw_peak(lg_ctx *ctx, int32_t reps, double *gfrag_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LIGHT_GARDEN_B200_H */
