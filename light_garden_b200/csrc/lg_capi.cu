// lg_capi.cu — implementation of include/light_garden_b200.h.
//
// One lg_ctx = one device, one stream, one (optional) NCCL communicator.
// See the header for the reference call sites each entry point replaces.
// There is no CPU fallback anywhere in this file: every compute entry point
// launches a kernel from lg_trace.cuh / lg_accum.cuh or fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/light_garden_b200.h"
#include "lg_accum.cuh"
#include "lg_bench.cuh"
#include "lg_tiles.cuh"
#include "lg_nested.cuh"
#include "lg_reduce.cuh"
#include <unistd.h>
#include "lg_scene.h"
#include "lg_tables.h"
#include "lg_trace.cuh"

static_assert(sizeof(LgGeoNode) == 112, "LgGeoNode ABI");
static_assert(sizeof(LgObject) == 16, "LgObject ABI");
static_assert(sizeof(LgTraceParams) == 56, "LgTraceParams ABI");
static_assert(sizeof(LgLight) == 96, "LgLight ABI");
static_assert(sizeof(LgRay) == 56, "LgRay ABI");
static_assert(sizeof(LgSegment) == 32, "LgSegment ABI");
static_assert(sizeof(LgVertexPair) == 64, "LgVertexPair ABI");
static_assert(sizeof(LgSegmentTag) == 24, "LgSegmentTag ABI");
static_assert(sizeof(LgSegmentF64) == 32, "LgSegmentF64 ABI");
static_assert(sizeof(LgModRemColor) == 32, "LgModRemColor ABI");
static_assert(sizeof(LgStringMod) == 80, "LgStringMod ABI");
static_assert(sizeof(LgTraceStats) == 56, "LgTraceStats ABI");

using namespace lg;

// ---- NCCL through dlopen: no link-time dependency, the process-wide libnccl.so.2
// (torch's or the system's) is the one that gets used -----------------------------------
namespace {
typedef struct {
  char internal[128];
} NcclUniqueId;
typedef void *NcclComm;
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(NcclComm *, int, const int *) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Reduce)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool tried = false;
  std::string why;
};
NcclApi g_nccl;
bool nccl_load() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    g_nccl.why = "libnccl.so.2 not found";
    return false;
  }
  bool ok = true;
  auto sym = [&](const char *s) {
    void *p = dlsym(g_nccl.handle, s);
    if (!p) {
      ok = false;
      g_nccl.why = std::string("missing symbol ") + s;
    }
    return p;
  };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.Reduce = (decltype(g_nccl.Reduce))sym("ncclReduce");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!ok) {
    dlclose(g_nccl.handle);
    g_nccl.handle = nullptr;
  }
  return ok;
}
constexpr int kNcclFloat32 = 7; // ncclFloat32
constexpr int kNcclSum = 0;     // ncclSum
constexpr int kNcclMax = 2;     // ncclMax
constexpr int kNcclInt32 = 2;   // ncclInt32
constexpr int kNcclUint8 = 1;   // ncclUint8
constexpr int kMaxPeers = 16;
} // namespace

// ---- exportable frames (lg_image_export_fd): the driver's virtual-memory API, looked up through the runtime
// (cudaGetDriverEntryPoint), so the library keeps no link-time dependency on libcuda
namespace {
struct VmmApi {
  CUresult (*GetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*AddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*Export)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*Import)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  bool tried = false, ok = false;
  std::string why;
};
VmmApi g_vmm;
bool vmm_load() {
  if (g_vmm.tried) return g_vmm.ok;
  g_vmm.tried = true;
  bool ok = true;
  auto sym = [&](const char *name) -> void * {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      ok = false;
      g_vmm.why = std::string("driver entry point ") + name + " not available";
      return nullptr;
    }
    return fn;
  };
  g_vmm.GetGranularity = (decltype(g_vmm.GetGranularity))sym("cuMemGetAllocationGranularity");
  g_vmm.Create = (decltype(g_vmm.Create))sym("cuMemCreate");
  g_vmm.Release = (decltype(g_vmm.Release))sym("cuMemRelease");
  g_vmm.AddressReserve = (decltype(g_vmm.AddressReserve))sym("cuMemAddressReserve");
  g_vmm.AddressFree = (decltype(g_vmm.AddressFree))sym("cuMemAddressFree");
  g_vmm.Map = (decltype(g_vmm.Map))sym("cuMemMap");
  g_vmm.Unmap = (decltype(g_vmm.Unmap))sym("cuMemUnmap");
  g_vmm.SetAccess = (decltype(g_vmm.SetAccess))sym("cuMemSetAccess");
  g_vmm.Export = (decltype(g_vmm.Export))sym("cuMemExportToShareableHandle");
  g_vmm.Import = (decltype(g_vmm.Import))sym("cuMemImportFromShareableHandle");
  g_vmm.ok = ok;
  return ok;
}
// one frame in memory another API can import (POSIX file descriptor handle)
struct ExportBuf {
  CUmemGenericAllocationHandle handle = 0;
  CUdeviceptr va = 0;
  size_t size = 0; // allocation size (the frame rounded up to the allocation granularity)
};
void export_release(ExportBuf &e) {
  if (e.va && g_vmm.ok) {
    g_vmm.Unmap(e.va, e.size);
    g_vmm.AddressFree(e.va, e.size);
  }
  if (e.handle && g_vmm.ok) g_vmm.Release(e.handle);
  e = ExportBuf{};
}
constexpr int kExportFormats = 4;
} // namespace

// ---- context -----------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct lg_ctx {
  int device = 0;
  int precision = LG_PRECISION_F32;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // lg_render's wave pipeline: the accumulate passes of wave k run on stream2 while wave k + 1 is traced on `stream`
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_pipe = nullptr, ev_t0[2] = {nullptr, nullptr}, ev_t1[2] = {nullptr, nullptr}, ev_a0[2] = {nullptr, nullptr},
              ev_a1[2] = {nullptr, nullptr};
  int render_overlap = 0;   // lg_render_overlap_set: 0 = one wave after the other (default: measured faster), 1 = automatic, 2 = always
  unsigned pipe_waves = 8;  // waves a frame is cut into when the pipeline runs
  int pipe_trace_ctas = -1; // resident trace CTAs per SM while it runs: < 0 = that many fewer than would fit (LG_PIPE_TRACE_CTAS)
  int pipe_raster_ctas = -1; // raster CTAs per SM next to a running trace kernel: < 0 = what its shared memory leaves (LG_PIPE_RASTER_CTAS)
  int raster_cap_now = 0;    // > 0: accumulate_tiled launches at most that many raster CTAs per SM
  size_t smem_per_sm = 0;
  size_t last_trace_smem = 0; // dynamic shared memory and resident CTAs per SM of the last trace launch
  int last_trace_ctas = 0;
  std::string err;
  int sm_count = 0;
  int slots = 2; // ray slots per thread (R)
  size_t smem_optin = 0;

  // scene
  bool have_scene = false;
  HostScene hs;
  bool have_drawing = false; // lg_drawing_object_set: chained behind the objects in the start-medium scan only
  HostScene drawing_hs;
  int n_obj = 0, n_pad = 0;
  bool scene_flat = false; // object i == token i (SceneArgs::flat)
  unsigned bounds_bytes = 0;
  DevBuf bounds, toks, obj_first, obj_count, obj_n, ovl_start, ovl_list;
  double canvas[8] = {0};
  std::vector<double> bound_c; // per object: bounding circle cx, cy, radius (f64)
  double coord_bound = 0;      // max |coordinate| the broad-phase margin was built for
  double delta = 0;

  // lights / shard
  std::vector<LgLight> lights;
  bool have_lights = false; // lg_lights_set was called (possibly with zero lights)
  std::vector<DevLight> dev_lights;
  DevBuf d_lights;
  uint32_t rank = 0, world = 1;
  unsigned long long shard_rays = 0;

  // segments
  unsigned long long seg_cap = 64ull << 20;
  bool tags_on = false;
  DevBuf seg, tags, seg64;
  unsigned long long seg_count = 0;
  DevBuf ctr;
  DevBuf stack;
  DevBuf rays; // explicit primary rays
  double segs_per_ray_est = 0;

  // image
  int W = 0, H = 0;
  DevBuf img, img16, img8, pixctr;
  ExportBuf exported[kExportFormats]; // lg_image_export_fd, indexed by format
  // tile-binned accumulation (lg_tiles.cuh)
  int accum_mode = 0; // 0 = auto, 1 = direct (one L2 reduction per fragment), 2 = tile-binned
  // auto mode: measured cost of each resolve on this context's recent work (ns per fragment, 0 = no sample yet)
  // uniform grid (lg_tile_map_enable)
  bool blend_generic = false; // lg_blend_set: a non-default (order-independent) blend state
  bool blend_linear = true;   // the image is a sum of fragment terms (Add / ReverseSubtract): partial images can be summed
  BlendCfg blend{};
  int fill_wide = 0; // LG_FILL_WIDE=1: always the 64-bit form of tile_fill (tests; otherwise only lists of 2^32 entries and more)
  int bin_grid = 0; // LG_BIN_GRID: CTAs of the count / fill passes, absolute (0 = sm_count x bin_ctas)
  int bin_ctas = 3, bin_threads = 1024; // count / fill passes (measured: 19.1 ms against 20.7 ms with 4 x 256): CTAs per SM and threads per CTA (LG_BIN_CTAS, LG_BIN_THREADS)
  int trace_merged = -1; // -1 = by scene size
  bool grid_on = false;
  double grid_density = 1.0; // cells per object (LG_GRID_DENSITY)
  int grid_slots = 1;        // ray slots per thread of the grid kernel (LG_GRID_SLOTS)
  DevBuf grid_start, grid_obj;
  // nested string mod: outer chords, per-block crossing counts, crossing points, inner chords
  DevBuf nest_lines, nest_counts, nest_points, nest_pairs;
  unsigned long long nest_lines_n = 0, nest_points_n = 0;
  double grid_x0 = 0, grid_y0 = 0, grid_x1 = 0, grid_y1 = 0, grid_cs = 1, grid_eta = 0;
  int grid_nx = 0, grid_ny = 0;
  // accumulate auto mode: measured cost of the two resolves per WORKLOAD (scene + lights + image for traced
  // segments, chord pattern for string mod, segment-count class for host lines)
  struct AutoStat {
    double ns_per_frag[3] = {0, 0, 0};
    unsigned samples[3] = {0, 0, 0};
    unsigned long long calls = 0;
  };
  std::unordered_map<unsigned long long, AutoStat> auto_stats;
  unsigned long long scene_hash = 0, lights_hash = 0;
  DevBuf tile_count, tile_cursor, tile_offset, item_prefix, tile_totals, item_counter, tile_list, seg2, tile_hist;
  double pairs_per_seg_est = 0;              // (segment, tile) pairs per segment of the last tiled resolve: sizes the next pair list
  TileArgs tiled_last{};                     // arguments of the last accumulate_tiled (tiled_finish runs fill + raster again from them)
  int tiled_raster_grid = 0;
  size_t tiled_raster_smem = 0;
  unsigned long long *h_totals = nullptr;    // page-locked: [0..3] totals of the last tiled resolve, [4..] wave status (lg_render)

  // comm
  NcclComm comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  // peer-memory reduce (lg_reduce.cuh): every rank's fp32 image and the root's fp16 frame, mapped here
  int reduce_mode = 0; // 0 = auto (peer-fused when every peer is reachable), 1 = ncclReduce, 2 = peer-fused only
  bool peers_ready = false;
  int peers_root = -1;
  void *peer_img[kMaxPeers] = {nullptr};
  void *peer_img16[kMaxPeers] = {nullptr};
  bool peer_opened[kMaxPeers] = {false};   // mapped through cudaIpcOpenMemHandle (another process)
  bool peer_opened16[kMaxPeers] = {false};
  bool peer_fused_ok = true;
  bool peer_size_mismatch = false; // the last handle exchange found ranks with images of different sizes
  DevBuf sync_buf, peer_xchg;
  // device-side barriers of the fused reduce (lg_reduce.cuh: peer_barrier_kernel): this rank's flag words, every
  // peer's mapped here; they outlive lg_image_configure (only the images move)
  DevBuf flags, flag_status;
  void *peer_flags[kMaxPeers] = {nullptr};
  bool peer_flags_opened[kMaxPeers] = {false};
  bool flags_ready = false;
  unsigned long long flag_epoch = 0;
  bool img16_valid = false; // the fp16 frame already holds the finalized image

  unsigned long long launches = 0;
};

namespace {

void close_peers(lg_ctx *c);
void close_peer_flags(lg_ctx *c);

int fail(lg_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  return code;
}
#define LG_CUDA(c, call)                                                                                  \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail((c), e_ == cudaErrorMemoryAllocation ? LG_ERR_NOMEM : LG_ERR_CUDA,                      \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                    \
  } while (0)

int ensure(lg_ctx *c, DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return LG_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(c, LG_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  }
  b.bytes = bytes;
  return LG_OK;
}
void release(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}
template <class V> int upload(lg_ctx *c, DevBuf &b, const std::vector<V> &v) {
  int rc = ensure(c, b, v.size() * sizeof(V));
  if (rc) return rc;
  if (!v.empty()) LG_CUDA(c, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice, c->stream));
  return LG_OK;
}

// device tables in precision T (built by lg_tables.h, uploaded here)
template <class T> int upload_scene(lg_ctx *c) {
  const HostScene &hs = c->hs;
  const std::vector<Tok<T>> toks = device_tokens<T>(hs);
  std::vector<int> obj_first, obj_count;
  std::vector<T> obj_n;
  for (size_t i = 0; i < hs.objs.size(); ++i) {
    const HostObj &o = hs.objs[i];
    obj_first.push_back(o.first);
    obj_count.push_back(o.count);
    obj_n.push_back(o.has_material ? (T)o.n : std::numeric_limits<T>::quiet_NaN());
  }
  c->n_obj = (int)hs.objs.size();
  c->scene_flat = true;
  for (size_t i = 0; i < hs.objs.size(); ++i)
    if (hs.objs[i].first != (int)i || hs.objs[i].count != 1) c->scene_flat = false;
  int rc;
  if ((rc = upload(c, c->toks, toks))) return rc;
  if ((rc = upload(c, c->obj_first, obj_first))) return rc;
  if ((rc = upload(c, c->obj_count, obj_count))) return rc;
  if ((rc = upload(c, c->obj_n, obj_n))) return rc;
  if ((rc = upload(c, c->ovl_start, hs.ovl_start))) return rc;
  if ((rc = upload(c, c->ovl_list, hs.ovl_list))) return rc;
  // canvas as a rect: Rect::from_tlbr(top, left, bottom, right), sub_render_pass.rs:156
  const double top = hs.params.canvas_tlbr[0], left = hs.params.canvas_tlbr[1], bottom = hs.params.canvas_tlbr[2],
               right = hs.params.canvas_tlbr[3];
  T cv[8];
  cv[0] = (T)((left + right) * 0.5);
  cv[1] = (T)((top + bottom) * 0.5);
  cv[2] = (T)((right - left) * 0.5);
  cv[3] = (T)0;
  cv[4] = (T)0;
  cv[5] = (T)((top - bottom) * 0.5);
  cv[6] = cv[2] * cv[2];
  cv[7] = cv[5] * cv[5];
  for (int k = 0; k < 8; ++k) c->canvas[k] = (double)cv[k];
  c->bound_c = object_circles(hs);
  c->coord_bound = 0; // forces the broad-phase table to be rebuilt at the next launch
  LG_CUDA(c, cudaStreamSynchronize(c->stream)); // host vectors die here
  return LG_OK;
}

// Broad-phase table and uniform grid for coordinate bound B (scene, canvas, lights, explicit ray origins)
template <class T> int upload_bounds(lg_ctx *c, double B) {
  const size_t n = c->bound_c.size() / 3;
  const BoundsTable<T> bt = build_bounds<T>(c->bound_c.data(), n, B);
  const HostGrid g = build_scene_grid<T>(bt, n, c->grid_density);
  c->grid_x0 = g.x0, c->grid_y0 = g.y0, c->grid_x1 = g.x1, c->grid_y1 = g.y1, c->grid_cs = g.cs;
  c->grid_nx = g.nx, c->grid_ny = g.ny, c->grid_eta = bt.delta;
  int rc;
  if ((rc = upload(c, c->grid_start, g.start))) return rc;
  if ((rc = upload(c, c->grid_obj, g.obj))) return rc;
  if (g.obj.empty() && (rc = ensure(c, c->grid_obj, 4))) return rc;
  c->n_pad = bt.n_pad;
  c->bounds_bytes = (unsigned)(bt.tab.size() * sizeof(T));
  c->coord_bound = B;
  c->delta = bt.delta;
  if ((rc = upload(c, c->bounds, bt.tab))) return rc;
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}
int ensure_bounds(lg_ctx *c, double ray_bound) {
  double B = std::fmax(c->hs.bound, ray_bound);
  for (const LgLight &l : c->lights) {
    B = std::fmax(B, std::fmax(std::fabs(l.position[0]), std::fabs(l.position[1])));
    if (l.kind == LG_LIGHT_DIRECTIONAL) B = std::fmax(B, std::fmax(std::fabs(l.b[0]), std::fabs(l.b[1])));
  }
  if (c->bounds.p && c->coord_bound >= B && c->coord_bound > 0) return LG_OK;
  return c->precision == LG_PRECISION_F64 ? upload_bounds<double>(c, B) : upload_bounds<float>(c, B);
}

void rebuild_dev_lights(lg_ctx *c) {
  c->dev_lights.clear();
  unsigned long long prefix = 0, id_base = 0;
  for (const LgLight &l : c->lights) {
    DevLight d{};
    d.kind = l.kind;
    const unsigned long long n = l.num_rays;
    // interleaved shard: ray i belongs to rank i mod world (every rank sees every direction of every light, so
    // the ranks' work is balanced whatever the scene looks like around the light)
    d.first = c->rank;
    d.stride = c->world;
    d.count = n > c->rank ? (n - c->rank + c->world - 1) / c->world : 0;
    d.prefix = prefix;
    d.id_base = id_base;
    d.n_rays = (double)n;
    d.n0 = c->have_scene ? host_start_medium(c->hs, l.position[0], l.position[1]) : 1.0;
    // `.chain(&self.drawing_object)` (tracer.rs:281): last in the chain, so it wins when it contains the light
    if (c->have_drawing && c->drawing_hs.objs.size() == 1 && c->drawing_hs.objs[0].has_material &&
        host_contains(c->drawing_hs, 0, l.position[0], l.position[1]))
      d.n0 = c->drawing_hs.objs[0].n;
    if (l.flags & LG_LIGHT_START_MEDIUM) d.n0 = l.start_medium;
    d.rsign = (l.flags & LG_LIGHT_DIRECTIONAL_NEG_R) ? -1.0 : 1.0;
    std::memcpy(d.color, l.color, 16);
    d.px = l.position[0], d.py = l.position[1];
    d.ex = l.b[0] - l.position[0], d.ey = l.b[1] - l.position[1];
    if (l.kind == LG_LIGHT_SPOT) { // light.rs:230-243; EPSILON: ORACLE.md §6.2
      const double PI = 3.14159265358979323846;
      const double dx = l.spot_direction[0], dy = l.spot_direction[1];
      const double da = std::fabs(dx) < 1e-10 ? (dy >= 0. ? PI * 0.5 : -PI * 0.5) : std::atan(dy / dx);
      d.min_angle = da - 0.5 * l.spot_angle;
      d.spot_angle = l.spot_angle;
      d.sign = std::isnan(dx) ? dx : (std::signbit(dx) ? -1.0 : 1.0); // f64::signum
    }
    c->dev_lights.push_back(d);
    prefix += d.count;
    id_base += n;
  }
  c->shard_rays = prefix;
}

template <class T> void fill_args(lg_ctx *c, TraceArgs<T> &A) {
  A.bounds = (const T *)c->bounds.p;
  A.bounds_bytes = c->bounds_bytes;
  A.n_pad = c->n_pad;
  A.delta = (T)c->delta;
  A.toks = (const Tok<T> *)c->toks.p;
  A.obj_first = (const int *)c->obj_first.p, A.obj_count = (const int *)c->obj_count.p;
  A.flat = c->scene_flat ? 1 : 0;
  A.obj_n = (const T *)c->obj_n.p;
  A.ovl_start = (const int *)c->ovl_start.p, A.ovl_list = (const int *)c->ovl_list.p;
  for (int k = 0; k < 8; ++k) A.canvas[k] = (T)c->canvas[k];
  A.n_obj = c->n_obj;
  A.max_bounce = c->hs.params.max_bounce;
  std::memcpy(A.cutoff, c->hs.params.cutoff_color, 16);
  A.lights = (const DevLight *)c->d_lights.p;
  A.n_lights = (int)c->dev_lights.size();
  A.rays = nullptr;
  A.seg = (LgSegment *)c->seg.p;
  A.tags = c->tags_on ? (LgSegmentTag *)c->tags.p : nullptr;
  A.seg64 = (c->tags_on && c->precision == LG_PRECISION_F64) ? (LgSegmentF64 *)c->seg64.p : nullptr;
  A.seg_cap = c->seg_cap;
  A.ctr = (TraceCounters *)c->ctr.p;
  A.stack = (uint4 *)c->stack.p;
  A.grid_x0 = (T)c->grid_x0, A.grid_y0 = (T)c->grid_y0, A.grid_x1 = (T)c->grid_x1, A.grid_y1 = (T)c->grid_y1;
  A.grid_cs = (T)c->grid_cs, A.grid_ics = (T)(1.0 / c->grid_cs), A.grid_eta = (T)c->grid_eta;
  A.grid_nx = c->grid_nx, A.grid_ny = c->grid_ny;
  A.grid_start = (const unsigned *)c->grid_start.p, A.grid_obj = (const unsigned *)c->grid_obj.p;
}

template <class T> struct KernelOf;
template <> struct KernelOf<float> {
  static const void *get(int slots, bool smem) { return trace_kernel_f32(slots, smem); }
  static const void *grid(int slots) { return trace_kernel_grid_f32(slots); }
  static int clamp(int slots) { return (slots == 1 || slots == 4) ? slots : 2; }
};
template <> struct KernelOf<double> {
  static const void *get(int slots, bool smem) { return trace_kernel_f64(slots, smem); }
  static const void *grid(int slots) { return trace_kernel_grid_f64(slots); }
  static int clamp(int slots) { return slots == 2 ? 2 : 1; }
};

// st: the stream to launch on; cta_cap > 0 limits the resident CTAs per SM (lg_render's wave pipeline leaves room
// on every SM for the accumulate kernels of the previous wave)
template <class T> int launch_trace(lg_ctx *c, TraceArgs<T> &A, cudaStream_t st, int cta_cap) {
  const bool grid_mode = c->grid_on && c->n_obj > 0;
  int R = KernelOf<T>::clamp(c->slots);
  if (grid_mode) R = (sizeof(T) == 4 && c->grid_slots == 2) ? 2 : 1; // divergent cell walks: one ray per thread by default
  const bool use_smem = !grid_mode && (size_t)A.bounds_bytes + 1024 <= c->smem_optin;
  const void *kern = grid_mode ? KernelOf<T>::grid(R) : KernelOf<T>::get(R, use_smem);
  // large scenes: per-slot narrow phase (see trace_kernel); LG_TRACE_MERGED=0/1 forces one or the other
  const bool per_slot = c->trace_merged < 0 ? c->n_obj >= 512 : c->trace_merged == 0;
  if (!grid_mode && sizeof(T) == 4 && R == 2 && per_slot) kern = trace_kernel_f32_per_slot(use_smem);
  if (!grid_mode && sizeof(T) == 8 && per_slot) kern = trace_kernel_f64_large(R, use_smem);
  const size_t smem = use_smem ? A.bounds_bytes : 0;
  if (smem > 48 * 1024) LG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  LG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTraceBlock, smem));
  if (per_sm < 1) return fail(c, LG_ERR_CUDA, "trace kernel does not fit on an SM");
  if (cta_cap < 0) per_sm = std::max(1, per_sm + cta_cap); // -1: one CTA per SM less than would fit
  else if (cta_cap > 0) per_sm = std::min(per_sm, cta_cap);
  c->last_trace_smem = smem, c->last_trace_ctas = per_sm;
  const int grid = c->sm_count * per_sm;
  const size_t nthreads = (size_t)grid * kTraceBlock;
  // split stack: one private stack per slot, depth <= max_bounce - 1 live branches
  int cap = (int)std::min<unsigned>(A.max_bounce > 0 ? A.max_bounce - 1 : 0, 64u);
  if (cap < 1) cap = 1;
  A.stack_cap = cap;
  int rc = ensure(c, c->stack, (size_t)cap * R * StackCodec<T>::kVecs * sizeof(uint4) * nthreads);
  if (rc) return rc;
  A.stack = (uint4 *)c->stack.p;
  void *args[] = {(void *)&A};
  LG_CUDA(c, cudaLaunchKernel(kern, dim3(grid), dim3(kTraceBlock), args, smem, st));
  c->launches++;
  return LG_OK;
}

// queues one trace launch over [first, end) of the ray index space on `st`: segments into seg[0, seg_cap), counters
// (cleared first) at `ctr`.  ensure_bounds() must have run.
int trace_async(lg_ctx *c, cudaStream_t st, const LgRay *d_rays, unsigned long long first, unsigned long long end,
                LgSegment *seg, unsigned long long seg_cap, TraceCounters *ctr, int cta_cap) {
  LG_CUDA(c, cudaMemsetAsync(ctr, 0, sizeof(TraceCounters), st));
  if (c->precision == LG_PRECISION_F64) {
    TraceArgs<double> A{};
    fill_args(c, A);
    A.rays = d_rays, A.ray_first = first, A.ray_end = end, A.seg = seg, A.seg_cap = seg_cap, A.ctr = ctr;
    return launch_trace(c, A, st, cta_cap);
  }
  TraceArgs<float> A{};
  fill_args(c, A);
  A.rays = d_rays, A.ray_first = first, A.ray_end = end, A.seg = seg, A.seg_cap = seg_cap, A.ctr = ctr;
  return launch_trace(c, A, st, cta_cap);
}

int prepare_trace_buffers(lg_ctx *c) {
  int rc;
  if ((rc = ensure(c, c->seg, c->seg_cap * sizeof(LgSegment)))) return rc;
  if (c->tags_on) {
    if ((rc = ensure(c, c->tags, c->seg_cap * sizeof(LgSegmentTag)))) return rc;
    if (c->precision == LG_PRECISION_F64 && (rc = ensure(c, c->seg64, c->seg_cap * sizeof(LgSegmentF64)))) return rc;
  }
  if ((rc = ensure(c, c->ctr, 2 * sizeof(TraceCounters)))) return rc; // two: the wave pipeline's buffers
  return LG_OK;
}

// one trace launch over [first, end) of the ray index space (or explicit rays);
// returns with the stream drained and the counters on the host
int trace_range(lg_ctx *c, const LgRay *d_rays, unsigned long long first, unsigned long long end, TraceCounters &out,
                float *ms, double ray_bound = 0.0) {
  int rb = ensure_bounds(c, ray_bound);
  if (rb) return rb;
  LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  int rc = trace_async(c, c->stream, d_rays, first, end, (LgSegment *)c->seg.p, c->seg_cap, (TraceCounters *)c->ctr.p, 0);
  if (rc) return rc;
  LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  LG_CUDA(c, cudaMemcpyAsync(&out, c->ctr.p, sizeof(TraceCounters), cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  float t = 0.f;
  LG_CUDA(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
  if (ms) *ms += t;
  if (out.stack_overflow) return fail(c, LG_ERR_OVERFLOW, "split stack overflow (max_bounce with refraction > 65)");
  return LG_OK;
}

AccumArgs accum_args(lg_ctx *c) {
  AccumArgs A;
  A.img = (float *)c->img.p;
  A.W = c->W, A.H = c->H;
  const float aspect = (float)c->W / (float)c->H; // sub_render_pass.rs:146
  A.m00 = 2.0f / (aspect - (-aspect));           // cgmath::ortho, renderer.rs:121
  A.m11 = 2.0f / (1.0f - (-1.0f));
  A.hw = (float)c->W * 0.5f;
  A.hh = (float)c->H * 0.5f;
  A.pixel_updates = (unsigned long long *)c->pixctr.p;
  A.blend = c->blend;
  return A;
}
int accum_grid(lg_ctx *c) { return c->sm_count * 8; }

constexpr unsigned long long kTiledMinSegments = 1ull << 17; // below this the direct kernels win

bool tiled_possible(lg_ctx *c, unsigned long long n) {
  if (n == 0 || n >= (1ull << 32)) return false; // the tile lists hold 32-bit segment indices
  const size_t n_lists = 2 * (size_t)((c->W + kTile - 1) / kTile) * ((c->H + kTile - 1) / kTile);
  return n_lists * 4 + 1024 <= c->smem_optin; // the per-CTA histogram must fit in shared memory
}

// 64-bit content hash (workload signatures of the auto mode; never used to skip work)
unsigned long long hash_bytes(const void *p, size_t n, unsigned long long h = 0x9E3779B97F4A7C15ull) {
  const unsigned char *b = (const unsigned char *)p;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    unsigned long long v;
    std::memcpy(&v, b + i, 8);
    h = (h ^ v) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  for (; i < n; ++i) h = (h ^ b[i]) * 0x100000001B3ull;
  return h;
}
unsigned long long mix_sig(unsigned long long a, unsigned long long b) { return hash_bytes(&b, 8, a); }
unsigned long long traced_signature(lg_ctx *c) {
  unsigned long long h = mix_sig(c->scene_hash, c->lights_hash);
  h = mix_sig(h, ((unsigned long long)c->W << 32) | (unsigned)c->H);
  return mix_sig(h, ((unsigned long long)c->rank << 32) | c->world);
}

// Which resolve to use.  Forced modes aside, small inputs go direct; large ones use whichever of the two has been
// cheaper per fragment on this workload (signature `sig`) so far: each is sampled twice first (the first call of a
// mode pays for its buffers), the loser is re-probed every 64th call.  Long coalescing segments favour the direct
// kernel, short or scattered ones the tile bins.
bool use_tiled(lg_ctx *c, unsigned long long n, unsigned long long sig) {
  if (!tiled_possible(c, n)) return false;
  if (c->blend_generic) return false; // non-default blend states go through the direct kernels (lg_blend_set)
  if (c->accum_mode == 1) return false;
  if (c->accum_mode == 2) return true;
  if (n < kTiledMinSegments) return false;
  if (c->auto_stats.size() > 256) c->auto_stats.clear();
  lg_ctx::AutoStat &a = c->auto_stats[sig];
  ++a.calls;
  if (a.samples[2] < 2) return true;
  if (a.samples[1] < 2) return false;
  const bool tiled_better = a.ns_per_frag[2] <= a.ns_per_frag[1];
  if (a.calls % 64 == 0) return !tiled_better;
  return tiled_better;
}

// feed the auto mode: elapsed time of one resolve over `frags` fragments
void note_accum_cost(lg_ctx *c, bool tiled, float ms, unsigned long long frags, unsigned long long n,
                     unsigned long long sig) {
  if (frags == 0 || n < kTiledMinSegments) return;
  const double v = (double)ms * 1e6 / (double)frags;
  lg_ctx::AutoStat &a = c->auto_stats[sig];
  const int m = tiled ? 2 : 1;
  // the best sample so far: a call that paid for an allocation, or one odd slow call, must not flip the choice
  // (round 2: one driver-style run settled on the direct resolve, 60 ms instead of 33 ms per step, after such a sample)
  a.ns_per_frag[m] = a.samples[m] < 1 ? v : std::min(a.ns_per_frag[m], v);
  if (a.samples[m] < 2) ++a.samples[m];
  if (getenv("LG_DEBUG_AUTO"))
    fprintf(stderr, "[lg auto] %s: %.3f ms, %llu fragments, %llu segments -> %.4f ns/fragment (direct %.4f, tiled %.4f)\n",
            tiled ? "tiled" : "direct", ms, frags, n, v, a.ns_per_frag[1], a.ns_per_frag[2]);
}

// count -> scan -> fill -> raster over device segments of type Seg (LgSegment or Seg2), queued on `st` without a
// host round trip: the pair list is sized from what this context has seen so far (pairs per segment of the previous
// call, with head room); tile_scan_kernel checks the real total against it and, when it does not fit, flags the
// overflow and leaves fill / raster nothing to do.  tiled_finish() -- after the caller's own synchronisation --
// reads the totals, grows the list and runs the two passes again in that case (first call of a workload at most).
// n_dev / skip_dev: the segment count and the "wave overflowed" flag may live on the device (lg_render's waves).
// The fill pass of the tile bins (lg_tiles.cuh: which form of the cursors).  T.n_ctas is the count pass's split of the segments.
template <class Seg> int launch_fill(lg_ctx *c, const TileArgs &T, cudaStream_t st, const Seg *d_seg, unsigned long long n) {
  // positions in the pair list fit 32 bits unless the list itself is larger than that (C2 at full size: 5.7e9 pairs; the
  // list keeps its size afterwards, so a later 4096 x 4096 frame sees it too)
  const bool wide = T.list_cap >= (1ull << 32) || c->fill_wide;
  const bool staged = wide && (size_t)T.n_tiles * 12 + 1024 <= c->smem_optin;
  const size_t smem = (size_t)T.n_tiles * (staged ? 12 : 4);
  auto k = !wide ? tile_fill_kernel<Seg, 0> : (staged ? tile_fill_kernel<Seg, 1> : tile_fill_kernel<Seg, 2>);
  if (smem > 48 * 1024) LG_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<T.n_ctas, c->bin_threads, smem, st>>>(T, d_seg, n);
  LG_CUDA(c, cudaGetLastError());
  return LG_OK;
}

template <class Seg>
int accumulate_tiled(lg_ctx *c, cudaStream_t st, const Seg *d_seg, unsigned long long n, const unsigned long long *n_dev,
                     const unsigned int *skip_dev, unsigned long long *h_totals, unsigned *launches,
                     unsigned long long n_expect = 0) {
  if (n_expect == 0) n_expect = n; // segments the pair list is sized for (n is only a capacity when n_dev is given)
  TileArgs T;
  T.A = accum_args(c);
  T.tiles_x = (c->W + kTile - 1) / kTile, T.tiles_y = (c->H + kTile - 1) / kTile;
  T.n_tiles = 2 * T.tiles_x * T.tiles_y; // two lists per tile: x-major and y-major segments
  int rc;
  if ((rc = ensure(c, c->tile_count, (size_t)T.n_tiles * 4))) return rc;
  if ((rc = ensure(c, c->tile_cursor, (size_t)T.n_tiles * 4))) return rc;
  if ((rc = ensure(c, c->tile_offset, ((size_t)T.n_tiles + 1) * 8))) return rc;
  if ((rc = ensure(c, c->item_prefix, ((size_t)T.n_tiles + 1) * 4))) return rc;
  if ((rc = ensure(c, c->tile_totals, 64))) return rc;
  if ((rc = ensure(c, c->item_counter, 4))) return rc;
  // pair list: pairs per segment seen so far x 1.25 (4 before the first sample), never shrinks
  const double pps = c->pairs_per_seg_est > 0 ? c->pairs_per_seg_est * 1.25 : 4.0;
  unsigned long long want = (unsigned long long)(pps * (double)n_expect) + 4096ull;
  if (want < c->tile_list.bytes / 4) want = c->tile_list.bytes / 4;
  if ((rc = ensure(c, c->tile_list, (size_t)want * 4))) return rc;
  T.tile_count = (unsigned *)c->tile_count.p, T.tile_cursor = (unsigned *)c->tile_cursor.p;
  T.tile_offset = (unsigned long long *)c->tile_offset.p, T.item_prefix = (unsigned *)c->item_prefix.p;
  T.totals = (unsigned long long *)c->tile_totals.p, T.item_counter = (unsigned *)c->item_counter.p;
  T.list = (unsigned *)c->tile_list.p;
  T.list_cap = c->tile_list.bytes / 4;
  T.n_dev = n_dev, T.skip_dev = skip_dev;
  const size_t hist_smem = (size_t)T.n_tiles * 4;
  auto count_k = tile_count_kernel<Seg>;
  if (hist_smem > 48 * 1024) LG_CUDA(c, cudaFuncSetAttribute(count_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
  // CTAs of the count and fill passes (the same split of the segments in both): one resident wave -- the histogram
  // in shared memory decides how many fit on an SM -- and no more than the segments can keep busy
  int hist_per_sm = 0;
  LG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&hist_per_sm, count_k, c->bin_threads, hist_smem));
  if (hist_per_sm < 1) return fail(c, LG_ERR_CUDA, "tile histogram does not fit in shared memory");
  int grid = c->bin_grid > 0 ? c->bin_grid : c->sm_count * c->bin_ctas;
  grid = (int)std::max<unsigned long long>(1ull, std::min<unsigned long long>((unsigned long long)grid, (n + 1023ull) / 1024ull));
  T.n_ctas = grid;
  if ((rc = ensure(c, c->tile_hist, (size_t)grid * T.n_tiles * 4))) return rc;
  T.hist = (unsigned *)c->tile_hist.p;
  const size_t smem = (size_t)kRasterWarps * ((size_t)kTileFloat4 * sizeof(float4) + sizeof(RasterScratch));
  auto kern = tile_raster_kernel<Seg>;
  LG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  LG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRasterWarps * 32, smem));
  if (per_sm < 1) return fail(c, LG_ERR_CUDA, "tile raster kernel does not fit on an SM");
  if (c->raster_cap_now > 0) per_sm = std::min(per_sm, c->raster_cap_now);
  c->tiled_last = T;
  c->tiled_raster_grid = c->sm_count * per_sm;
  c->tiled_raster_smem = smem;
  count_k<<<grid, c->bin_threads, hist_smem, st>>>(T, d_seg, n);
  LG_CUDA(c, cudaGetLastError());
  tile_rowscan_kernel<<<(T.n_tiles + 31) / 32, 32 * kRowscanWarps, 0, st>>>(T);
  LG_CUDA(c, cudaGetLastError());
  tile_scan_kernel<<<1, 1024, 0, st>>>(T);
  LG_CUDA(c, cudaGetLastError());
  if ((rc = launch_fill<Seg>(c, T, st, d_seg, n))) return rc;
  kern<<<c->tiled_raster_grid, kRasterWarps * 32, smem, st>>>(T, d_seg);
  LG_CUDA(c, cudaGetLastError());
  LG_CUDA(c, cudaMemcpyAsync(h_totals, T.totals, 32, cudaMemcpyDeviceToHost, st));
  c->launches += 5;
  if (launches) *launches += 5;
  return LG_OK;
}

// After the stream that ran accumulate_tiled() has been synchronised: the pair list was too small -> grow it and run
// fill + raster again (the counts and offsets of the first three passes are still valid).  Blocking.
template <class Seg>
int tiled_finish(lg_ctx *c, cudaStream_t st, const Seg *d_seg, unsigned long long n, const unsigned long long *h_totals,
                 unsigned *launches) {
  if (n > 0 && h_totals[0] > 0) c->pairs_per_seg_est = std::max(1.0, (double)h_totals[0] / (double)n);
  int rc;
  // the list is grown HERE, outside the caller's timed region, to what the next call will ask for (pairs per segment
  // x 1.25 and head room), whether or not this call overflowed it
  const size_t next = (size_t)((double)h_totals[0] * 1.25 + (double)n + 8192.0) * 4;
  if (!h_totals[2]) {
    if (next > c->tile_list.bytes && (rc = ensure(c, c->tile_list, next))) return rc;
    return LG_OK;
  }
  TileArgs T = c->tiled_last;
  if ((rc = ensure(c, c->tile_list, next))) return rc;
  T.list = (unsigned *)c->tile_list.p;
  T.list_cap = c->tile_list.bytes / 4;
  const unsigned long long fixed[4] = {h_totals[0], h_totals[3], 0ull, h_totals[3]};
  LG_CUDA(c, cudaMemcpyAsync(T.totals, fixed, 32, cudaMemcpyHostToDevice, st));
  LG_CUDA(c, cudaMemsetAsync(T.item_counter, 0, 4, st)); // the raster's first (empty) launch moved it
  if ((rc = launch_fill<Seg>(c, T, st, d_seg, n))) return rc; // (the grown list may have crossed 2^32 entries)
  tile_raster_kernel<Seg><<<c->tiled_raster_grid, kRasterWarps * 32, c->tiled_raster_smem, st>>>(T, d_seg);
  LG_CUDA(c, cudaGetLastError());
  LG_CUDA(c, cudaStreamSynchronize(st));
  c->launches += 2;
  if (launches) *launches += 2;
  return LG_OK;
}

int read_pixel_counter(lg_ctx *c, uint64_t *out);

int accumulate_device_segments(lg_ctx *c, unsigned long long n, float *ms, unsigned *launches) {
  if (n == 0) return LG_OK;
  c->img16_valid = false;
  const unsigned long long sig = traced_signature(c);
  const bool tiled = use_tiled(c, n, sig);
  uint64_t before = 0, after = 0;
  int rc;
  if (c->accum_mode == 0 && n >= kTiledMinSegments && (rc = read_pixel_counter(c, &before))) return rc;
  LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  if (tiled) {
    if ((rc = accumulate_tiled<LgSegment>(c, c->stream, (const LgSegment *)c->seg.p, n, nullptr, nullptr, c->h_totals, launches))) return rc;
  } else {
    if (c->blend_generic)
      accumulate_segments_kernel<true><<<accum_grid(c), kAccumBlock, 0, c->stream>>>(accum_args(c), (const LgSegment *)c->seg.p, n);
    else
      accumulate_segments_kernel<false><<<accum_grid(c), kAccumBlock, 0, c->stream>>>(accum_args(c), (const LgSegment *)c->seg.p, n);
    LG_CUDA(c, cudaGetLastError());
    c->launches++;
    if (launches) (*launches)++;
  }
  LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (tiled && (rc = tiled_finish<LgSegment>(c, c->stream, (const LgSegment *)c->seg.p, n, c->h_totals, launches))) return rc;
  float t = 0.f;
  LG_CUDA(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
  if (ms) *ms += t;
  if (c->accum_mode == 0 && n >= kTiledMinSegments) {
    if ((rc = read_pixel_counter(c, &after))) return rc;
    note_accum_cost(c, tiled, t, after - before, n, sig);
  }
  return LG_OK;
}

int read_pixel_counter(lg_ctx *c, uint64_t *out) {
  unsigned long long v = 0;
  LG_CUDA(c, cudaMemcpyAsync(&v, c->pixctr.p, 8, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = v;
  return LG_OK;
}

int need_scene_lights(lg_ctx *c, bool lights) {
  if (!c) return LG_ERR_INVALID;
  if (!c->have_scene) return fail(c, LG_ERR_STATE, "lg_scene_set has not been called");
  if (lights && !c->have_lights) return fail(c, LG_ERR_STATE, "lg_lights_set has not been called");
  return LG_OK;
}
int need_image(lg_ctx *c) {
  if (!c) return LG_ERR_INVALID;
  if (!c->img.p || c->W <= 0) return fail(c, LG_ERR_STATE, "lg_image_configure has not been called");
  return LG_OK;
}

} // namespace

// ===========================================================================================
extern "C" {

int32_t lg_abi_version(void) { return LG_ABI_VERSION; }

int32_t lg_device_count(int32_t *count) {
  if (!count) return LG_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return LG_ERR_CUDA;
  }
  *count = n;
  return LG_OK;
}

int32_t lg_create(int32_t device, int32_t precision, lg_ctx **out) {
  if (!out) return LG_ERR_INVALID;
  *out = nullptr;
  if (precision != LG_PRECISION_F32 && precision != LG_PRECISION_F64) return LG_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return LG_ERR_CUDA; // no CPU fallback
  }
  if (device < 0 || device >= n) return LG_ERR_INVALID;
  lg_ctx *c = new lg_ctx();
  c->device = device;
  c->precision = precision;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    cudaGetLastError();
    delete c;
    return LG_ERR_CUDA;
  }
  bool ev_ok = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreate(&c->ev_pipe) == cudaSuccess;
  for (int b = 0; b < 2 && ev_ok; ++b)
    ev_ok = cudaEventCreate(&c->ev_t0[b]) == cudaSuccess && cudaEventCreate(&c->ev_t1[b]) == cudaSuccess &&
            cudaEventCreate(&c->ev_a0[b]) == cudaSuccess && cudaEventCreate(&c->ev_a1[b]) == cudaSuccess;
  if (!ev_ok) {
    cudaGetLastError();
    delete c;
    return LG_ERR_CUDA;
  }
  if (cudaHostAlloc((void **)&c->h_totals, 4096, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    delete c;
    return LG_ERR_NOMEM;
  }
  std::memset(c->h_totals, 0, 4096);
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->smem_per_sm = prop.sharedMemPerMultiprocessor;
  c->slots = precision == LG_PRECISION_F64 ? 1 : 2;
  if (const char *e = getenv("LG_ACCUM_MODE")) {
    int v = atoi(e);
    if (v >= 0 && v <= 2) c->accum_mode = v;
  }
  if (const char *e = getenv("LG_TRACE_MERGED")) c->trace_merged = atoi(e) ? 1 : 0;
  if (const char *e = getenv("LG_BIN_GRID")) c->bin_grid = std::max(0, atoi(e));
  if (const char *e = getenv("LG_FILL_WIDE")) c->fill_wide = atoi(e) ? 1 : 0;
  if (const char *e = getenv("LG_BIN_CTAS")) {
    int v = atoi(e);
    if (v >= 1 && v <= 16) c->bin_ctas = v;
  }
  if (const char *e = getenv("LG_BIN_THREADS")) {
    int v = atoi(e);
    if (v == 128 || v == 256 || v == 512 || v == 1024) c->bin_threads = v;
  }
  if (const char *e = getenv("LG_GRID_DENSITY")) {
    double v = atof(e);
    if (v > 0.01 && v < 100.0) c->grid_density = v;
  }
  if (const char *e = getenv("LG_GRID_SLOTS")) {
    int v = atoi(e);
    if (v == 1 || v == 2) c->grid_slots = v;
  }
  if (const char *e = getenv("LG_RENDER_OVERLAP")) {
    int v = atoi(e);
    if (v >= 0 && v <= 2) c->render_overlap = v;
  }
  if (const char *e = getenv("LG_PIPE_WAVES")) {
    int v = atoi(e);
    if (v >= 2 && v <= 256) c->pipe_waves = (unsigned)v;
  }
  if (const char *e = getenv("LG_PIPE_TRACE_CTAS")) {
    int v = atoi(e);
    if (v >= -4 && v <= 8) c->pipe_trace_ctas = v;
  }
  if (const char *e = getenv("LG_PIPE_RASTER_CTAS")) {
    int v = atoi(e);
    if (v >= -1 && v <= 8) c->pipe_raster_ctas = v;
  }
  if (const char *e = getenv("LG_TRACE_SLOTS")) {
    int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) c->slots = v;
  }
  *out = c;
  return LG_OK;
}

int32_t lg_destroy(lg_ctx *c) {
  if (!c) return LG_ERR_INVALID;
  cudaSetDevice(c->device);
  close_peers(c);
  close_peer_flags(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  DevBuf *bufs[] = {&c->bounds,    &c->toks,
                    &c->obj_first, &c->obj_count, &c->obj_n,    &c->ovl_start, &c->ovl_list, &c->d_lights, &c->seg,
                    &c->tags,      &c->seg64,    &c->ctr,      &c->stack,   &c->rays,     &c->img,     &c->img16,
                    &c->pixctr,    &c->tile_count, &c->tile_cursor, &c->tile_offset, &c->item_prefix,
                    &c->tile_totals, &c->item_counter, &c->tile_list, &c->seg2,       &c->tile_hist,
                    &c->sync_buf,  &c->peer_xchg,  &c->img8,       &c->grid_start,  &c->grid_obj, &c->flags, &c->flag_status,
                    &c->nest_lines, &c->nest_counts, &c->nest_points, &c->nest_pairs};
  for (DevBuf *b : bufs) release(*b);
  for (ExportBuf &e : c->exported) export_release(e);
  if (c->h_totals) cudaFreeHost(c->h_totals);
  for (int b = 0; b < 2; ++b)
    for (cudaEvent_t e : {c->ev_t0[b], c->ev_t1[b], c->ev_a0[b], c->ev_a1[b]})
      if (e) cudaEventDestroy(e);
  if (c->ev_pipe) cudaEventDestroy(c->ev_pipe);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return LG_OK;
}

const char *lg_last_error(const lg_ctx *c) { return c ? c->err.c_str() : "null context"; }

int32_t lg_scene_set(lg_ctx *c, const LgObject *objects, uint32_t n_objects, const LgGeoNode *nodes, uint32_t n_nodes,
                     const LgTraceParams *params) {
  if (!c) return LG_ERR_INVALID;
  if (!params || (n_objects && (!objects || !nodes))) return fail(c, LG_ERR_INVALID, "null scene arrays");
  LG_CUDA(c, cudaSetDevice(c->device));
  std::string err;
  int rc = lower_scene(objects, n_objects, nodes, n_nodes, *params, c->hs, err);
  if (rc) {
    c->have_scene = false;
    return fail(c, rc, err);
  }
  rc = c->precision == LG_PRECISION_F64 ? upload_scene<double>(c) : upload_scene<float>(c);
  if (rc) return rc;
  c->have_scene = true;
  c->scene_hash = hash_bytes(params, sizeof(LgTraceParams), hash_bytes(nodes, (size_t)n_nodes * sizeof(LgGeoNode),
                                                                        hash_bytes(objects, (size_t)n_objects * sizeof(LgObject))));
  if (!c->lights.empty()) { // start media depend on the scene
    rebuild_dev_lights(c);
    if ((rc = upload(c, c->d_lights, c->dev_lights))) return rc;
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return LG_OK;
}

int32_t lg_lights_set(lg_ctx *c, const LgLight *lights, uint32_t n) {
  if (!c) return LG_ERR_INVALID;
  if (n && !lights) return fail(c, LG_ERR_INVALID, "null lights");
  for (uint32_t i = 0; i < n; ++i) {
    if (lights[i].kind < LG_LIGHT_POINT || lights[i].kind > LG_LIGHT_SPOT) return fail(c, LG_ERR_INVALID, "light kind");
    if (lights[i].flags & ~(LG_LIGHT_DIRECTIONAL_NEG_R | LG_LIGHT_START_MEDIUM)) return fail(c, LG_ERR_INVALID, "light flags");
    if ((lights[i].flags & LG_LIGHT_START_MEDIUM) && !(lights[i].start_medium > 0.0 && lights[i].start_medium < 1e30))
      return fail(c, LG_ERR_INVALID, "start_medium must be a positive refractive index");
  }
  LG_CUDA(c, cudaSetDevice(c->device));
  c->lights.assign(lights, lights + n);
  c->have_lights = true;
  c->lights_hash = hash_bytes(lights, (size_t)n * sizeof(LgLight));
  rebuild_dev_lights(c);
  int rc = upload(c, c->d_lights, c->dev_lights);
  if (rc) return rc;
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

int32_t lg_drawing_object_set(lg_ctx *c, const LgObject *object, const LgGeoNode *nodes, uint32_t n_nodes) {
  if (!c) return LG_ERR_INVALID;
  LG_CUDA(c, cudaSetDevice(c->device));
  if (!object) {
    c->have_drawing = false;
  } else {
    if (!nodes || n_nodes == 0) return fail(c, LG_ERR_INVALID, "null geometry nodes");
    LgTraceParams prm{};
    prm.max_bounce = 1;
    prm.canvas_tlbr[0] = 1.0, prm.canvas_tlbr[1] = -1.0, prm.canvas_tlbr[2] = -1.0, prm.canvas_tlbr[3] = 1.0;
    std::string err;
    HostScene tmp;
    int rc = lower_scene(object, 1, nodes, n_nodes, prm, tmp, err);
    if (rc) return fail(c, rc, err);
    c->drawing_hs = tmp;
    c->have_drawing = true;
  }
  if (c->have_lights) { // start media may have changed
    rebuild_dev_lights(c);
    int rc = upload(c, c->d_lights, c->dev_lights);
    if (rc) return rc;
    c->lights_hash = mix_sig(hash_bytes(c->lights.data(), c->lights.size() * sizeof(LgLight)), c->have_drawing ? 1 : 0);
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return LG_OK;
}

int32_t lg_shard_set(lg_ctx *c, uint32_t rank, uint32_t world) {
  if (!c) return LG_ERR_INVALID;
  if (world == 0 || rank >= world) return fail(c, LG_ERR_INVALID, "rank/world");
  LG_CUDA(c, cudaSetDevice(c->device));
  c->rank = rank, c->world = world;
  rebuild_dev_lights(c);
  int rc = upload(c, c->d_lights, c->dev_lights);
  if (rc) return rc;
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

int32_t lg_segment_capacity_set(lg_ctx *c, uint64_t n) {
  if (!c) return LG_ERR_INVALID;
  if (n == 0) return fail(c, LG_ERR_INVALID, "capacity 0");
  LG_CUDA(c, cudaSetDevice(c->device));
  c->seg_cap = n;
  release(c->seg), release(c->tags), release(c->seg64);
  c->seg_count = 0;
  return LG_OK;
}

int32_t lg_accumulate_mode_set(lg_ctx *c, int32_t mode) {
  if (!c) return LG_ERR_INVALID;
  if (mode < 0 || mode > 2) return fail(c, LG_ERR_INVALID, "accumulate mode");
  if (mode == 2 && c->blend_generic) return fail(c, LG_ERR_UNSUPPORTED, "the tile-binned resolve implements the default blend state only");
  c->accum_mode = mode;
  return LG_OK;
}

int32_t lg_blend_set(lg_ctx *c, const LgBlendState *st) {
  if (!c) return LG_ERR_INVALID;
  LgBlendState def{};
  def.color = {LG_BF_ONE, LG_BF_ONE, LG_BO_ADD};       // mod.rs:65-69
  def.alpha = {LG_BF_SRC_ALPHA, LG_BF_ONE, LG_BO_ADD}; // mod.rs:60-64
  const LgBlendState s = st ? *st : def;
  auto source_only = [](int f) {
    return f == LG_BF_ZERO || f == LG_BF_ONE || f == LG_BF_SRC || f == LG_BF_ONE_MINUS_SRC || f == LG_BF_SRC_ALPHA ||
           f == LG_BF_ONE_MINUS_SRC_ALPHA || f == LG_BF_CONSTANT || f == LG_BF_ONE_MINUS_CONSTANT;
  };
  bool linear = true;
  for (const LgBlendComponent *k : {&s.color, &s.alpha}) {
    if (k->operation < LG_BO_ADD || k->operation > LG_BO_MAX || k->src_factor < LG_BF_ZERO ||
        k->src_factor > LG_BF_ONE_MINUS_CONSTANT || k->dst_factor < LG_BF_ZERO || k->dst_factor > LG_BF_ONE_MINUS_CONSTANT)
      return fail(c, LG_ERR_INVALID, "blend state: unknown factor or operation");
    if (k->operation == LG_BO_MIN || k->operation == LG_BO_MAX) {
      linear = false; // wgpu ignores the factors of Min / Max
    } else if (k->operation == LG_BO_SUBTRACT) {
      return fail(c, LG_ERR_UNSUPPORTED, "blend Subtract (src - dst) depends on the fragment order");
    } else if (k->dst_factor != LG_BF_ONE || !source_only(k->src_factor)) {
      return fail(c, LG_ERR_UNSUPPORTED,
                  "blend Add / ReverseSubtract is order-independent only with dst_factor One and a source-only src_factor");
    }
  }
  if (c->accum_mode == 2 && std::memcmp(&s, &def, sizeof s) != 0)
    return fail(c, LG_ERR_UNSUPPORTED, "non-default blend states use the direct resolve (lg_accumulate_mode_set 0 or 1)");
  c->blend_generic = std::memcmp(&s, &def, sizeof s) != 0;
  c->blend_linear = linear;
  c->blend.color_factor = s.color.src_factor, c->blend.alpha_factor = s.alpha.src_factor;
  c->blend.color_op = s.color.operation, c->blend.alpha_op = s.alpha.operation;
  std::memcpy(c->blend.constant, s.constant, 16);
  return LG_OK;
}

int32_t lg_tile_map_enable(lg_ctx *c, int32_t enable) {
  if (!c) return LG_ERR_INVALID;
  c->grid_on = enable != 0;
  return LG_OK;
}

int32_t lg_tags_enable(lg_ctx *c, int32_t enable) {
  if (!c) return LG_ERR_INVALID;
  c->tags_on = enable != 0;
  return LG_OK;
}

int32_t lg_emit_rays(lg_ctx *c, uint32_t light, uint64_t first, uint64_t count, LgRay *dst) {
  if (!c) return LG_ERR_INVALID;
  if (light >= c->dev_lights.size() || (count && !dst)) return fail(c, LG_ERR_INVALID, "light index / dst");
  if (count == 0) return LG_OK;
  LG_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure(c, c->rays, count * sizeof(LgRay));
  if (rc) return rc;
  const unsigned grid = (unsigned)((count + 255) / 256);
  emit_rays_kernel<<<grid, 256, 0, c->stream>>>(c->dev_lights[light], first, count, (LgRay *)c->rays.p);
  LG_CUDA(c, cudaGetLastError());
  c->launches++;
  LG_CUDA(c, cudaMemcpyAsync(dst, c->rays.p, count * sizeof(LgRay), cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

static void fill_stats(lg_ctx *c, LgTraceStats *st, unsigned long long rays, unsigned long long steps,
                       unsigned long long segs, float tms, unsigned tl) {
  if (!st) return;
  st->primary_rays += rays;
  st->ray_steps += steps;
  st->object_tests += steps * (unsigned long long)c->n_obj;
  st->segments += segs;
  st->trace_ms += tms;
  st->trace_launches += tl;
}

int32_t lg_trace(lg_ctx *c, LgTraceStats *stats) {
  int rc = need_scene_lights(c, true);
  if (rc) return rc;
  LG_CUDA(c, cudaSetDevice(c->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  if ((rc = prepare_trace_buffers(c))) return rc;
  TraceCounters k{};
  float ms = 0.f;
  c->seg_count = 0;
  if ((rc = trace_range(c, nullptr, 0, c->shard_rays, k, &ms))) return rc;
  fill_stats(c, stats, c->shard_rays, k.ray_steps, k.seg_count, ms, 1);
  if (k.seg_overflow) {
    c->seg_count = 0;
    return fail(c, LG_ERR_OVERFLOW,
                "segment buffer too small: " + std::to_string(k.seg_count) + " segments, capacity " +
                    std::to_string(c->seg_cap) + " (raise lg_segment_capacity_set or use lg_render)");
  }
  c->seg_count = k.seg_count;
  return LG_OK;
}

int32_t lg_trace_rays(lg_ctx *c, const LgRay *rays, uint64_t n, LgTraceStats *stats) {
  int rc = need_scene_lights(c, false);
  if (rc) return rc;
  if (n && !rays) return fail(c, LG_ERR_INVALID, "null rays");
  LG_CUDA(c, cudaSetDevice(c->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  if ((rc = prepare_trace_buffers(c))) return rc;
  if ((rc = ensure(c, c->rays, n * sizeof(LgRay)))) return rc;
  if (n) LG_CUDA(c, cudaMemcpyAsync(c->rays.p, rays, n * sizeof(LgRay), cudaMemcpyHostToDevice, c->stream));
  TraceCounters k{};
  float ms = 0.f;
  c->seg_count = 0;
  double ray_bound = 0;
  // directions must be unit vectors (Ray::from_origin normalises, light.rs:172): the broad phase and the range
  // filter of the trace kernel rely on it, a longer or shorter one would silently lose hits
  const double tol = 16.0 * (c->precision == LG_PRECISION_F64 ? std::numeric_limits<double>::epsilon()
                                                              : (double)std::numeric_limits<float>::epsilon());
  for (uint64_t i = 0; i < n; ++i) {
    ray_bound = std::fmax(ray_bound, std::fmax(std::fabs(rays[i].origin[0]), std::fabs(rays[i].origin[1])));
    const double dx = rays[i].direction[0], dy = rays[i].direction[1];
    if (!(std::fabs(dx * dx + dy * dy - 1.0) <= tol))
      return fail(c, LG_ERR_INVALID, "ray " + std::to_string(i) + ": direction is not a unit vector (|d|^2 - 1 = " +
                                         std::to_string(dx * dx + dy * dy - 1.0) + ")");
  }
  if (!(ray_bound < 1e30)) return fail(c, LG_ERR_INVALID, "ray origin is not finite");
  if ((rc = trace_range(c, (const LgRay *)c->rays.p, 0, n, k, &ms, ray_bound))) return rc;
  fill_stats(c, stats, n, k.ray_steps, k.seg_count, ms, 1);
  if (k.seg_overflow) return fail(c, LG_ERR_OVERFLOW, "segment buffer too small");
  c->seg_count = k.seg_count;
  return LG_OK;
}

int32_t lg_segments_count(lg_ctx *c, uint64_t *n) {
  if (!c || !n) return LG_ERR_INVALID;
  *n = c->seg_count;
  return LG_OK;
}

int32_t lg_segments_read(lg_ctx *c, LgSegment *dst, LgSegmentTag *tags, LgSegmentF64 *f64, uint64_t cap, uint64_t *n) {
  if (!c || !n) return LG_ERR_INVALID;
  LG_CUDA(c, cudaSetDevice(c->device));
  const uint64_t m = std::min<uint64_t>(cap, c->seg_count);
  *n = m;
  if (m == 0) return LG_OK;
  if (dst) LG_CUDA(c, cudaMemcpyAsync(dst, c->seg.p, m * sizeof(LgSegment), cudaMemcpyDeviceToHost, c->stream));
  if (tags) {
    if (!c->tags_on || !c->tags.p) return fail(c, LG_ERR_STATE, "tags were not enabled for the last trace");
    LG_CUDA(c, cudaMemcpyAsync(tags, c->tags.p, m * sizeof(LgSegmentTag), cudaMemcpyDeviceToHost, c->stream));
  }
  if (f64) {
    if (!c->tags_on || c->precision != LG_PRECISION_F64 || !c->seg64.p)
      return fail(c, LG_ERR_STATE, "f64 endpoints need an F64 context with tags enabled");
    LG_CUDA(c, cudaMemcpyAsync(f64, c->seg64.p, m * sizeof(LgSegmentF64), cudaMemcpyDeviceToHost, c->stream));
  }
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

int32_t lg_image_configure(lg_ctx *c, uint32_t width, uint32_t height) {
  if (!c) return LG_ERR_INVALID;
  if (width == 0 || height == 0 || width > 32768 || height > 32768) return fail(c, LG_ERR_INVALID, "image size");
  LG_CUDA(c, cudaSetDevice(c->device));
  if (c->W != (int)width || c->H != (int)height)
    for (ExportBuf &e : c->exported) export_release(e); // a frame of another size: importers must fetch a new handle
  c->W = (int)width, c->H = (int)height;
  c->peers_ready = false; // the buffers may move: the next lg_image_reduce re-exchanges the peer mappings
  c->img16_valid = false;
  int rc;
  if ((rc = ensure(c, c->img, (size_t)width * height * 16))) return rc;
  if ((rc = ensure(c, c->pixctr, 8))) return rc;
  return lg_image_clear(c, 1.0f);
}

int32_t lg_image_clear(lg_ctx *c, float clear_alpha) {
  int rc = need_image(c);
  if (rc) return rc;
  LG_CUDA(c, cudaSetDevice(c->device));
  c->img16_valid = false;
  clear_image_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>((float4 *)c->img.p, (size_t)c->W * c->H, clear_alpha);
  LG_CUDA(c, cudaGetLastError());
  c->launches++;
  LG_CUDA(c, cudaMemsetAsync(c->pixctr.p, 0, 8, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

int32_t lg_accumulate_traced(lg_ctx *c, LgTraceStats *stats) {
  int rc = need_image(c);
  if (rc) return rc;
  LG_CUDA(c, cudaSetDevice(c->device));
  float ms = 0.f;
  unsigned l = 0;
  if ((rc = accumulate_device_segments(c, c->seg_count, &ms, &l))) return rc;
  if (stats) {
    stats->accumulate_ms += ms;
    stats->accumulate_launches += l;
    if ((rc = read_pixel_counter(c, &stats->pixel_updates))) return rc;
  }
  return LG_OK;
}

} // extern "C"
namespace {
// update_vertex_buffer + render for `n` vertex pairs that already live on the device
int accumulate_device_pairs(lg_ctx *c, const LgVertexPair *d_pairs, uint64_t n, unsigned long long sig, LgTraceStats *stats) {
  int rc;
  const bool tiled = use_tiled(c, n, sig);
  uint64_t frag0 = 0;
  if ((rc = read_pixel_counter(c, &frag0))) return rc;
  LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  unsigned nl = 0;
  if (tiled) {
    if ((rc = ensure(c, c->seg2, n * sizeof(Seg2)))) return rc;
    pairs_to_seg2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_pairs, (Seg2 *)c->seg2.p, n);
    LG_CUDA(c, cudaGetLastError());
    c->launches++, nl++;
    if ((rc = accumulate_tiled<Seg2>(c, c->stream, (const Seg2 *)c->seg2.p, n, nullptr, nullptr, c->h_totals, &nl))) return rc;
  } else {
    if (c->blend_generic)
      accumulate_pairs_kernel<true><<<accum_grid(c), kAccumBlock, 0, c->stream>>>(accum_args(c), d_pairs, n);
    else
      accumulate_pairs_kernel<false><<<accum_grid(c), kAccumBlock, 0, c->stream>>>(accum_args(c), d_pairs, n);
    LG_CUDA(c, cudaGetLastError());
    c->launches++, nl++;
  }
  LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (tiled && (rc = tiled_finish<Seg2>(c, c->stream, (const Seg2 *)c->seg2.p, n, c->h_totals, &nl))) return rc;
  float t = 0.f;
  LG_CUDA(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
  uint64_t frag1 = 0;
  if ((rc = read_pixel_counter(c, &frag1))) return rc;
  if (c->accum_mode == 0) note_accum_cost(c, tiled, t, frag1 - frag0, n, sig);
  if (stats) {
    stats->accumulate_ms += t;
    stats->accumulate_launches += nl;
    stats->segments += n;
    stats->pixel_updates = frag1;
  }
  return LG_OK;
}
} // namespace
extern "C" {

int32_t lg_accumulate_segments(lg_ctx *c, const LgVertexPair *pairs, uint64_t n, LgTraceStats *stats) {
  int rc = need_image(c);
  if (rc) return rc;
  if (n && !pairs) return fail(c, LG_ERR_INVALID, "null pairs");
  if (n == 0) return LG_OK;
  c->img16_valid = false;
  LG_CUDA(c, cudaSetDevice(c->device));
  // reuse the ray staging buffer for the upload (update_vertex_buffer makes a NEW buffer per frame)
  if ((rc = ensure(c, c->rays, n * sizeof(LgVertexPair)))) return rc;
  LG_CUDA(c, cudaMemcpyAsync(c->rays.p, pairs, n * sizeof(LgVertexPair), cudaMemcpyHostToDevice, c->stream));
  // host lines: workloads are told apart by image and segment-count class only
  const unsigned long long sig = mix_sig(mix_sig(2, ((unsigned long long)c->W << 32) | (unsigned)c->H), 63 - __builtin_clzll(n | 1));
  return accumulate_device_pairs(c, (const LgVertexPair *)c->rays.p, n, sig, stats);
}

// StringMod::draw with `nested: Some(inner)` (string_mod.rs:87-101,152-158)
int32_t lg_string_mod_nested(lg_ctx *c, const LgStringMod *outer, const LgStringMod *inner, const LgModRemColor *inner_rules,
                             uint32_t n_inner_rules, LgTraceStats *stats) {
  int rc = need_image(c);
  if (rc) return rc;
  if (!outer || !inner || (n_inner_rules && !inner_rules)) return fail(c, LG_ERR_INVALID, "null string mod");
  for (const LgStringMod *sm : {outer, inner}) {
    if (sm->curve < LG_CURVE_CIRCLE || sm->curve > LG_CURVE_LISSAJOUS) return fail(c, LG_ERR_INVALID, "Curve");
    if (sm->mode < LG_SM_ADD || sm->mode > LG_SM_BASE) return fail(c, LG_ERR_INVALID, "StringModMode");
  }
  c->nest_lines_n = c->nest_points_n = 0;
  const unsigned long long L = outer->modulo;
  if (L > (1ull << 20)) return fail(c, LG_ERR_UNSUPPORTED, "nested string mod: more than 2^20 outer chords (L^2 crossing tests)");
  if (L < 2 || inner->modulo == 0) return LG_OK; // no crossings, or draw_init_points of nothing (string_mod.rs:106-108)
  c->img16_valid = false;
  LG_CUDA(c, cudaSetDevice(c->device));
  // 1. the outer pattern's chords, f64 (their colours play no role)
  StringModArgs S;
  S.sm = *outer, S.rules = nullptr, S.n_rules = 0, S.first = 0, S.count = L;
  if ((rc = ensure(c, c->nest_lines, L * sizeof(LgVertexPair)))) return rc;
  string_mod_pairs_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(S, (LgVertexPair *)c->nest_lines.p);
  LG_CUDA(c, cudaGetLastError());
  c->nest_lines_n = L;
  // 2. crossings in the reference's order: count per block, scan, ordered write
  const unsigned long long n_cand = L * (L - 1ull);
  const unsigned long long per_block = (unsigned long long)kNestBlock * kNestPer;
  const unsigned long long n_blocks = (n_cand + per_block - 1) / per_block;
  if (n_blocks > 0x7fffffffull) return fail(c, LG_ERR_UNSUPPORTED, "nested string mod: too many crossing candidates");
  if ((rc = ensure(c, c->nest_counts, (n_blocks + 1) * 8))) return rc;
  unsigned long long *counts = (unsigned long long *)c->nest_counts.p;
  nested_count_kernel<<<(unsigned)n_blocks, kNestBlock, 0, c->stream>>>((const LgVertexPair *)c->nest_lines.p, L, n_cand, counts);
  LG_CUDA(c, cudaGetLastError());
  nested_scan_kernel<<<1, 1024, 0, c->stream>>>(counts, n_blocks, counts + n_blocks);
  LG_CUDA(c, cudaGetLastError());
  unsigned long long P = 0;
  LG_CUDA(c, cudaMemcpyAsync(&P, counts + n_blocks, 8, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  c->launches += 3;
  if (stats) stats->accumulate_launches += 3;
  if (P == 0) return LG_OK;
  if ((rc = ensure(c, c->nest_points, P * sizeof(double2)))) return rc;
  nested_write_kernel<<<(unsigned)n_blocks, kNestBlock, 0, c->stream>>>((const LgVertexPair *)c->nest_lines.p, L, n_cand, counts,
                                                                        (double2 *)c->nest_points.p);
  LG_CUDA(c, cudaGetLastError());
  c->nest_points_n = P;
  // 3. the inner pattern's chords between those points, then the line pass
  std::vector<LgModRemColor> rv(inner_rules, inner_rules + n_inner_rules);
  if ((rc = upload(c, c->rays, rv))) return rc;
  LG_CUDA(c, cudaStreamSynchronize(c->stream)); // rv dies at the end of this call; the kernel below reads the copy
  StringModArgs I;
  I.sm = *inner, I.rules = (const LgModRemColor *)c->rays.p, I.n_rules = n_inner_rules, I.first = 0, I.count = inner->modulo;
  if ((rc = ensure(c, c->nest_pairs, inner->modulo * sizeof(LgVertexPair)))) return rc;
  nested_pairs_kernel<<<(unsigned)((inner->modulo + 255) / 256), 256, 0, c->stream>>>(I, (const double2 *)c->nest_points.p, P,
                                                                                      (LgVertexPair *)c->nest_pairs.p);
  LG_CUDA(c, cudaGetLastError());
  c->launches += 2;
  if (stats) stats->accumulate_launches += 2;
  unsigned long long sig = hash_bytes(inner, sizeof(LgStringMod), hash_bytes(outer, sizeof(LgStringMod), 4));
  sig = mix_sig(sig, ((unsigned long long)c->W << 32) | (unsigned)c->H);
  return accumulate_device_pairs(c, (const LgVertexPair *)c->nest_pairs.p, inner->modulo, sig, stats);
}

// The outer chords and the crossing points of the last lg_string_mod_nested call (tests, inspection)
int32_t lg_string_mod_nested_read(lg_ctx *c, LgVertexPair *outer_chords, uint64_t chord_cap, double *crossings_xy,
                                  uint64_t crossing_cap, uint64_t *n_chords, uint64_t *n_crossings) {
  if (!c) return LG_ERR_INVALID;
  LG_CUDA(c, cudaSetDevice(c->device));
  if (n_chords) *n_chords = c->nest_lines_n;
  if (n_crossings) *n_crossings = c->nest_points_n;
  const uint64_t nl = std::min<uint64_t>(chord_cap, c->nest_lines_n), np = std::min<uint64_t>(crossing_cap, c->nest_points_n);
  if (outer_chords && nl)
    LG_CUDA(c, cudaMemcpyAsync(outer_chords, c->nest_lines.p, nl * sizeof(LgVertexPair), cudaMemcpyDeviceToHost, c->stream));
  if (crossings_xy && np)
    LG_CUDA(c, cudaMemcpyAsync(crossings_xy, c->nest_points.p, np * 16, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

int32_t lg_string_mod(lg_ctx *c, const LgStringMod *sm, const LgModRemColor *rules, uint32_t n_rules, uint64_t first,
                      uint64_t count, LgTraceStats *stats) {
  int rc = need_image(c);
  if (rc) return rc;
  if (!sm || (n_rules && !rules)) return fail(c, LG_ERR_INVALID, "null string mod");
  if (sm->curve < LG_CURVE_CIRCLE || sm->curve > LG_CURVE_LISSAJOUS) return fail(c, LG_ERR_INVALID, "Curve");
  if (sm->mode < LG_SM_ADD || sm->mode > LG_SM_BASE) return fail(c, LG_ERR_INVALID, "StringModMode");
  if (sm->modulo == 0) return LG_OK; // draw_init_points: points.is_empty() -> no lines (string_mod.rs:106-108)
  c->img16_valid = false;
  LG_CUDA(c, cudaSetDevice(c->device));
  if (count == 0) { // this context's shard of all chords: a contiguous block (chords are uniform work)
    first = (uint64_t)(((unsigned __int128)sm->modulo * c->rank) / c->world);
    count = (uint64_t)(((unsigned __int128)sm->modulo * (c->rank + 1)) / c->world) - first;
  }
  if (first + count > sm->modulo) return fail(c, LG_ERR_INVALID, "chord range beyond modulo");
  std::vector<LgModRemColor> rv(rules, rules + n_rules);
  if ((rc = upload(c, c->rays, rv))) return rc;
  StringModArgs S;
  S.sm = *sm;
  S.rules = (const LgModRemColor *)c->rays.p;
  S.n_rules = n_rules;
  S.first = first, S.count = count;
  unsigned long long sig = hash_bytes(sm, sizeof(LgStringMod), 3);
  sig = mix_sig(mix_sig(sig, ((unsigned long long)c->W << 32) | (unsigned)c->H), mix_sig(first, count));
  const bool tiled = count && use_tiled(c, count, sig);
  uint64_t frag0 = 0;
  if ((rc = read_pixel_counter(c, &frag0))) return rc;
  LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  unsigned nl = 0;
  if (tiled) {
    // the tiled path needs the chords in memory; the direct path generates them in-kernel
    if ((rc = ensure(c, c->seg2, count * sizeof(Seg2)))) return rc;
    string_mod_seg2_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(S, (Seg2 *)c->seg2.p);
    LG_CUDA(c, cudaGetLastError());
    c->launches++, nl++;
    if ((rc = accumulate_tiled<Seg2>(c, c->stream, (const Seg2 *)c->seg2.p, count, nullptr, nullptr, c->h_totals, &nl))) return rc;
  } else if (count) {
    if (c->blend_generic)
      string_mod_kernel<true><<<accum_grid(c), kAccumBlock, 0, c->stream>>>(accum_args(c), S);
    else
      string_mod_kernel<false><<<accum_grid(c), kAccumBlock, 0, c->stream>>>(accum_args(c), S);
    LG_CUDA(c, cudaGetLastError());
    c->launches++, nl++;
  }
  LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (tiled && (rc = tiled_finish<Seg2>(c, c->stream, (const Seg2 *)c->seg2.p, count, c->h_totals, &nl))) return rc;
  float t = 0.f;
  LG_CUDA(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
  uint64_t frag1 = 0;
  if ((rc = read_pixel_counter(c, &frag1))) return rc;
  if (c->accum_mode == 0 && count) note_accum_cost(c, tiled, t, frag1 - frag0, count, sig);
  if (stats) {
    stats->accumulate_ms += t;
    stats->accumulate_launches += nl;
    stats->segments += count;
    stats->pixel_updates = frag1;
  }
  return LG_OK;
}

} // extern "C"
namespace {

// Rays [first, end) of this context's shard, one wave after the other through the whole segment buffer: trace,
// wait, accumulate, wait.  A wave that overflows is traced again with half the rays (nothing of it has been
// accumulated yet).
int render_range_sync(lg_ctx *c, LgTraceStats *stats, unsigned long long first, unsigned long long end, double &est) {
  int rc;
  unsigned long long done = first;
  unsigned long long limit = ~0ull; // shrinks when a wave overflowed
  while (done < end) {
    unsigned long long wave = (unsigned long long)((double)c->seg_cap / (est * 1.25 + 1.0));
    if (wave < 1) wave = 1;
    if (wave > limit) wave = limit;
    if (wave > end - done) wave = end - done;
    TraceCounters k{};
    float ms = 0.f;
    if ((rc = trace_range(c, nullptr, done, done + wave, k, &ms))) return rc;
    stats->trace_ms += ms;
    stats->trace_launches += 1;
    if (k.seg_overflow) {
      if (wave <= 1) return fail(c, LG_ERR_OVERFLOW, "one ray does not fit the segment buffer");
      est = std::max(est, (double)k.seg_count / (double)wave);
      limit = wave / 2; // guaranteed progress: the retry traces at most half as many rays
      continue;
    }
    limit = ~0ull;
    est = std::max(1.0, (double)k.seg_count / (double)wave);
    fill_stats(c, stats, wave, k.ray_steps, k.seg_count, 0.f, 0);
    c->seg_count = k.seg_count;
    if ((rc = accumulate_device_segments(c, k.seg_count, &stats->accumulate_ms, &stats->accumulate_launches))) return rc;
    done += wave;
  }
  return LG_OK;
}

// Has the automatic accumulate mode settled on the tile bins for this workload (and is this not one of its probe calls)?
bool auto_settled_on_tiles(lg_ctx *c, unsigned long long sig) {
  auto it = c->auto_stats.find(sig);
  if (it == c->auto_stats.end()) return false;
  const lg_ctx::AutoStat &a = it->second;
  if (a.samples[1] < 2 || a.samples[2] < 2) return false;
  if ((a.calls + 1) % 64 == 0) return false; // the next call re-probes the other resolve: let the sequential path take it
  return a.ns_per_frag[2] <= a.ns_per_frag[1];
}

// The frame as a two-stage pipeline over waves of rays (tile-binned resolve only).  The segment buffer is used as
// two halves; wave k is traced into half k & 1 on `stream` with one CTA per SM less than would fit, and its
// count / scan / fill / raster passes are queued on `stream2` behind an event -- they read the segment count from the
// trace kernel's counters on the device, so the host never waits in between: while the passes of wave k run, wave
// k + 1 is being traced on the SMs' remaining issue slots, registers and shared memory.  The host looks at a wave
// again only when its half of the buffer is needed for wave k + 2 (by then it has long finished): statistics, and the
// two rare repairs -- a wave that overflowed its half (skipped by the passes on the device: its rays are traced
// again at the end, in smaller waves) and a pair list that turned out too small (that wave's passes run again).
int render_pipelined(lg_ctx *c, LgTraceStats *stats, double &est) {
  const unsigned long long total = c->shard_rays;
  const unsigned long long half = c->seg_cap / 2;
  LgSegment *seg[2] = {(LgSegment *)c->seg.p, (LgSegment *)c->seg.p + half};
  TraceCounters *ctr = (TraceCounters *)c->ctr.p;
  unsigned long long *hst[2] = {c->h_totals + 8, c->h_totals + 24}; // [0..3] tile totals, [4..7] TraceCounters
  struct Wave {
    unsigned long long first, end;
    bool busy;
  } w[2] = {{0, 0, false}, {0, 0, false}};
  std::vector<std::pair<unsigned long long, unsigned long long>> redo; // ray ranges whose wave overflowed its half
  int rc;
  // everything queued on `stream` so far (the clear) comes first on stream2 as well
  LG_CUDA(c, cudaEventRecord(c->ev_pipe, c->stream));
  LG_CUDA(c, cudaStreamWaitEvent(c->stream2, c->ev_pipe, 0));

  auto queue_passes = [&](int b, unsigned long long expect) -> int {
    LG_CUDA(c, cudaEventRecord(c->ev_a0[b], c->stream2));
    int r = accumulate_tiled<LgSegment>(c, c->stream2, seg[b], half, &ctr[b].seg_count, &ctr[b].seg_overflow, hst[b],
                                        &stats->accumulate_launches, expect);
    if (r) return r;
    LG_CUDA(c, cudaMemcpyAsync(hst[b] + 4, &ctr[b], sizeof(TraceCounters), cudaMemcpyDeviceToHost, c->stream2));
    LG_CUDA(c, cudaEventRecord(c->ev_a1[b], c->stream2));
    return LG_OK;
  };
  auto reclaim = [&](int b) -> int { // the wave in half b has been drawn (or skipped): statistics and repairs
    if (!w[b].busy) return LG_OK;
    w[b].busy = false;
    LG_CUDA(c, cudaEventSynchronize(c->ev_a1[b]));
    TraceCounters k;
    std::memcpy(&k, hst[b] + 4, sizeof k);
    float tms = 0.f, ams = 0.f;
    LG_CUDA(c, cudaEventElapsedTime(&tms, c->ev_t0[b], c->ev_t1[b]));
    LG_CUDA(c, cudaEventElapsedTime(&ams, c->ev_a0[b], c->ev_a1[b]));
    stats->trace_ms += tms, stats->accumulate_ms += ams;
    if (k.stack_overflow) return fail(c, LG_ERR_OVERFLOW, "split stack overflow (max_bounce with refraction > 65)");
    const unsigned long long rays = w[b].end - w[b].first;
    if (k.seg_overflow) { // nothing of this wave was drawn
      est = std::max(est, (double)k.seg_count / (double)rays);
      redo.emplace_back(w[b].first, w[b].end);
      return LG_OK;
    }
    est = std::max(1.0, (double)k.seg_count / (double)rays);
    fill_stats(c, stats, rays, k.ray_steps, k.seg_count, 0.f, 0);
    if (k.seg_count > 0) c->pairs_per_seg_est = std::max(1.0, (double)hst[b][0] / (double)k.seg_count);
    if (hst[b][2]) { // the pair list was too small: the later wave's passes (queued already) share the bin buffers, so
      // let them finish, then run this wave's passes again with the list sized from its own count
      LG_CUDA(c, cudaStreamSynchronize(c->stream2)); // (the other wave's own status is looked at when it is reclaimed)
      if ((rc = accumulate_tiled<LgSegment>(c, c->stream2, seg[b], k.seg_count, nullptr, nullptr, c->h_totals,
                                            &stats->accumulate_launches, k.seg_count)))
        return rc;
      LG_CUDA(c, cudaStreamSynchronize(c->stream2));
      if (c->h_totals[2]) return fail(c, LG_ERR_NOMEM, "pair list still too small after it was sized from the wave's own count");
    }
    return LG_OK;
  };

  unsigned long long wave_rays = std::max<unsigned long long>((total + c->pipe_waves - 1) / c->pipe_waves, 1ull);
  unsigned long long done = 0;
  for (unsigned k = 0; done < total; ++k) {
    const int b = (int)(k & 1u);
    if ((rc = reclaim(b))) return rc;
    unsigned long long fit = (unsigned long long)((double)half / (est * 1.25 + 1.0));
    if (fit < 1) fit = 1;
    const unsigned long long wave = std::min(std::min(wave_rays, fit), total - done);
    LG_CUDA(c, cudaEventRecord(c->ev_t0[b], c->stream));
    if ((rc = trace_async(c, c->stream, nullptr, done, done + wave, seg[b], half, &ctr[b], c->pipe_trace_ctas))) return rc;
    LG_CUDA(c, cudaEventRecord(c->ev_t1[b], c->stream));
    stats->trace_launches += 1;
    LG_CUDA(c, cudaStreamWaitEvent(c->stream2, c->ev_t1[b], 0));
    // this wave's passes run next to the NEXT wave's trace kernel: leave its CTAs their shared memory (the last
    // wave's passes have the device to themselves)
    c->raster_cap_now = 0;
    if (done + wave < total) {
      if (c->pipe_raster_ctas >= 0) {
        c->raster_cap_now = c->pipe_raster_ctas;
      } else if (c->last_trace_smem > 0) {
        const size_t raster_smem = (size_t)kRasterWarps * ((size_t)kTileFloat4 * sizeof(float4) + sizeof(RasterScratch)) + 1024;
        const size_t taken = (size_t)c->last_trace_ctas * (c->last_trace_smem + 1024 + 64);
        c->raster_cap_now = (int)std::max<size_t>(1, c->smem_per_sm > taken ? (c->smem_per_sm - taken) / raster_smem : 0);
      }
    }
    rc = queue_passes(b, (unsigned long long)((double)wave * est * 1.1) + 1024ull);
    c->raster_cap_now = 0;
    if (rc) return rc;
    w[b] = {done, done + wave, true};
    done += wave;
  }
  // drain, older wave first
  const int last = w[0].busy && w[1].busy ? (w[0].first < w[1].first ? 0 : 1) : (w[0].busy ? 0 : 1);
  if ((rc = reclaim(last))) return rc;
  if ((rc = reclaim(last ^ 1))) return rc;
  LG_CUDA(c, cudaStreamSynchronize(c->stream2));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  for (const auto &r : redo)
    if ((rc = render_range_sync(c, stats, r.first, r.second, est))) return rc;
  return LG_OK;
}

} // namespace
extern "C" {

int32_t lg_render(lg_ctx *c, LgTraceStats *stats) {
  int rc = need_scene_lights(c, true);
  if (rc) return rc;
  if ((rc = need_image(c))) return rc;
  LG_CUDA(c, cudaSetDevice(c->device));
  LgTraceStats local;
  if (!stats) stats = &local;
  std::memset(stats, 0, sizeof *stats);
  if ((rc = prepare_trace_buffers(c))) return rc;
  if ((rc = ensure_bounds(c, 0.0))) return rc;
  const unsigned mb = c->hs.params.max_bounce;
  double est = c->segs_per_ray_est > 0 ? c->segs_per_ray_est : std::min<double>(2.0 * (mb ? mb : 1), 64.0);
  // the wave pipeline: tile-binned resolve (forced, or what the automatic mode has settled on for this workload),
  // enough rays to cut into waves that still fill the device
  const unsigned long long total = c->shard_rays;
  const unsigned long long expect_segments = (unsigned long long)((double)total * est);
  bool pipe = c->render_overlap != 0 && !c->tags_on && !c->blend_generic && c->accum_mode != 1 && total >= 2 &&
              tiled_possible(c, std::max<unsigned long long>(1ull, std::min<unsigned long long>(expect_segments, (1ull << 32) - 1)));
  if (pipe && c->render_overlap == 1) {
    pipe = total >= (1ull << 20) && expect_segments / c->pipe_waves >= kTiledMinSegments &&
           (c->accum_mode == 2 || auto_settled_on_tiles(c, traced_signature(c)));
  }
  if (pipe && c->seg_cap / 2 >= (1ull << 32)) pipe = false; // 32-bit segment indices in the pair lists
  c->img16_valid = false;
  if (pipe) {
    if (c->accum_mode == 0) ++c->auto_stats[traced_signature(c)].calls;
    rc = render_pipelined(c, stats, est);
  } else {
    rc = render_range_sync(c, stats, 0, total, est);
  }
  if (rc) {
    cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream);
    return rc;
  }
  c->segs_per_ray_est = est;
  if (pipe) c->seg_count = 0; // the halves hold the last two waves: nothing lg_segments_read could hand out as one list
  return read_pixel_counter(c, &stats->pixel_updates);
}

int32_t lg_image_read(lg_ctx *c, int32_t format, void *dst, size_t pitch) {
  int rc = need_image(c);
  if (rc) return rc;
  if (!dst) return fail(c, LG_ERR_INVALID, "null dst");
  LG_CUDA(c, cudaSetDevice(c->device));
  const size_t npx = (size_t)c->W * c->H;
  if (format == LG_RGBA32F) {
    const size_t row = (size_t)c->W * 16;
    if (pitch == 0) pitch = row;
    if (pitch < row) return fail(c, LG_ERR_INVALID, "pitch");
    LG_CUDA(c, cudaMemcpy2DAsync(dst, pitch, c->img.p, row, row, c->H, cudaMemcpyDeviceToHost, c->stream));
  } else if (format == LG_RGBA16F) {
    const size_t row = (size_t)c->W * 8;
    if (pitch == 0) pitch = row;
    if (pitch < row) return fail(c, LG_ERR_INVALID, "pitch");
    if ((rc = ensure(c, c->img16, npx * 8))) return rc;
    if (!c->img16_valid) { // the peer-fused reduce already left the finalized frame here
      finalize_f16_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>((const float4 *)c->img.p, (uint2 *)c->img16.p, npx);
      LG_CUDA(c, cudaGetLastError());
      c->launches++;
    }
    LG_CUDA(c, cudaMemcpy2DAsync(dst, pitch, c->img16.p, row, row, c->H, cudaMemcpyDeviceToHost, c->stream));
  } else if (format == LG_BGRA8_GAMMA) {
    const size_t row = (size_t)c->W * 4;
    if (pitch == 0) pitch = row;
    if (pitch < row) return fail(c, LG_ERR_INVALID, "pitch");
    if ((rc = ensure(c, c->img8, npx * 4))) return rc;
    screenshot_bgra8_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>((const float4 *)c->img.p, (uchar4 *)c->img8.p, npx);
    LG_CUDA(c, cudaGetLastError());
    c->launches++;
    LG_CUDA(c, cudaMemcpy2DAsync(dst, pitch, c->img8.p, row, row, c->H, cudaMemcpyDeviceToHost, c->stream));
  } else if (format == LG_BGRA8_SRGB) {
    const size_t row = (size_t)c->W * 4;
    if (pitch == 0) pitch = row;
    if (pitch < row) return fail(c, LG_ERR_INVALID, "pitch");
    if ((rc = ensure(c, c->img8, npx * 4))) return rc;
    static const SrgbThresholds thresholds = srgb_thresholds();
    surface_bgra8_srgb_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>((const float4 *)c->img.p, (uchar4 *)c->img8.p, npx,
                                                                      thresholds);
    LG_CUDA(c, cudaGetLastError());
    c->launches++;
    LG_CUDA(c, cudaMemcpy2DAsync(dst, pitch, c->img8.p, row, row, c->H, cudaMemcpyDeviceToHost, c->stream));
  } else {
    return fail(c, LG_ERR_INVALID, "format");
  }
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

// ---- display hand-off without the host bounce ----------------------------------------------
namespace {
size_t format_bytes_per_pixel(int32_t format) {
  return format == LG_RGBA32F ? 16 : format == LG_RGBA16F ? 8 : (format == LG_BGRA8_GAMMA || format == LG_BGRA8_SRGB) ? 4 : 0;
}
int vmm_fail(lg_ctx *c, const char *what, CUresult r) {
  return fail(c, LG_ERR_UNSUPPORTED, std::string(what) + " failed (CUresult " + std::to_string((int)r) + ")");
}
} // namespace

int32_t lg_image_export_fd(lg_ctx *c, int32_t format, int32_t *fd, uint64_t *bytes) {
  int rc = need_image(c);
  if (rc) return rc;
  const size_t bpp = format_bytes_per_pixel(format);
  if (!fd || !bytes || bpp == 0) return fail(c, LG_ERR_INVALID, "format / null output");
  if (!vmm_load()) return fail(c, LG_ERR_UNSUPPORTED, g_vmm.why);
  LG_CUDA(c, cudaSetDevice(c->device));
  LG_CUDA(c, cudaFree(0)); // make sure the primary context is current for the driver calls
  ExportBuf &e = c->exported[format];
  const size_t frame = (size_t)c->W * c->H * bpp;
  if (!e.handle) {
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = c->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    CUresult r = g_vmm.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    if (r != CUDA_SUCCESS || gran == 0) return vmm_fail(c, "cuMemGetAllocationGranularity", r);
    ExportBuf n;
    n.size = (frame + gran - 1) / gran * gran;
    if ((r = g_vmm.Create(&n.handle, n.size, &prop, 0)) != CUDA_SUCCESS) return vmm_fail(c, "cuMemCreate (exportable)", r);
    if ((r = g_vmm.AddressReserve(&n.va, n.size, 0, 0, 0)) != CUDA_SUCCESS) {
      g_vmm.Release(n.handle);
      return vmm_fail(c, "cuMemAddressReserve", r);
    }
    if ((r = g_vmm.Map(n.va, n.size, 0, n.handle, 0)) != CUDA_SUCCESS) {
      g_vmm.AddressFree(n.va, n.size);
      g_vmm.Release(n.handle);
      return vmm_fail(c, "cuMemMap", r);
    }
    CUmemAccessDesc acc{};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if ((r = g_vmm.SetAccess(n.va, n.size, &acc, 1)) != CUDA_SUCCESS) {
      export_release(n);
      return vmm_fail(c, "cuMemSetAccess", r);
    }
    e = n;
    LG_CUDA(c, cudaMemsetAsync((void *)e.va, 0, e.size, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  int out = -1;
  CUresult r = g_vmm.Export(&out, e.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS || out < 0) return vmm_fail(c, "cuMemExportToShareableHandle", r);
  *fd = out;
  *bytes = e.size;
  return LG_OK;
}

int32_t lg_image_export_refresh(lg_ctx *c, int32_t format) {
  int rc = need_image(c);
  if (rc) return rc;
  if (format_bytes_per_pixel(format) == 0) return fail(c, LG_ERR_INVALID, "format");
  ExportBuf &e = c->exported[format];
  if (!e.handle) return fail(c, LG_ERR_STATE, "lg_image_export_fd has not been called for this format");
  LG_CUDA(c, cudaSetDevice(c->device));
  const size_t npx = (size_t)c->W * c->H;
  const float4 *src = (const float4 *)c->img.p;
  if (format == LG_RGBA32F) {
    LG_CUDA(c, cudaMemcpyAsync((void *)e.va, src, npx * 16, cudaMemcpyDeviceToDevice, c->stream));
  } else if (format == LG_RGBA16F) {
    if (c->img16_valid) { // the peer-fused reduce already left the finalized frame in img16
      LG_CUDA(c, cudaMemcpyAsync((void *)e.va, c->img16.p, npx * 8, cudaMemcpyDeviceToDevice, c->stream));
    } else {
      finalize_f16_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, (uint2 *)e.va, npx);
      c->launches++;
    }
  } else if (format == LG_BGRA8_GAMMA) {
    screenshot_bgra8_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, (uchar4 *)e.va, npx);
    c->launches++;
  } else {
    static const SrgbThresholds thresholds = srgb_thresholds();
    surface_bgra8_srgb_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, (uchar4 *)e.va, npx, thresholds);
    c->launches++;
  }
  LG_CUDA(c, cudaGetLastError());
  LG_CUDA(c, cudaStreamSynchronize(c->stream)); // the importer may read as soon as this returns
  return LG_OK;
}

int32_t lg_import_fd_read(int32_t device, int32_t fd, uint64_t bytes, void *dst, uint64_t dst_bytes) {
  if (fd < 0 || !dst || dst_bytes > bytes || bytes == 0) return LG_ERR_INVALID;
  if (!vmm_load()) return LG_ERR_UNSUPPORTED;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) {
    cudaGetLastError();
    return LG_ERR_CUDA;
  }
  CUmemGenericAllocationHandle h = 0;
  if (g_vmm.Import(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) return LG_ERR_UNSUPPORTED;
  CUdeviceptr va = 0;
  int32_t rc = LG_OK;
  if (g_vmm.AddressReserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) {
    g_vmm.Release(h);
    return LG_ERR_NOMEM;
  }
  if (g_vmm.Map(va, bytes, 0, h, 0) != CUDA_SUCCESS) {
    rc = LG_ERR_CUDA;
  } else {
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READ;
    if (g_vmm.SetAccess(va, bytes, &acc, 1) != CUDA_SUCCESS ||
        cudaMemcpy(dst, (const void *)va, dst_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaGetLastError();
      rc = LG_ERR_CUDA;
    }
    g_vmm.Unmap(va, bytes);
  }
  g_vmm.AddressFree(va, bytes);
  g_vmm.Release(h);
  return rc;
}

// ---- multi-GPU ---------------------------------------------------------------------------
int32_t lg_comm_unique_id(void *id128) {
  if (!id128) return LG_ERR_INVALID;
  if (!nccl_load()) return LG_ERR_NCCL;
  NcclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return LG_ERR_NCCL;
  std::memcpy(id128, &id, 128);
  return LG_OK;
}

int32_t lg_comm_init_rank(lg_ctx *c, const void *id128, int32_t rank, int32_t world) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(c, LG_ERR_INVALID, "comm arguments");
  if (!nccl_load()) return fail(c, LG_ERR_NCCL, g_nccl.why);
  LG_CUDA(c, cudaSetDevice(c->device));
  if (c->comm) g_nccl.CommDestroy(c->comm), c->comm = nullptr;
  NcclUniqueId id;
  std::memcpy(&id, id128, 128);
  int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) return fail(c, LG_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
  c->comm_rank = rank, c->comm_world = world;
  return LG_OK;
}

int32_t lg_comm_init_all(lg_ctx **ctxs, int32_t n) {
  if (!ctxs || n < 1) return LG_ERR_INVALID;
  if (!nccl_load()) return fail(ctxs[0], LG_ERR_NCCL, g_nccl.why);
  std::vector<int> devs(n);
  std::vector<NcclComm> comms(n);
  for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
  int r = g_nccl.CommInitAll(comms.data(), n, devs.data());
  if (r != 0) return fail(ctxs[0], LG_ERR_NCCL, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
  for (int i = 0; i < n; ++i) {
    ctxs[i]->comm = comms[i];
    ctxs[i]->comm_rank = i, ctxs[i]->comm_world = n;
  }
  return LG_OK;
}

} // extern "C"

namespace {

struct PeerInfo { // what every rank tells the others about its buffers
  cudaIpcMemHandle_t h_img, h_img16, h_flags;
  unsigned long long ptr_img, ptr_img16, ptr_flags, epoch;
  long long pid;
  int device, has16;
  unsigned long long bytes_img;
};

void close_peers(lg_ctx *c) {
  for (int p = 0; p < kMaxPeers; ++p) {
    if (c->peer_opened[p] && c->peer_img[p]) cudaIpcCloseMemHandle(c->peer_img[p]);
    if (c->peer_opened16[p] && c->peer_img16[p]) cudaIpcCloseMemHandle(c->peer_img16[p]);
    c->peer_img[p] = c->peer_img16[p] = nullptr;
    c->peer_opened[p] = c->peer_opened16[p] = false;
  }
  c->peers_ready = false;
  cudaGetLastError();
}

void close_peer_flags(lg_ctx *c) {
  for (int p = 0; p < kMaxPeers; ++p) {
    if (c->peer_flags_opened[p] && c->peer_flags[p]) cudaIpcCloseMemHandle(c->peer_flags[p]);
    c->peer_flags[p] = nullptr;
    c->peer_flags_opened[p] = false;
  }
  c->flags_ready = false;
  cudaGetLastError();
}

// stream-ordered barrier + consensus: max over ranks of `value` (also orders all ranks' earlier work on their streams)
int comm_max(lg_ctx *c, int value, int *out) {
  int rc = ensure(c, c->sync_buf, 16);
  if (rc) return rc;
  LG_CUDA(c, cudaMemcpyAsync(c->sync_buf.p, &value, 4, cudaMemcpyHostToDevice, c->stream));
  int r = g_nccl.AllReduce(c->sync_buf.p, (char *)c->sync_buf.p + 8, 1, kNcclInt32, kNcclMax, c->comm, c->stream);
  if (r != 0) return fail(c, LG_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
  LG_CUDA(c, cudaMemcpyAsync(out, (char *)c->sync_buf.p + 8, 4, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  return LG_OK;
}

// all ranks exchange their buffer handles and map each other's images (collective)
int exchange_peers(lg_ctx *c, int root) {
  close_peers(c);
  close_peer_flags(c);
  const int n = c->comm_world;
  int rc;
  if (c->comm_rank == root && (rc = ensure(c, c->img16, (size_t)c->W * c->H * 8))) return rc;
  if (!c->flags.p) { // allocated once per context, never moved: peers keep their mapping of it
    if ((rc = ensure(c, c->flags, (size_t)kFlagPhases * 16 * 8))) return rc;
    if ((rc = ensure(c, c->flag_status, 16))) return rc;
    LG_CUDA(c, cudaMemsetAsync(c->flags.p, 0, c->flags.bytes, c->stream));
  }
  PeerInfo mine{};
  mine.ptr_flags = (unsigned long long)(uintptr_t)c->flags.p;
  mine.epoch = c->flag_epoch;
  mine.pid = (long long)getpid();
  mine.device = c->device;
  mine.ptr_img = (unsigned long long)(uintptr_t)c->img.p;
  mine.bytes_img = (unsigned long long)c->W * c->H * 16;
  bool ok = cudaIpcGetMemHandle(&mine.h_img, c->img.p) == cudaSuccess;
  ok = ok && cudaIpcGetMemHandle(&mine.h_flags, c->flags.p) == cudaSuccess;
  if (c->comm_rank == root) {
    mine.has16 = 1;
    mine.ptr_img16 = (unsigned long long)(uintptr_t)c->img16.p;
    ok = ok && cudaIpcGetMemHandle(&mine.h_img16, c->img16.p) == cudaSuccess;
  }
  cudaGetLastError();
  if ((rc = ensure(c, c->peer_xchg, sizeof(PeerInfo) * (size_t)(n + 1)))) return rc;
  PeerInfo *d_mine = (PeerInfo *)c->peer_xchg.p, *d_all = d_mine + 1;
  LG_CUDA(c, cudaMemcpyAsync(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
  int r = g_nccl.AllGather(d_mine, d_all, sizeof(PeerInfo), kNcclUint8, c->comm, c->stream);
  if (r != 0) return fail(c, LG_ERR_NCCL, std::string("ncclAllGather: ") + g_nccl.GetErrorString(r));
  std::vector<PeerInfo> all(n);
  LG_CUDA(c, cudaMemcpyAsync(all.data(), d_all, sizeof(PeerInfo) * n, cudaMemcpyDeviceToHost, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  bool mismatch = false;
  for (int p = 0; p < n; ++p)
    if (all[p].bytes_img != mine.bytes_img) mismatch = true, ok = false; // ranks disagree on the image size
  for (int p = 0; p < n && ok; ++p) {
    c->flag_epoch = std::max(c->flag_epoch, all[p].epoch); // every rank continues from the same epoch
    if (p == c->comm_rank) {
      c->peer_img[p] = c->img.p;
      c->peer_flags[p] = c->flags.p;
      if (p == root) c->peer_img16[p] = c->img16.p;
      continue;
    }
    if (all[p].pid == mine.pid) { // same process (lg_comm_init_all): plain peer access
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, c->device, all[p].device) != cudaSuccess || !can) {
        ok = false;
        break;
      }
      static bool enabled[kMaxPeers][kMaxPeers] = {{false}}; // per (this device, peer device): enable once per process
      if (c->device < kMaxPeers && all[p].device < kMaxPeers && !enabled[c->device][all[p].device]) {
        cudaError_t e = cudaDeviceEnablePeerAccess(all[p].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
        else enabled[c->device][all[p].device] = true;
        cudaGetLastError();
      }
      c->peer_img[p] = (void *)(uintptr_t)all[p].ptr_img;
      c->peer_flags[p] = (void *)(uintptr_t)all[p].ptr_flags;
      if (p == root) c->peer_img16[p] = (void *)(uintptr_t)all[p].ptr_img16;
    } else { // another process (one rank per GPU under torchrun): CUDA IPC
      if (cudaIpcOpenMemHandle(&c->peer_img[p], all[p].h_img, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        c->peer_img[p] = nullptr;
        ok = false;
      } else {
        c->peer_opened[p] = true;
      }
      if (ok && cudaIpcOpenMemHandle(&c->peer_flags[p], all[p].h_flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        c->peer_flags[p] = nullptr;
        ok = false;
      } else if (ok) {
        c->peer_flags_opened[p] = true;
      }
      if (ok && p == root) {
        if (cudaIpcOpenMemHandle(&c->peer_img16[p], all[p].h_img16, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          c->peer_img16[p] = nullptr;
          ok = false;
        } else {
          c->peer_opened16[p] = true;
        }
      }
      cudaGetLastError();
    }
  }
  // every rank must take the same path: consensus on "somebody could not map a peer"
  int any_bad = 0;
  if ((rc = comm_max(c, ok ? 0 : (mismatch ? 2 : 1), &any_bad))) return rc;
  c->peer_fused_ok = any_bad == 0;
  c->peer_size_mismatch = any_bad == 2;
  if (!c->peer_fused_ok) close_peers(c), close_peer_flags(c);
  c->flags_ready = c->peer_fused_ok;
  c->peers_ready = true; // the exchange happened (even if it ended in the NCCL fallback)
  c->peers_root = root;
  return LG_OK;
}

} // namespace

extern "C" {

int32_t lg_reduce_mode_set(lg_ctx *c, int32_t mode) {
  if (!c) return LG_ERR_INVALID;
  if (mode < 0 || mode > 2) return fail(c, LG_ERR_INVALID, "reduce mode");
  c->reduce_mode = mode;
  return LG_OK;
}

int32_t lg_image_reduce(lg_ctx *c, int32_t root, float *reduce_ms) {
  int rc = need_image(c);
  if (rc) return rc;
  if (reduce_ms) *reduce_ms = 0.f;
  if (c->comm_world <= 1 || !c->comm) return LG_OK; // one partial image: nothing to sum
  if (root < 0 || root >= c->comm_world) return fail(c, LG_ERR_INVALID, "root");
  if (!c->blend_linear) return fail(c, LG_ERR_UNSUPPORTED, "Min / Max blend states: partial images cannot be summed");
  LG_CUDA(c, cudaSetDevice(c->device));
  LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  bool fused = c->reduce_mode != 1 && c->comm_world <= kMaxPeers;
  const int mine_need = (!c->peers_ready || c->peers_root != root) ? 1 : 0;
  bool started = false; // a barrier in front of the reduce kernel has run (every rank's accumulation is complete)
  if (fused && !c->flags_ready) {
    // first call of this communicator (or after a failed mapping): NCCL is the barrier and carries the consensus on
    // "does anybody need the handle exchange", then the handles travel through ncclAllGather
    int need = 0;
    if ((rc = comm_max(c, mine_need, &need))) return rc;
    if (need && (rc = exchange_peers(c, root))) return rc;
    if (c->peer_size_mismatch) return fail(c, LG_ERR_INVALID, "lg_image_reduce: the ranks' images differ in size");
    fused = c->peer_fused_ok;
    if (!fused && c->reduce_mode == 2) return fail(c, LG_ERR_UNSUPPORTED, "peer memory is not reachable from every rank");
    started = true;
  }
  if (fused) {
    // the steady state: two device-side barriers over peer-mapped flag words around the kernel, no NCCL call and no
    // host synchronisation until the end.  Barrier 0 ("my accumulation is complete", + the need-exchange bit of every
    // rank), the reduce of this rank's band, barrier 1 ("my band has landed in the root's buffers": nobody touches its
    // partial image or reads the frame before that).
    for (int attempt = 0; attempt < 2; ++attempt) {
      PeerFlags F{};
      F.n = c->comm_world, F.rank = c->comm_rank, F.mine = (unsigned long long *)c->flags.p;
      for (int p = 0; p < F.n; ++p) F.peer[p] = (unsigned long long *)c->peer_flags[p];
      const unsigned long long epoch = ++c->flag_epoch;
      const long long timeout = 20ll * 1000 * 1000 * 1000; // ~10 s of SM clocks: a peer that is not in this protocol
      unsigned int *status = (unsigned int *)c->flag_status.p;
      LG_CUDA(c, cudaMemsetAsync(status, 0, 16, c->stream));
      const int need_bit = (attempt == 0 && !started) ? mine_need : 0;
      peer_barrier_kernel<<<1, 32, 0, c->stream>>>(F, 0, epoch, (unsigned)need_bit, status, timeout);
      LG_CUDA(c, cudaGetLastError());
      PeerPtrs P{};
      P.n = c->comm_world;
      for (int p = 0; p < P.n; ++p) P.img[p] = (const float4 *)c->peer_img[p];
      P.root_img = (float4 *)c->peer_img[root];
      P.root_img16 = (uint2 *)c->peer_img16[root];
      const size_t rows0 = (size_t)c->H * c->comm_rank / c->comm_world, rows1 = (size_t)c->H * (c->comm_rank + 1) / c->comm_world;
      const size_t px0 = rows0 * c->W, px1 = rows1 * c->W;
      if (px1 > px0 && P.root_img && P.root_img16) {
        reduce_finalize_peer_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(P, px0, px1, status);
        LG_CUDA(c, cudaGetLastError());
        c->launches++;
      }
      peer_barrier_kernel<<<1, 32, 0, c->stream>>>(F, 1, epoch, 0u, status + 1, timeout);
      LG_CUDA(c, cudaGetLastError());
      c->launches += 2;
      unsigned int *h_status = (unsigned int *)(c->h_totals + 40);
      LG_CUDA(c, cudaMemcpyAsync(h_status, status, 8, cudaMemcpyDeviceToHost, c->stream));
      LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
      LG_CUDA(c, cudaStreamSynchronize(c->stream));
      if ((h_status[0] | h_status[1]) & 2u)
        return fail(c, LG_ERR_NCCL, "lg_image_reduce: a peer did not arrive at the barrier (is every rank calling it?)");
      if (!(h_status[0] & 1u)) break; // done
      // some rank's buffers moved (lg_image_configure): nobody reduced anything; exchange the handles again (collective:
      // every rank saw the same bit) and run the three kernels once more
      if (attempt == 1) return fail(c, LG_ERR_STATE, "lg_image_reduce: handle exchange requested twice");
      if ((rc = exchange_peers(c, root))) return rc;
      if (c->peer_size_mismatch) return fail(c, LG_ERR_INVALID, "lg_image_reduce: the ranks' images differ in size");
      if (!c->peer_fused_ok) return fail(c, LG_ERR_UNSUPPORTED, "peer memory is no longer reachable from every rank");
    }
    if (c->comm_rank == root) c->img16_valid = true;
    if (reduce_ms) LG_CUDA(c, cudaEventElapsedTime(reduce_ms, c->ev0, c->ev1));
    return LG_OK;
  } else {
    // ncclReduce with different counts per rank does not fail, it hangs or corrupts: agree on the size first
    const int px = (int)(((size_t)c->W << 16) ^ (size_t)c->H);
    int hi = 0, lo = 0;
    if ((rc = comm_max(c, px, &hi)) || (rc = comm_max(c, -px, &lo))) return rc;
    if (hi != -lo) return fail(c, LG_ERR_INVALID, "lg_image_reduce: the ranks' images differ in size");
    int r = g_nccl.Reduce(c->img.p, c->img.p, (size_t)c->W * c->H * 4, kNcclFloat32, kNcclSum, root, c->comm, c->stream);
    if (r != 0) return fail(c, LG_ERR_NCCL, std::string("ncclReduce: ") + g_nccl.GetErrorString(r));
    c->img16_valid = false;
  }
  LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  LG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (reduce_ms) LG_CUDA(c, cudaEventElapsedTime(reduce_ms, c->ev0, c->ev1));
  return LG_OK;
}

int32_t lg_comm_destroy(lg_ctx *c) {
  if (!c) return LG_ERR_INVALID;
  cudaSetDevice(c->device);
  close_peers(c);
  close_peer_flags(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  c->comm = nullptr;
  c->comm_world = 1, c->comm_rank = 0;
  return LG_OK;
}

int32_t lg_measure_fma_peak(lg_ctx *c, int32_t precision, int32_t reps, double *tflops) {
  if (!c || !tflops) return LG_ERR_INVALID;
  LG_CUDA(c, cudaSetDevice(c->device));
  DevBuf tmp;
  int rc = ensure(c, tmp, 64);
  if (rc) return rc;
  const int grid = c->sm_count * 8, iters = 4096;
  double best = 0;
  for (int r = 0; r < std::max(1, reps) + 1; ++r) {
    LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    if (precision == LG_PRECISION_F64)
      dfma_peak_kernel<<<grid, 256, 0, c->stream>>>((double *)tmp.p, iters, 0.999, 0.001);
    else
      fma_peak_kernel<<<grid, 256, 0, c->stream>>>((float *)tmp.p, iters, 0.999f, 0.001f);
    LG_CUDA(c, cudaGetLastError());
    c->launches++;
    LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    LG_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double fl = 2.0 * kFmaPerIter * (double)iters * (double)grid * 256.0;
    if (r > 0 && ms > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  release(tmp);
  *tflops = best;
  return LG_OK;
}

int32_t lg_measure_red_peak(lg_ctx *c, uint64_t span_px, int32_t pattern, int32_t reps, double *gred) {
  if (!c || !gred || span_px == 0) return LG_ERR_INVALID;
  LG_CUDA(c, cudaSetDevice(c->device));
  DevBuf tmp;
  int rc = ensure(c, tmp, span_px * 16);
  if (rc) return rc;
  LG_CUDA(c, cudaMemsetAsync(tmp.p, 0, span_px * 16, c->stream));
  const int grid = c->sm_count * 8, iters = 2048;
  double best = 0;
  for (int r = 0; r < std::max(1, reps) + 1; ++r) {
    LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    red_peak_kernel<<<grid, 256, 0, c->stream>>>((float *)tmp.p, span_px, iters, pattern);
    LG_CUDA(c, cudaGetLastError());
    c->launches++;
    LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    LG_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double n = (double)iters * (double)grid * 256.0;
    if (r > 0 && ms > 0) best = std::max(best, n / (ms * 1e-3) / 1e9);
  }
  release(tmp);
  *gred = best;
  return LG_OK;
}

int32_t lg_measure_tile_rmw_peak(lg_ctx *c, int32_t reps, double *gfrag) {
  if (!c || !gfrag) return LG_ERR_INVALID;
  LG_CUDA(c, cudaSetDevice(c->device));
  DevBuf tmp;
  int rc = ensure(c, tmp, 64);
  if (rc) return rc;
  const size_t smem = (size_t)kRasterWarps * ((size_t)kTileFloat4 * sizeof(float4) + sizeof(RasterScratch));
  LG_CUDA(c, cudaFuncSetAttribute(tile_rmw_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  LG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_rmw_peak_kernel, kRasterWarps * 32, smem));
  if (per_sm < 1) return fail(c, LG_ERR_CUDA, "tile microbenchmark does not fit on an SM");
  const int grid = c->sm_count * per_sm, iters = 1 << 15;
  double best = 0;
  for (int r = 0; r < std::max(1, reps) + 1; ++r) {
    LG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    tile_rmw_peak_kernel<<<grid, kRasterWarps * 32, smem, c->stream>>>((float *)tmp.p, iters);
    LG_CUDA(c, cudaGetLastError());
    c->launches++;
    LG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    LG_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double n = 4.0 * (double)iters * (double)grid * (double)(kRasterWarps * 32);
    if (r > 0 && ms > 0) best = std::max(best, n / (ms * 1e-3) / 1e9);
  }
  release(tmp);
  *gfrag = best;
  return LG_OK;
}

int32_t lg_render_overlap_set(lg_ctx *c, int32_t mode, uint32_t waves) {
  if (!c) return LG_ERR_INVALID;
  if (mode < 0 || mode > 2) return fail(c, LG_ERR_INVALID, "overlap mode");
  if (waves == 1 || waves > 4096) return fail(c, LG_ERR_INVALID, "waves");
  c->render_overlap = mode;
  if (waves) c->pipe_waves = waves;
  return LG_OK;
}

int32_t lg_stream_handle(lg_ctx *c, uint64_t *s) {
  if (!c || !s) return LG_ERR_INVALID;
  *s = (uint64_t)(uintptr_t)c->stream;
  return LG_OK;
}
int32_t lg_image_device_ptr(lg_ctx *c, uint64_t *p) {
  if (!c || !p) return LG_ERR_INVALID;
  *p = (uint64_t)(uintptr_t)c->img.p;
  return LG_OK;
}
int32_t lg_host_alloc(size_t bytes, void **out) {
  if (!out || bytes == 0) return LG_ERR_INVALID;
  *out = nullptr;
  if (cudaHostAlloc(out, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return LG_ERR_NOMEM;
  }
  return LG_OK;
}
int32_t lg_host_free(void *p) {
  if (!p) return LG_ERR_INVALID;
  return cudaFreeHost(p) == cudaSuccess ? LG_OK : LG_ERR_CUDA;
}
int32_t lg_launch_count(lg_ctx *c, uint64_t *n) {
  if (!c || !n) return LG_ERR_INVALID;
  *n = c->launches;
  return LG_OK;
}

} // extern "C"
