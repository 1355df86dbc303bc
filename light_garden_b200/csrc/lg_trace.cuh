// lg_trace.cuh — K1 (ray emission) + K2 (trace) for sm_100a.
//
// Replaces Tracer::trace_all / Tracer::trace (src/light_garden/tracer.rs:276-493)
// and Light::set_num_rays (src/light_garden/light.rs:103-115,163-174,225-249).
//
// Shape of the kernel
//   * persistent CTAs; every thread owns R ray slots.  A slot holds one ray of
//     the reference's work list (tracer.rs:368-369).  The reference walks a
//     primary ray's split tree breadth-first with two Vecs; here each slot
//     walks it depth-first with a private stack in global memory — the set of
//     processed rays (and therefore of emitted segments) is identical, the
//     emission order is restored from the (ray, generation, path) tag.
//   * brute force over every object (tracer.rs:412-424) in two phases.  Broad
//     phase: the object table — one conservative bounding circle per object,
//     SoA, staged once per CTA into shared memory with one cp.async.bulk (TMA
//     bulk copy, SASS UBLKCP) — is swept by all lanes in lock step (broadcast
//     LDS.128 of 4 objects, no bank conflicts, no branches): 3 FFMA + 1 funnel
//     shift per ray x object decide "the ray's line misses this object for
//     sure" and collect the outcome in a bit mask.  Narrow phase: the rare
//     survivors run the exact ORACLE.md test.  The broad phase only discards
//     provable misses (its radius carries the rounding margin), so results are
//     bit-identical to testing every object exactly.
//   * rays that die (left the canvas, culled by cutoff_color, generation
//     limit) free their slot; idle slots are re-filled every iteration from the
//     slot's stack or, warp-aggregated (ballot + one atomicAdd per warp), from
//     the global primary-ray counter, so lanes stay packed with live rays
//     however divergent the bounce depth is.
//   * segments leave through warp-aggregated slot allocation (ballot + one
//     atomicAdd per warp) as two 16-byte stores per segment.
#pragma once
#include "../../include/light_garden_b200.h"
#include "lg_geom.cuh"
#include "lg_nearest.cuh"

#ifndef LG_MERGED_CTAS
#define LG_MERGED_CTAS 3 // CTAs per SM the shared-narrow-phase f32 kernel is compiled for
#endif


namespace lg {

constexpr int kTraceBlock = 256;

struct DevLight {
  int32_t kind;
  int32_t _pad;
  unsigned long long first;    // this context's shard of the light: rays first, first + stride, ...
  unsigned long long stride;
  unsigned long long count;    // rays in the shard
  unsigned long long prefix;   // offset of the shard in this context's ray index space
  unsigned long long id_base;  // global id of ray 0 of this light
  double n_rays;
  double n0; // start medium, tracer.rs:280-287
  float color[4];
  double px, py;               // position / segment a
  double ex, ey;               // directional: b - a
  double min_angle, spot_angle, sign; // spot: light.rs:230-243
  double rsign;                // directional: +1 origins a + (i/n)(b - a); -1 (LG_LIGHT_DIRECTIONAL_NEG_R) a - (i/n)(b - a)
};

struct TraceCounters {
  unsigned long long next_ray;   // work counter
  unsigned long long seg_count;  // segments emitted
  unsigned long long ray_steps;  // popped rays that passed the cutoff test
  unsigned int seg_overflow;     // a segment did not fit
  unsigned int stack_overflow;   // a split did not fit the slot's stack
};

template <class T> struct TraceArgs : SceneArgs<T> { // + toks, obj_first/count, delta, grid (lg_nearest.cuh)
  // broad-phase table, contiguous SoA of n_pad entries each: centre x, centre y,
  // (radius + margin)^2, radius + margin.  n_pad is a multiple of 32; padding
  // entries have a negative squared radius and can never become candidates.
  const T *bounds;
  unsigned int bounds_bytes;
  int n_pad;
  const T *obj_n; // refractive index, NaN = no material
  const int *ovl_start, *ovl_list;
  T canvas[8];
  int n_obj;
  unsigned int max_bounce;
  float cutoff[4];
  // rays
  const DevLight *lights;
  int n_lights;
  const LgRay *rays;             // explicit primary rays (lg_trace_rays) or nullptr
  unsigned long long ray_first;  // this launch covers [ray_first, ray_end) of the index space
  unsigned long long ray_end;
  // outputs
  LgSegment *seg;
  LgSegmentTag *tags;
  LgSegmentF64 *seg64;
  unsigned long long seg_cap;
  TraceCounters *ctr;
  uint4 *stack;
  int stack_cap; // entries per slot
};

// ---- stack entry packing -------------------------------------------------------
template <class T> struct StackCodec;
template <> struct StackCodec<float> {
  static constexpr int kVecs = 3;
  static __device__ __forceinline__ void put(uint4 *s, size_t stride, V2<float> o, V2<float> d, float n, float r,
                                             float g, float b, unsigned gen, unsigned long long path) {
    s[0] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(d.x), __float_as_uint(d.y));
    s[stride] = make_uint4(__float_as_uint(n), __float_as_uint(r), __float_as_uint(g), __float_as_uint(b));
    s[2 * stride] = make_uint4(gen, (unsigned)path, (unsigned)(path >> 32), 0u);
  }
  static __device__ __forceinline__ void get(const uint4 *s, size_t stride, V2<float> &o, V2<float> &d, float &n,
                                             float &r, float &g, float &b, unsigned &gen, unsigned long long &path) {
    uint4 a = s[0], c = s[stride], e = s[2 * stride];
    o = {__uint_as_float(a.x), __uint_as_float(a.y)};
    d = {__uint_as_float(a.z), __uint_as_float(a.w)};
    n = __uint_as_float(c.x);
    r = __uint_as_float(c.y);
    g = __uint_as_float(c.z);
    b = __uint_as_float(c.w);
    gen = e.x;
    path = (unsigned long long)e.y | ((unsigned long long)e.z << 32);
  }
};
template <> struct StackCodec<double> {
  static constexpr int kVecs = 4;
  static __device__ __forceinline__ uint2 d2u(double v) {
    long long x = __double_as_longlong(v);
    return make_uint2((unsigned)x, (unsigned)((unsigned long long)x >> 32));
  }
  static __device__ __forceinline__ double u2d(unsigned lo, unsigned hi) {
    return __longlong_as_double((long long)((unsigned long long)lo | ((unsigned long long)hi << 32)));
  }
  static __device__ __forceinline__ void put(uint4 *s, size_t stride, V2<double> o, V2<double> d, double n, float r,
                                             float g, float b, unsigned gen, unsigned long long path) {
    uint2 a = d2u(o.x), c = d2u(o.y), e = d2u(d.x), f = d2u(d.y), h = d2u(n);
    s[0] = make_uint4(a.x, a.y, c.x, c.y);
    s[stride] = make_uint4(e.x, e.y, f.x, f.y);
    s[2 * stride] = make_uint4(h.x, h.y, __float_as_uint(r), __float_as_uint(g));
    s[3 * stride] = make_uint4(__float_as_uint(b), gen, (unsigned)path, (unsigned)(path >> 32));
  }
  static __device__ __forceinline__ void get(const uint4 *s, size_t stride, V2<double> &o, V2<double> &d, double &n,
                                             float &r, float &g, float &b, unsigned &gen, unsigned long long &path) {
    uint4 a = s[0], c = s[stride], e = s[2 * stride], f = s[3 * stride];
    o = {u2d(a.x, a.y), u2d(a.z, a.w)};
    d = {u2d(c.x, c.y), u2d(c.z, c.w)};
    n = u2d(e.x, e.y);
    r = __uint_as_float(e.z);
    g = __uint_as_float(e.w);
    b = __uint_as_float(f.x);
    gen = f.y;
    path = (unsigned long long)f.z | ((unsigned long long)f.w << 32);
  }
};

// ---- K1: ray emission (always f64, like the reference) ---------------------------
__device__ __forceinline__ void unit_from(double x, double y, double &ux, double &uy) {
  // Ray::from_origin -> nalgebra Unit::new_normalize: v / sqrt(x*x + y*y)
  double n = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
  ux = __ddiv_rn(x, n);
  uy = __ddiv_rn(y, n);
}
__device__ __forceinline__ void emit_ray(const DevLight &l, unsigned long long i, double &ox, double &oy, double &dx,
                                         double &dy) {
  const double PI = 3.14159265358979323846;
  if (l.kind == LG_LIGHT_POINT) { // light.rs:163-174
    double f = __ddiv_rn(__dmul_rn(__dmul_rn((double)i, PI), 2.0), l.n_rays);
    double s, c;
    sincos(f, &s, &c);
    ox = l.px;
    oy = l.py;
    unit_from(c, s, dx, dy);
  } else if (l.kind == LG_LIGHT_SPOT) { // light.rs:225-249
    double step = (double)(i + 1ull);
    double angle = __dadd_rn(l.min_angle, __dmul_rn(__ddiv_rn(step, l.n_rays), l.spot_angle));
    double s, c;
    sincos(angle, &s, &c);
    ox = l.px;
    oy = l.py;
    unit_from(__dmul_rn(l.sign, c), __dmul_rn(l.sign, s), dx, dy);
  } else { // directional, light.rs:103-115 (ORACLE.md §6.3)
    double rr = __dmul_rn(l.rsign, __ddiv_rn((double)i, l.n_rays)); // eval_at_r(-(i / n)) under the chosen convention
    ox = __dadd_rn(l.px, __dmul_rn(rr, l.ex));
    oy = __dadd_rn(l.py, __dmul_rn(rr, l.ey));
    unit_from(-l.ey, l.ex, dx, dy);
  }
}

static __global__ void emit_rays_kernel(DevLight l, unsigned long long first, unsigned long long count, LgRay *dst) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  LgRay r;
  emit_ray(l, first + i, r.origin[0], r.origin[1], r.direction[0], r.direction[1]);
  r.color[0] = l.color[0];
  r.color[1] = l.color[1];
  r.color[2] = l.color[2];
  r.color[3] = l.color[3];
  r.refractive_index = l.n0;
  dst[i] = r;
}

template <class T> struct SignBits;
template <> struct SignBits<float> {
  static __device__ __forceinline__ unsigned get(float v) { return __float_as_uint(v); }
};
template <> struct SignBits<double> {
  static __device__ __forceinline__ unsigned get(double v) { return (unsigned)__double2hiint(v); }
};
template <class T> struct Vec4;
template <> struct Vec4<float> {
  typedef float4 type;
};
template <> struct Vec4<double> {
  typedef double4 type;
};

// Broad phase for 4 consecutive table entries and one ray slot: appends 4 "certainly missed" bits to m.
// cr = cross(c - o, d) = cx*dy - cy*dx - cross(o, d); the object can only be hit if cr^2 <= (radius + margin)^2.
template <class T> struct Broad {
  static __device__ __forceinline__ unsigned test4(const typename Vec4<T>::type &x4, const typename Vec4<T>::type &y4,
                                                   const typename Vec4<T>::type &r4, T sdx, T sdy, T nk, unsigned m) {
    const T xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, rs[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const T t = Real<T>::fma(xs[e], sdy, nk);
      const T c = Real<T>::fma(-ys[e], sdx, t);
      const T disc = Real<T>::fma(-c, c, rs[e]);
      m = __funnelshift_l(SignBits<T>::get(disc), m, 1);
    }
    return m;
  }
};
// f32: Blackwell's packed fma.rn.f32x2 (SASS FFMA2) does two tests per instruction.  Each half is an
// IEEE fused multiply-add, so the bits are those of the scalar form.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <> struct Broad<float> {
  static __device__ __forceinline__ unsigned pair(float x0, float x1, float y0, float y1, float r0, float r1, float sdx,
                                                  float sdy, float nk, unsigned m) {
    const unsigned long long t = fma2(pack2(x0, x1), pack2(sdy, sdy), pack2(nk, nk));
    const unsigned long long c = fma2(pack2(-y0, -y1), pack2(sdx, sdx), t);
    float c0, c1;
    unpack2(c, c0, c1);
    const unsigned long long d = fma2(pack2(-c0, -c1), c, pack2(r0, r1));
    float d0, d1;
    unpack2(d, d0, d1);
    m = __funnelshift_l(__float_as_uint(d0), m, 1);
    return __funnelshift_l(__float_as_uint(d1), m, 1);
  }
  static __device__ __forceinline__ unsigned test4(const float4 &x4, const float4 &y4, const float4 &r4, float sdx,
                                                   float sdy, float nk, unsigned m) {
    m = pair(x4.x, x4.y, y4.x, y4.y, r4.x, r4.y, sdx, sdy, nk, m);
    return pair(x4.z, x4.w, y4.z, y4.w, r4.z, r4.w, sdx, sdy, nk, m);
  }
};


// Table access.  With the table in shared memory the loads are issued in the shared state space from a 32-bit
// address computed once per kernel: through the generic pointer the compiler rebuilt the shared window base
// (S2UR SR_CgaCtaId + ULEA) in every 32-object chunk and in every candidate-filter iteration (ncu: ~3 % of the
// kernel's samples).  In global memory (tables larger than shared memory) the generic path is kept.
template <class T, bool kSmem> struct TabLoad {
  typedef typename Vec4<T>::type T4;
  static __device__ __forceinline__ T4 v4(const T *gen, unsigned, int idx) { return *reinterpret_cast<const T4 *>(gen + idx); }
  static __device__ __forceinline__ T s(const T *gen, unsigned, int idx) { return gen[idx]; }
};
template <> struct TabLoad<float, true> {
  static __device__ __forceinline__ float4 v4(const float *, unsigned sh, int idx) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sh + 4u * (unsigned)idx));
    return v;
  }
  static __device__ __forceinline__ float s(const float *, unsigned sh, int idx) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sh + 4u * (unsigned)idx));
    return v;
  }
};
template <> struct TabLoad<double, true> {
  static __device__ __forceinline__ double4 v4(const double *, unsigned sh, int idx) {
    double4 v;
    const unsigned a = sh + 8u * (unsigned)idx;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(a + 16u));
    return v;
  }
  static __device__ __forceinline__ double s(const double *, unsigned sh, int idx) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sh + 8u * (unsigned)idx));
    return v;
  }
};

template <class T> __device__ __forceinline__ bool contains_object(const TraceArgs<T> &A, int obj, V2<T> p) {
  return contains_range(A.toks + A.obj_first[obj], 0, A.obj_count[obj] - 1, p);
}

__device__ __forceinline__ bool culled(float r, float g, float b, float a, const float *cut) {
  // tracer.rs:378-384
  return (r < cut[0] && g < cut[1] && b < cut[2]) || a < cut[3];
}

// ---- K2 ---------------------------------------------------------------------------
// kMerged: one narrow-phase instance shared by the R slots (small scenes: the exact tests dominate, lanes busy with
// different slots run them together, C2 43 vs 78 ms) or one instance per slot (large scenes: the broad phase and the
// candidate filter dominate and the simpler per-slot loops win, C5 114 vs 123 ms).  Same results either way.
// kLarge: the form for scenes of hundreds of objects and more, where the 32-object sweep is nearly all of the work --
// table loads in the shared state space from addresses held in registers and the sweep fully unrolled (C5: -3 %,
// f64 -9 %).  Small scenes (C2, C3) run 3-9 % slower with it (four more live registers in kernels that already spill)
// and keep the generic loads.
template <class T, int R, bool kSmem, bool kGrid = false, bool kMerged = true, bool kLarge = !kMerged>
__global__ void __launch_bounds__(kTraceBlock, (R >= 4 || sizeof(T) == 8) ? 2 : ((kMerged && !kGrid && R == 2) ? LG_MERGED_CTAS : 3))
    trace_kernel(const __grid_constant__ TraceArgs<T> A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const T *tab = A.bounds;
  if (kSmem) {
    // stage the object table: one TMA bulk copy, completion on an mbarrier
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (A.bounds_bytes > 0) {
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(A.bounds_bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst),
                     "l"(A.bounds), "r"(A.bounds_bytes), "r"(bar)
                     : "memory");
      }
      unsigned done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(0u)
                     : "memory");
      }
    }
    tab = reinterpret_cast<const T *>(smem_raw);
  }
  const T *bx = tab;
  const T *by = bx + A.n_pad;
  const T *br2 = by + A.n_pad;
  const T *brb = br2 + A.n_pad;
  typedef typename Vec4<T>::type T4;
  typedef TabLoad<T, kSmem && kLarge> TL;
  // the same four arrays as 32-bit shared-state-space addresses (used when kSmem)
  unsigned sx = 0u;
  if (kSmem && kLarge) // opaque move: keeps the compiler from rematerialising the window base inside the loops
    asm volatile("mov.u32 %0, %1;" : "=r"(sx) : "r"((unsigned)__cvta_generic_to_shared(smem_raw)));
  const unsigned sy = sx + (unsigned)(A.n_pad * sizeof(T));
  const unsigned sr2 = sy + (unsigned)(A.n_pad * sizeof(T));
  const unsigned srb = sr2 + (unsigned)(A.n_pad * sizeof(T));

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  constexpr int kVecs = StackCodec<T>::kVecs;

  // per-slot state (fully unrolled => registers)
  V2<T> o[R], d[R];
  T nmed[R];
  float cr[R], cg[R], cb[R], ca[R];
  unsigned gen[R];
  unsigned long long path[R], rid[R];
  int sp[R];
  bool alive[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    alive[r] = false;
    sp[r] = 0;
    o[r] = {(T)0, (T)0};
    d[r] = {(T)1, (T)0};
    nmed[r] = (T)1;
    cr[r] = cg[r] = cb[r] = ca[r] = 0.f;
    gen[r] = 0;
    path[r] = rid[r] = 0;
  }
  bool exhausted = false;
  unsigned long long steps = 0;

  while (true) {
    // ---- 1. refill idle slots -----------------------------------------------------
    bool any_alive = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!alive[r] && sp[r] > 0) {
        --sp[r];
        const uint4 *s = A.stack + ((size_t)(sp[r] * R + r) * kVecs) * nthreads + tid;
        StackCodec<T>::get(s, nthreads, o[r], d[r], nmed[r], cr[r], cg[r], cb[r], gen[r], path[r]);
        alive[r] = true; // pushed rays already passed the cutoff test
      }
      const bool want = !alive[r] && !exhausted;
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(&A.ctr->next_ray, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (want) {
          const unsigned long long j = A.ray_first + base + __popc(m & lt_mask);
          if (j < A.ray_end) {
            double ox, oy, dx, dy, n0;
            if (A.rays) {
              const LgRay &ry = A.rays[j];
              ox = ry.origin[0], oy = ry.origin[1], dx = ry.direction[0], dy = ry.direction[1];
              cr[r] = ry.color[0], cg[r] = ry.color[1], cb[r] = ry.color[2], ca[r] = ry.color[3];
              n0 = ry.refractive_index;
              rid[r] = j;
            } else {
              int li = 0;
              while (li + 1 < A.n_lights && j >= A.lights[li + 1].prefix) ++li;
              const DevLight &l = A.lights[li];
              const unsigned long long i = l.first + (j - l.prefix) * l.stride;
              emit_ray(l, i, ox, oy, dx, dy);
              cr[r] = l.color[0], cg[r] = l.color[1], cb[r] = l.color[2], ca[r] = l.color[3];
              n0 = l.n0;
              rid[r] = l.id_base + i;
            }
            o[r] = {(T)ox, (T)oy};
            d[r] = {(T)dx, (T)dy};
            nmed[r] = (T)n0;
            gen[r] = 0;
            path[r] = 0;
            // generation 0 is popped like any other ray: cutoff test, tracer.rs:378-384
            alive[r] = A.max_bounce > 0 && !culled(cr[r], cg[r], cb[r], ca[r], A.cutoff);
          } else {
            exhausted = true;
          }
        }
      }
      any_alive |= alive[r];
    }
    if (!__any_sync(0xffffffffu, any_alive || !exhausted)) break;

    // ---- 2. nearest hit over all objects (tracer.rs:412-424) -----------------------
    Best<T> best[R];
    T sdx[R], sdy[R], nk[R], nkd[R], tb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      best[r].d2 = Real<T>::max_value();
      best[r].obj = -1;
      best[r].tok = -1;
      best[r].px = best[r].py = best[r].aux = (T)0;
      tb[r] = Real<T>::max_value();
      if (alive[r]) {
        ++steps;
        sdx[r] = d[r].x, sdy[r] = d[r].y;
        nk[r] = -cross(o[r], d[r]);  // cross(c - o, d) = cx*dy - cy*dx - cross(o, d)
        nkd[r] = -dot(o[r], d[r]);   // dot(c - o, d)   = cx*dx + cy*dy - dot(o, d)
      } else {                       // an idle slot sweeps a line that misses everything
        sdx[r] = sdy[r] = (T)0;
        nk[r] = (T)-1e30;
        nkd[r] = (T)0;
      }
    }
    if (kGrid) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (alive[r]) best[r] = grid_nearest(A, best[r], o[r], d[r]);
    }
    for (int c0 = 0; !kGrid && c0 < A.n_pad; c0 += 32) {
      // broad phase over 32 objects: bit (31 - i) of m[r] = "object c0 + i cannot be hit"
      unsigned m[R];
#pragma unroll
      for (int r = 0; r < R; ++r) m[r] = 0u;
      if (kLarge) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const T4 x4 = TL::v4(bx, sx, c0 + q);
          const T4 y4 = TL::v4(by, sy, c0 + q);
          const T4 r4 = TL::v4(br2, sr2, c0 + q);
#pragma unroll
          for (int r = 0; r < R; ++r) m[r] = Broad<T>::test4(x4, y4, r4, sdx[r], sdy[r], nk[r], m[r]);
        }
      } else {
#pragma unroll 4
        for (int q = 0; q < 32; q += 4) {
          const T4 x4 = *reinterpret_cast<const T4 *>(bx + c0 + q);
          const T4 y4 = *reinterpret_cast<const T4 *>(by + c0 + q);
          const T4 r4 = *reinterpret_cast<const T4 *>(br2 + c0 + q);
#pragma unroll
          for (int r = 0; r < R; ++r) m[r] = Broad<T>::test4(x4, y4, r4, sdx[r], sdy[r], nk[r], m[r]);
        }
      }
      if (kMerged) {
      // narrow phase, ONE instance for all slots: a lane walks the survivors of its slots one after the other
      // (ascending object order within a slot; take() resolves equal distances towards the lower object index, like
      // the strict `<` of the in-order loop, tracer.rs:417), so lanes busy with different slots run the exact
      // test together and the kernel carries one copy of its code.
      unsigned cand[R];
#pragma unroll
      for (int r = 0; r < R; ++r) cand[r] = ~m[r];
      while (true) {
        // the lane's next candidate that survives the range filter, from whichever slot has one
        int s = -1, j = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          while (s < 0 && cand[r]) {
            const int bit = 31 - __clz(cand[r]);
            cand[r] ^= 1u << bit;
            const int jj = c0 + 31 - bit;
            // all hits of object jj have t in [tca - rb, tca + rb]: skip it when that lies
            // behind the origin or beyond the nearest hit found so far
            const T tca = Real<T>::fma(TL::s(bx, sx, jj), sdx[r], Real<T>::fma(TL::s(by, sy, jj), sdy[r], nkd[r]));
            const T rb = TL::s(brb, srb, jj);
            if (jj >= A.n_obj || tca < -rb || tca - rb > tb[r]) continue; // (jj >= n_obj: see the per-slot loop)
            s = r, j = jj;
          }
        }
        if (s < 0) break;
        V2<T> os = o[0], ds = d[0];
        Best<T> bs = best[0];
#pragma unroll
        for (int r = 1; r < R; ++r)
          if (s == r) os = o[r], ds = d[r], bs = best[r];
        const T before = bs.d2;
        bs = narrow_phase(A, bs, j, os, ds);
        if (bs.d2 != before) {
          const T nt = Real<T>::sqrt(bs.d2) * (T)1.000001 + A.delta;
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (s == r) best[r] = bs, tb[r] = nt;
        }
      }
      } else {
      // narrow phase: survivors in ascending object order, so the strict `<` of
      // tracer.rs:417 resolves equal distances exactly like the in-order loop.  Most chunks leave no survivor in
      // any slot of any lane: one branch skips them all (one test per slot cost ~2 % of the kernel's samples each)
      unsigned all_miss = m[0];
#pragma unroll
      for (int r = 1; r < R; ++r) all_miss &= m[r];
      if (all_miss != 0xffffffffu) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        unsigned cand = ~m[r];
        while (cand) {
          const int bit = 31 - __clz(cand);
          cand ^= 1u << bit;
          const int j = c0 + 31 - bit;
          // all hits of object j have t in [tca - rb, tca + rb]: skip it when that lies
          // behind the origin or beyond the nearest hit found so far
          const T tca = Real<T>::fma(TL::s(bx, sx, j), sdx[r], Real<T>::fma(TL::s(by, sy, j), sdy[r], nkd[r]));
          const T rb = TL::s(brb, srb, j);
          // j >= n_obj: a padding entry of the table.  Its squared radius is negative, so a ray never selects it -- except
          // a ray whose direction is NaN (reflected off a zero-radius circle, say): NaN's sign bit says "not missed" for
          // every entry, and the range test below lets NaN through.  Found by compute-sanitizer on the scene of
          // degenerate shapes (a 4-byte read past obj_first[]); such a ray hits nothing, here as in the oracle.
          if (j >= A.n_obj || tca < -rb || tca - rb > tb[r]) continue;
          const T before = best[r].d2;
          best[r] = narrow_phase(A, best[r], j, o[r], d[r]);
          if (best[r].d2 != before) tb[r] = Real<T>::sqrt(best[r].d2) * (T)1.000001 + A.delta;
        }
      }
      }
      }
    }

    // ---- 3. resolve: shade, emit, spawn (tracer.rs:426-489) ------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bool do_emit = false;
      V2<T> ea = o[r], eb = o[r];
      float e_r = cr[r], e_g = cg[r], e_b = cb[r], e_a = ca[r];
      unsigned e_gen = gen[r];
      unsigned long long e_path = path[r];
      int e_hit = -1;
      if (alive[r]) {
        const Best<T> &bh = best[r];
        if (bh.obj >= 0) {
          const V2<T> hp{bh.px, bh.py};
          const V2<T> nrm = hit_normal(A.toks[bh.tok], hp, bh.aux);
          const T mat_n = A.obj_n[bh.obj];
          do_emit = true;
          eb = hp;
          e_hit = bh.obj;
          const unsigned ngen = gen[r] + 1u;
          const bool child_ok = ngen < A.max_bounce;
          if (mat_n == mat_n) { // has a material, tracer.rs:428
            T n2 = (T)1;       // air
            // tracer.rs:431 obj.contains(&ray.get_origin()), sampled at the
            // midpoint of (origin, hit): ORACLE.md §5.2
            const V2<T> mid{(o[r].x + hp.x) * (T)0.5, (o[r].y + hp.y) * (T)0.5};
            if (contains_object(A, bh.obj, mid)) {
              // tracer.rs:432-439: first OTHER object containing the hit point
              const int q0 = A.ovl_start[bh.obj], q1 = A.ovl_start[bh.obj + 1];
              for (int q = q0; q < q1; ++q) {
                const int ix = A.ovl_list[q];
                if (contains_object(A, ix, hp)) {
                  n2 = A.obj_n[ix];
                  break;
                }
              }
            } else {
              n2 = mat_n; // tracer.rs:441
            }
            V2<T> rfl, rfr;
            bool has;
            const T Rf = refract_dir(d[r], nrm, nmed[r], n2, rfl, rfr, has); // tracer.rs:444-450
            const float refl = (float)Rf;                                    // tracer.rs:454
            const float om = 1.f - refl;
            const float ar = cr[r] * refl, ag = cg[r] * refl, ab = cb[r] * refl; // reflected colour
            const float br = cr[r] * om, bg = cg[r] * om, bb = cb[r] * om;       // refracted colour
            const bool live_a = child_ok && !culled(ar, ag, ab, ca[r], A.cutoff);
            const bool live_b = child_ok && has && !culled(br, bg, bb, ca[r], A.cutoff);
            const unsigned long long pa = path[r] << 1, pb = (path[r] << 1) | 1ull;
            if (live_a && live_b) {
              // keep the refracted ray in the slot, park the reflected one
              if (sp[r] < A.stack_cap) {
                uint4 *s = A.stack + ((size_t)(sp[r] * R + r) * kVecs) * nthreads + tid;
                StackCodec<T>::put(s, nthreads, hp, rfl, nmed[r], ar, ag, ab, ngen, pa);
                ++sp[r];
              } else {
                atomicExch(&A.ctr->stack_overflow, 1u);
              }
            }
            if (live_b) {
              o[r] = hp, d[r] = rfr, nmed[r] = n2;
              cr[r] = br, cg[r] = bg, cb[r] = bb;
              gen[r] = ngen, path[r] = pb;
            } else if (live_a) {
              o[r] = hp, d[r] = rfl;
              cr[r] = ar, cg[r] = ag, cb[r] = ab;
              gen[r] = ngen, path[r] = pa;
            } else {
              alive[r] = false;
            }
          } else { // mirror, tracer.rs:473-481 (colour and medium unchanged)
            if (child_ok) {
              d[r] = reflect_dir(d[r], nrm);
              o[r] = hp;
              gen[r] = ngen;
              path[r] = path[r] << 1;
            } else {
              alive[r] = false;
            }
          }
        } else { // canvas, tracer.rs:482-488
          CandList<T> hl;
          hl.n = 0;
          hit_rect(A.canvas, o[r], d[r], hl);
          if (hl.n > 0) {
            int f = 0;
            for (int q = 1; q < hl.n; ++q)
              if (hl.h[q].t < hl.h[f].t) f = q;
            do_emit = true;
            eb = hl.h[f].p;
          }
          alive[r] = false;
        }
      }
      // warp-aggregated segment slot allocation, two 16-byte stores per segment
      const unsigned em = __ballot_sync(0xffffffffu, do_emit);
      if (em) {
        const int leader = __ffs(em) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(&A.ctr->seg_count, (unsigned long long)__popc(em));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (do_emit) {
          const unsigned long long slot = base + __popc(em & lt_mask);
          if (slot < A.seg_cap) {
            float4 *dst = reinterpret_cast<float4 *>(A.seg + slot);
            dst[0] = make_float4((float)ea.x, (float)ea.y, (float)eb.x, (float)eb.y); // `as f32`, sub_render_pass.rs:192
            dst[1] = make_float4(e_r, e_g, e_b, e_a);
            if (A.tags) {
              LgSegmentTag tg;
              tg.ray = rid[r];
              tg.path = e_path;
              tg.generation = e_gen;
              tg.hit_object = e_hit;
              A.tags[slot] = tg;
            }
            if (A.seg64) {
              LgSegmentF64 s64;
              s64.a[0] = (double)ea.x, s64.a[1] = (double)ea.y, s64.b[0] = (double)eb.x, s64.b[1] = (double)eb.y;
              A.seg64[slot] = s64;
            }
          } else {
            atomicExch(&A.ctr->seg_overflow, 1u);
          }
        }
      }
    }
  }
  // ray_steps: warp reduce, one atomic per warp
  for (int off = 16; off > 0; off >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, off);
  if (lane == 0 && steps) atomicAdd(&A.ctr->ray_steps, steps);
}

// Type-erased access to the instantiations (one translation unit per precision,
// lg_trace_f32.cu / lg_trace_f64.cu, so they compile in parallel).
const void *trace_kernel_f32(int slots, bool smem);
const void *trace_kernel_f64(int slots, bool smem);
const void *trace_kernel_f32_per_slot(bool smem);
const void *trace_kernel_f64_large(int slots, bool smem);
const void *trace_kernel_grid_f32(int slots);
const void *trace_kernel_grid_f64(int slots);

} // namespace lg
