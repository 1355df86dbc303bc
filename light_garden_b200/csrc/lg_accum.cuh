// lg_accum.cuh — K3 (string-mod chords), K4 (line accumulation), K5 (finalize).
//
// Replaces, for the LineList pass of the reference:
//   SubRenderPass::update_vertex_buffer  src/sub_render_pass.rs:188-203  (P2 f64 -> [f32;2])
//   vs_main / fs_main                    src/shader.wgsl:14-28            (ortho transform, colour pass-through)
//   Renderer::generate_matrix            src/renderer.rs:120-124          (OPENGL_TO_WGPU * ortho(-a,a,-1,1,0,1))
//   pipeline state                       src/sub_render_pass.rs:43-101    (LineList, 1 px, no MSAA)
//   default BlendState                   src/light_garden/mod.rs:57-73    (rgb: src+dst, a: src.a*src.a + dst.a)
//   Rgba16Float target + clear           src/texture_renderer.rs:5,69-80, src/renderer.rs:174-177
//   StringMod::draw                      src/light_garden/string_mod.rs:33-158 (Circle curve)
// Coverage rule and arithmetic order: oracle/ORACLE.md §8 (one fragment per
// major-axis pixel centre inside the segment, colour lerped between the ends).
//
// The working image is RGBA fp32 (16 B/pixel); one blended fragment is one
// 16-byte vector reduction (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4)
// resolved in L2.  Compiled with -fmad=false: the only fused operations are the
// explicit fmaf calls the spec names.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/light_garden_b200.h"
#include "lg_srgb.h"

namespace lg {

constexpr int kAccumBlock = 256;

// non-default blend states (lg_blend_set): per component a source factor and how the product meets the image
struct BlendCfg {
  int color_factor, alpha_factor; // LG_BF_* (source-only factors)
  int color_op, alpha_op;         // LG_BO_ADD, LG_BO_REVERSE_SUBTRACT, LG_BO_MIN, LG_BO_MAX
  float constant[4];
};

struct AccumArgs {
  float *img; // RGBA fp32, row-major, y down
  int W, H;
  float m00, m11, hw, hh; // projection + viewport (ORACLE.md §8.1)
  unsigned long long *pixel_updates;
  BlendCfg blend;         // read by the <kBlend = true> kernels only
};

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- generic blend (ORACLE.md §8.6): source factor, then Add / ReverseSubtract / Min / Max against the image ----
__device__ __forceinline__ float blend_factor(int f, float c, float alpha, float constant) {
  switch (f) {
  case LG_BF_ZERO: return 0.f;
  case LG_BF_ONE: return 1.f;
  case LG_BF_SRC: return c;
  case LG_BF_ONE_MINUS_SRC: return 1.f - c;
  case LG_BF_SRC_ALPHA: return alpha;
  case LG_BF_ONE_MINUS_SRC_ALPHA: return 1.f - alpha;
  case LG_BF_CONSTANT: return constant;
  default: return 1.f - constant; // LG_BF_ONE_MINUS_CONSTANT
  }
}
// float min / max on a word that always holds a float: order-preserving integer views of the two sign classes
__device__ __forceinline__ void atomic_max_f32(float *p, float v) {
  if (v != v) return;
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(p), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(p), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float *p, float v) {
  if (v != v) return;
  if (v >= 0.f) atomicMin(reinterpret_cast<int *>(p), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int *>(p), __float_as_uint(v));
}
__device__ __forceinline__ void blend_channel(float *p, int op, float src, float factor) {
  if (op == LG_BO_MIN) atomic_min_f32(p, src);          // wgpu: Min / Max ignore the factors
  else if (op == LG_BO_MAX) atomic_max_f32(p, src);
  else if (op == LG_BO_ADD) atomicAdd(p, src * factor); // dst * One + src * factor
  else atomicAdd(p, -(src * factor));                   // ReverseSubtract: dst * One - src * factor
}
__device__ __forceinline__ void blend_fragment(const BlendCfg &B, float *px, float c0, float c1, float c2, float c3) {
  blend_channel(px + 0, B.color_op, c0, blend_factor(B.color_factor, c0, c3, B.constant[0]));
  blend_channel(px + 1, B.color_op, c1, blend_factor(B.color_factor, c1, c3, B.constant[1]));
  blend_channel(px + 2, B.color_op, c2, blend_factor(B.color_factor, c2, c3, B.constant[2]));
  blend_channel(px + 3, B.alpha_op, c3, blend_factor(B.alpha_factor, c3, c3, B.constant[3]));
}

// Per-segment raster setup (ORACLE.md §8.1-8.2), computed by ONE lane for its own segment and
// parked in shared memory; the warp then walks the 32 parked segments and every lane takes one
// major-axis step of the current one.  A segment with nothing to draw has i0 >= i1.
struct RasterSetup {
  float m0, inv, dn, n0; // major start, 1/(m1 - m0), minor delta, minor start
  int i0, i1, xmajor, _pad;
};

__device__ __forceinline__ RasterSetup raster_setup(const AccumArgs &A, float ax, float ay, float bx, float by) {
  RasterSetup S;
  S.m0 = S.inv = S.dn = S.n0 = 0.f;
  S.i0 = S.i1 = 0;
  S.xmajor = 1;
  S._pad = 0;
  // vs_main + viewport
  const float x0 = __fmaf_rn(A.m00 * ax, A.hw, A.hw), y0 = __fmaf_rn(-(A.m11 * ay), A.hh, A.hh);
  const float x1 = __fmaf_rn(A.m00 * bx, A.hw, A.hw), y1 = __fmaf_rn(-(A.m11 * by), A.hh, A.hh);
  const float dx = x1 - x0, dy = y1 - y0;
  if (!(fabsf(dx) < 1e30f) || !(fabsf(dy) < 1e30f)) return S;
  const bool xmajor = fabsf(dx) >= fabsf(dy);
  const float m0 = xmajor ? x0 : y0, m1 = xmajor ? x1 : y1;
  const float n0 = xmajor ? y0 : x0, n1 = xmajor ? y1 : x1;
  const float dm = m1 - m0;
  if (dm == 0.f) return S;
  const float lo = m0 < m1 ? m0 : m1, hi = m0 < m1 ? m1 : m0;
  const int Nmaj = xmajor ? A.W : A.H;
  float flo = ceilf(lo - 0.5f), fhi = ceilf(hi - 0.5f);
  if (flo < 0.f) flo = 0.f;
  if (fhi > (float)Nmaj) fhi = (float)Nmaj;
  if (!(flo < fhi)) return S;
  S.m0 = m0;
  S.inv = __fdiv_rn(1.0f, dm);
  S.dn = n1 - n0;
  S.n0 = n0;
  S.i0 = (int)flo;
  S.i1 = (int)fhi;
  S.xmajor = xmajor ? 1 : 0;
  return S;
}

// One warp rasterises one parked segment; returns this lane's number of blended fragments.
template <bool kLerp, bool kBlend = false>
__device__ __forceinline__ unsigned raster_walk(const AccumArgs &A, const RasterSetup &S, const float4 ca,
                                                const float4 dc, unsigned lane) {
  const int Nmin = S.xmajor ? A.H : A.W;
  unsigned n = 0;
  for (int i = S.i0 + (int)lane; i < S.i1; i += 32) {
    const float mc = (float)i + 0.5f;
    const float s = (mc - S.m0) * S.inv;
    const float nv = __fmaf_rn(s, S.dn, S.n0);
    const float fj = floorf(nv);
    if (!(fj >= 0.f) || !(fj < (float)Nmin)) continue;
    const int j = (int)fj;
    const int px = S.xmajor ? i : j, py = S.xmajor ? j : i;
    float c0 = ca.x, c1 = ca.y, c2 = ca.z, c3 = ca.w;
    if (kLerp) { // fmaf(s, 0, c) == c: single-colour segments skip the arithmetic, not the semantics
      c0 = __fmaf_rn(s, dc.x, ca.x), c1 = __fmaf_rn(s, dc.y, ca.y);
      c2 = __fmaf_rn(s, dc.z, ca.z), c3 = __fmaf_rn(s, dc.w, ca.w);
    }
    // blend: rgb = src*1 + dst*1 ; a = src.a*src.a + dst.a   (mod.rs:57-73)
    if (kBlend) blend_fragment(A.blend, A.img + ((size_t)py * A.W + px) * 4, c0, c1, c2, c3);
    else red_add_v4(A.img + ((size_t)py * A.W + px) * 4, c0, c1, c2, c3 * c3);
    ++n;
  }
  return n;
}

struct WarpScratch { // one per warp, in shared memory
  RasterSetup s[32];
  float4 ca[32];
  float4 dc[32];
};
constexpr int kWarpsPerBlock = kAccumBlock / 32;

__device__ __forceinline__ void flush_count(const AccumArgs &A, unsigned long long n, unsigned lane) {
  for (int off = 16; off > 0; off >>= 1) n += __shfl_down_sync(0xffffffffu, n, off);
  if (lane == 0 && n) atomicAdd(A.pixel_updates, n);
}

// walk the m parked segments of this warp; the next segment's parameters are fetched from shared
// memory while the current one is being rasterised (hides the LDS latency between short segments)
template <bool kLerp, bool kBlend = false>
__device__ __forceinline__ unsigned long long walk_parked(const AccumArgs &A, const WarpScratch &W, int m,
                                                          unsigned lane) {
  unsigned long long cnt = 0;
  if (m <= 0) return 0;
  RasterSetup S = W.s[0]; // broadcast loads
  float4 ca = W.ca[0];
  float4 dc = kLerp ? W.dc[0] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < m; ++k) {
    const int kn = k + 1 < m ? k + 1 : k;
    const RasterSetup Sn = W.s[kn];
    const float4 can = W.ca[kn];
    const float4 dcn = kLerp ? W.dc[kn] : dc;
    if (S.i0 < S.i1) cnt += raster_walk<kLerp, kBlend>(A, S, ca, dc, lane); // warp-uniform condition
    S = Sn, ca = can, dc = dcn;
  }
  return cnt;
}

// K4 over the compact device segments written by the trace kernel
template <bool kBlend>
__global__ void __launch_bounds__(kAccumBlock) accumulate_segments_kernel(AccumArgs A, const LgSegment *seg,
                                                                           unsigned long long n) {
  __shared__ WarpScratch scratch[kWarpsPerBlock];
  WarpScratch &W = scratch[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long cnt = 0;
  // the next batch's 32-byte segment is loaded while the current batch is rasterised
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), c = p;
  unsigned long long base = warp * 32ull;
  if (base + lane < n) {
    const float4 *s = reinterpret_cast<const float4 *>(seg + base + lane);
    p = __ldg(s), c = __ldg(s + 1);
  }
  for (; base < n; base += nwarps * 32ull) {
    if (base + lane < n) { // one raster setup per lane
      W.s[lane] = raster_setup(A, p.x, p.y, p.z, p.w);
      W.ca[lane] = c;
    }
    const unsigned long long nb = base + nwarps * 32ull;
    if (nb + lane < n) {
      const float4 *s = reinterpret_cast<const float4 *>(seg + nb + lane);
      p = __ldg(s), c = __ldg(s + 1);
    }
    __syncwarp();
    const int m = (int)((n - base) < 32ull ? (n - base) : 32ull);
    cnt += walk_parked<false, kBlend>(A, W, m, lane);
    __syncwarp();
  }
  flush_count(A, cnt, lane);
}

// K4 over host supplied vertex pairs (two colours, f64 positions cast `as f32`)
template <bool kBlend>
__global__ void __launch_bounds__(kAccumBlock) accumulate_pairs_kernel(AccumArgs A, const LgVertexPair *vp,
                                                                        unsigned long long n) {
  __shared__ WarpScratch scratch[kWarpsPerBlock];
  WarpScratch &W = scratch[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long cnt = 0;
  for (unsigned long long base = warp * 32ull; base < n; base += nwarps * 32ull) {
    if (base + lane < n) {
      const LgVertexPair &s = vp[base + lane];
      W.s[lane] = raster_setup(A, (float)s.a[0], (float)s.a[1], (float)s.b[0], (float)s.b[1]);
      const float4 ca = make_float4(s.color_a[0], s.color_a[1], s.color_a[2], s.color_a[3]);
      W.ca[lane] = ca;
      W.dc[lane] = make_float4(s.color_b[0] - ca.x, s.color_b[1] - ca.y, s.color_b[2] - ca.z, s.color_b[3] - ca.w);
    }
    __syncwarp();
    const int m = (int)((n - base) < 32ull ? (n - base) : 32ull);
    cnt += walk_parked<true, kBlend>(A, W, m, lane);
    __syncwarp();
  }
  flush_count(A, cnt, lane);
}

// ---- K3: string mod, fused into K4 (no chord list in memory) -----------------------
struct StringModArgs {
  LgStringMod sm;
  const LgModRemColor *rules;
  unsigned int n_rules;
  unsigned long long first, count;
};

__device__ __forceinline__ unsigned long long wrapping_pow(unsigned long long base, unsigned int e) {
  unsigned long long acc = 1ull; // Rust u64::pow in release builds wraps
  while (e) {
    if (e & 1u) acc *= base;
    base *= base;
    e >>= 1;
  }
  return acc;
}
__device__ __forceinline__ unsigned long long sm_target(const LgStringMod &sm, unsigned long long i) {
  const unsigned long long m = sm.modulo; // string_mod.rs:111-116
  switch (sm.mode) {
  case LG_SM_ADD: return (i + sm.num) % m;
  case LG_SM_MUL: return (i * sm.num) % m;
  case LG_SM_POW: return wrapping_pow(i, (unsigned int)sm.num) % m;
  default: return wrapping_pow(sm.num, (unsigned int)i) % m;
  }
}
// StringMod::init_points (string_mod.rs:33-85) in f64 like the reference
__device__ __forceinline__ void sm_point64(const LgStringMod &sm, unsigned long long n, double &x, double &y);
// ... then `as f32` (sub_render_pass.rs:192)
__device__ __forceinline__ void sm_point(const LgStringMod &sm, unsigned long long n, float &x, float &y) {
  double px, py;
  sm_point64(sm, n, px, py);
  x = (float)px;
  y = (float)py;
}
__device__ __forceinline__ void sm_point64(const LgStringMod &sm, unsigned long long n, double &x, double &y) {
  const double TAU = 6.28318530717958647692;
  const unsigned long long tn = sm.turns * n; // u64 product, wraps
  double px, py;
  if (sm.curve == LG_CURVE_COMPLEX_EXP) {
    // complex.powu((turns * n) as u32): num_traits::pow, exponentiation by squaring with plain complex products
    unsigned int e = (unsigned int)tn;
    double br = sm.curve_p[0], bi = sm.curve_p[1];
    if (e == 0) {
      px = 1.0, py = 0.0;
    } else {
      while ((e & 1u) == 0u) {
        const double r = __dsub_rn(__dmul_rn(br, br), __dmul_rn(bi, bi)), i = __dadd_rn(__dmul_rn(br, bi), __dmul_rn(bi, br));
        br = r, bi = i;
        e >>= 1;
      }
      double ar = br, ai = bi;
      while (e > 1u) {
        e >>= 1;
        const double r = __dsub_rn(__dmul_rn(br, br), __dmul_rn(bi, bi)), i = __dadd_rn(__dmul_rn(br, bi), __dmul_rn(bi, br));
        br = r, bi = i;
        if (e & 1u) {
          const double cr = __dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi)), ci = __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br));
          ar = cr, ai = ci;
        }
      }
      px = ar, py = ai;
    }
  } else {
    const double angle = __ddiv_rn(__dmul_rn((double)tn, TAU), (double)sm.modulo);
    if (sm.curve == LG_CURVE_HYPOTROCHOID) { // string_mod.rs:56-71
      const double small_r = (double)(unsigned long long)sm.curve_p[0], big_r = (double)(unsigned long long)sm.curve_p[1];
      const double off = (double)(unsigned long long)sm.curve_p[2];
      const double smr = __dsub_rn(big_r, small_r), ratio = __dadd_rn(smr, off);
      const double inner = __ddiv_rn(__dmul_rn(angle, smr), small_r);
      const double xx = __dadd_rn(__dmul_rn(smr, cos(angle)), __dmul_rn(off, cos(inner)));
      const double yy = __dsub_rn(__dmul_rn(smr, sin(angle)), __dmul_rn(off, sin(inner)));
      px = __ddiv_rn(xx, ratio), py = __ddiv_rn(yy, ratio);
    } else if (sm.curve == LG_CURVE_LISSAJOUS) { // string_mod.rs:73-83
      const double a = (double)(unsigned long long)sm.curve_p[0], b = (double)(unsigned long long)sm.curve_p[1];
      px = sin(__dadd_rn(__dmul_rn(a, angle), sm.curve_p[2]));
      py = sin(__dmul_rn(b, angle));
    } else { // Circle, string_mod.rs:45-55
      double s, c;
      sincos(angle, &s, &c);
      px = c, py = s;
    }
  }
  x = px;
  y = py;
}
__device__ __forceinline__ void sm_color(const StringModArgs &S, unsigned long long ix, float out[4]) {
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f; // string_mod.rs:124-150
  int cnt = 0;
  for (unsigned int k = 0; k < S.n_rules; ++k) {
    const LgModRemColor r = S.rules[k];
    if (r.modulo != 0ull && (ix % r.modulo) == r.rem) {
      c0 += r.color[0], c1 += r.color[1], c2 += r.color[2], c3 += r.color[3];
      ++cnt;
    }
  }
  if (cnt == 0) {
    out[0] = S.sm.color[0], out[1] = S.sm.color[1], out[2] = S.sm.color[2], out[3] = S.sm.color[3];
  } else {
    const float f = (float)cnt;
    out[0] = __fdiv_rn(c0, f), out[1] = __fdiv_rn(c1, f), out[2] = __fdiv_rn(c2, f), out[3] = __fdiv_rn(c3, f);
  }
}

template <bool kBlend>
__global__ void __launch_bounds__(kAccumBlock) string_mod_kernel(AccumArgs A, StringModArgs S) {
  __shared__ WarpScratch scratch[kWarpsPerBlock];
  WarpScratch &W = scratch[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long cnt = 0;
  for (unsigned long long base = warp * 32ull; base < S.count; base += nwarps * 32ull) {
    if (base + lane < S.count) { // lane = one chord: end points (f64 sincos), colours, raster setup
      const unsigned long long iix = S.first + base + lane;
      const unsigned long long ix = sm_target(S.sm, iix);
      float ax, ay, bx, by, ca[4], cb[4];
      sm_point(S.sm, iix, ax, ay);
      sm_point(S.sm, ix, bx, by);
      sm_color(S, iix, ca);
      sm_color(S, ix, cb);
      W.s[lane] = raster_setup(A, ax, ay, bx, by);
      W.ca[lane] = make_float4(ca[0], ca[1], ca[2], ca[3]);
      W.dc[lane] = make_float4(cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2], cb[3] - ca[3]);
    }
    __syncwarp();
    const int m = (int)((S.count - base) < 32ull ? (S.count - base) : 32ull);
    cnt += walk_parked<true, kBlend>(A, W, m, lane);
    __syncwarp();
  }
  flush_count(A, cnt, lane);
}

// ---- clear + K5 finalize -------------------------------------------------------------
__global__ void clear_image_kernel(float4 *img, size_t n_px, float alpha) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride)
    img[i] = make_float4(0.f, 0.f, 0.f, alpha); // LoadOp::Clear(BLACK), renderer.rs:174-177
}

// fp32 RGBA -> Rgba16Float (round to nearest even, what the ROP store does)
__global__ void finalize_f16_kernel(const float4 *img, uint2 *dst, size_t n_px) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
    const float4 v = img[i];
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned int *>(&lo);
    o.y = *reinterpret_cast<const unsigned int *>(&hi);
    dst[i] = o;
  }
}

// Screenshot conversion of the reference (src/renderer.rs:313-328): every Rgba16Float channel becomes
// `(f.powf(1. / 2.2) * 255.) as u8` (saturating cast, NaN -> 0) and the pixel is stored as [b, g, r, a].
// powf is evaluated in f64 and rounded once, which reproduces a correctly rounded f32 powf.
__device__ __forceinline__ unsigned char f16_to_u8(float v) {
  const float f = __half2float(__float2half_rn(v)); // the value the fp16 texture holds
  const float g = (float)pow((double)f, (double)(1.f / 2.2f)) * 255.f;
  if (!(g == g)) return 0;
  if (g <= 0.f) return 0;
  if (g >= 255.f) return 255;
  return (unsigned char)g; // truncation, as `as u8`
}
__global__ void screenshot_bgra8_kernel(const float4 *img, uchar4 *dst, size_t n_px) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
    const float4 v = img[i];
    dst[i] = make_uchar4(f16_to_u8(v.z), f16_to_u8(v.y), f16_to_u8(v.x), f16_to_u8(v.w));
  }
}

// The 8-bit surface target (render_to_texture off, src/sub_render_pass.rs:59-63; screenshot format
// Bgra8UnormSrgb, src/renderer.rs:207-209), ORACLE.md 8.7.  A unorm target saturates at 1; with non-negative
// additive fragments the order-free limit of that blend is min(sum, 1).  Colour is stored sRGB-encoded and rounded
// to nearest: byte = #{k in 1..255 : T[k] <= c}, T[k] = the linear value whose encoding is (k - 0.5) / 255
// (f64 on the host, rounded to f32) -- a table look-up, so the bytes do not depend on any device pow().
__global__ void surface_bgra8_srgb_kernel(const float4 *img, uchar4 *dst, size_t n_px,
                                          const __grid_constant__ SrgbThresholds T) {
  __shared__ float t[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) t[i] = T.t[i];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
    const float4 v = img[i];
    dst[i] = make_uchar4(srgb_byte(t, v.z), srgb_byte(t, v.y), srgb_byte(t, v.x), unorm_byte(v.w));
  }
}

} // namespace lg
