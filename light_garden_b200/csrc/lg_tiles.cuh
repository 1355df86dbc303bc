// lg_tiles.cuh — K4', the tile-binned form of the line accumulation.
//
// Same semantics as lg_accum.cuh (ORACLE.md §8; reference: src/sub_render_pass.rs:188-212, src/shader.wgsl,
// blend src/light_garden/mod.rs:57-73): what changes is where the additive blend is resolved.  The direct kernels
// send one 16-byte red.global.add per fragment to L2; at 500..70 000 fragments per pixel (BASELINE configs C1..C5)
// that is the bottleneck.  Here fragments are summed in SHARED MEMORY first:
//
//   1. tile_count : lane = segment; walk the 32x32-pixel tiles the segment's fragments fall into and count them in
//                   a per-CTA SHARED-MEMORY histogram (hot tiles next to a light take millions of hits: global
//                   atomics on one address would serialise), written out as hist[cta][tile]
//   2. tile_rowscan + tile_scan: exclusive scans -> every CTA's write position in every tile's list; work items =
//                   (tile, chunk of <= kChunk list entries)
//   3. tile_fill  : same walk by the same CTA over the same segments, positions from a shared-memory cursor
//   4. tile_raster: persistent warps; a warp owns a PRIVATE 32x32 RGBA fp32 tile in shared memory (33-pixel pitch:
//                   conflict-free for x-major and y-major lines), adds every fragment of its work item with plain
//                   LDS.128 / FADD / STS.128 — lanes are distinct major-axis steps of one segment, so there are no
//                   collisions and no shared-memory atomics (those are CAS spin loops for float) — and finally
//                   flushes the touched pixels with one red.global.add.v4.f32 each.
//
// Fragment coordinates use exactly the arithmetic of raster_walk(); a fragment lands in the tile that contains its
// pixel, every (segment, tile) pair that can own a fragment is listed (the walk brackets the minor coordinate of a
// clip interval by its two end fragments: j(i) is monotone), so coverage is identical to the direct kernels.
#pragma once
#include "lg_accum.cuh"

namespace lg {

constexpr int kTile = 32;          // tile edge in pixels
constexpr int kTileShift = 5;
constexpr int kTilePitch = 33;     // float4 per tile row in shared memory
constexpr int kChunk = 2048;       // list entries per work item
constexpr int kRasterWarps = 4;    // warps per CTA of tile_raster_kernel
constexpr int kTileFloat4 = kTile * kTilePitch;

// device segment with two end colours (string-mod chords, host vertex pairs)
struct Seg2 {
  float4 ab; // a.x a.y b.x b.y
  float4 ca, cb;
};

struct TileArgs {
  AccumArgs A;
  int tiles_x, tiles_y, n_tiles;
  unsigned int *tile_count;   // [n_tiles]
  unsigned int *tile_cursor;  // [n_tiles]
  unsigned long long *tile_offset; // [n_tiles + 1]
  unsigned int *item_prefix;  // [n_tiles + 1]
  unsigned long long *totals; // [0] = pairs, [1] = items
  unsigned int *list;         // segment index per pair
  unsigned int *item_counter;
  unsigned int *hist;         // [n_ctas][n_tiles] per-CTA counts, then per-CTA exclusive offsets within a tile
  int n_ctas;
};

template <class Seg> struct SegIO;
template <> struct SegIO<LgSegment> {
  static constexpr bool kLerp = false;
  static __device__ __forceinline__ void load(const LgSegment *s, unsigned long long i, float4 &ab, float4 &ca, float4 &dc) {
    const float4 *p = reinterpret_cast<const float4 *>(s + i);
    ab = __ldg(p);
    ca = __ldg(p + 1);
    dc = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  static __device__ __forceinline__ float4 load_ab(const LgSegment *s, unsigned long long i) {
    return __ldg(reinterpret_cast<const float4 *>(s + i));
  }
};
template <> struct SegIO<Seg2> {
  static constexpr bool kLerp = true;
  static __device__ __forceinline__ void load(const Seg2 *s, unsigned long long i, float4 &ab, float4 &ca, float4 &dc) {
    const float4 *p = reinterpret_cast<const float4 *>(s + i);
    ab = __ldg(p);
    ca = __ldg(p + 1);
    const float4 cb = __ldg(p + 2);
    dc = make_float4(cb.x - ca.x, cb.y - ca.y, cb.z - ca.z, cb.w - ca.w);
  }
  static __device__ __forceinline__ float4 load_ab(const Seg2 *s, unsigned long long i) {
    return __ldg(reinterpret_cast<const float4 *>(s + i));
  }
};

// minor pixel coordinate of the fragment at major pixel i (the arithmetic of raster_walk)
__device__ __forceinline__ float frag_minor(const RasterSetup &S, int i) {
  const float mc = (float)i + 0.5f;
  const float s = (mc - S.m0) * S.inv;
  return floorf(__fmaf_rn(s, S.dn, S.n0));
}

// calls f(tile index) for every tile that can own a fragment of the segment
template <class F> __device__ __forceinline__ void for_each_tile(const RasterSetup &S, int W, int H, int tiles_x, F f) {
  if (S.i0 >= S.i1) return;
  const int Nmin = S.xmajor ? H : W;
  const float fmax = (float)(Nmin - 1);
  for (int t = S.i0 >> kTileShift; t <= (S.i1 - 1) >> kTileShift; ++t) {
    const int a = max(S.i0, t << kTileShift), b = min(S.i1, (t + 1) << kTileShift) - 1;
    const float ja = frag_minor(S, a), jb = frag_minor(S, b);
    float lo = fminf(ja, jb), hi = fmaxf(ja, jb);
    if (!(hi >= 0.f) || !(lo <= fmax)) continue; // every fragment of this stretch is off the canvas (or NaN)
    lo = fmaxf(lo, 0.f), hi = fminf(hi, fmax);
    for (int u = (int)lo >> kTileShift; u <= (int)hi >> kTileShift; ++u) f(S.xmajor ? u * tiles_x + t : t * tiles_x + u);
  }
}

// segments [lo, hi) of CTA b out of g: the SAME split in the count and the fill pass
__device__ __forceinline__ void cta_range(unsigned long long n, unsigned b, unsigned g, unsigned long long &lo,
                                          unsigned long long &hi) {
  const unsigned long long per = (n + g - 1) / g;
  lo = per * b < n ? per * b : n;
  hi = lo + per < n ? lo + per : n;
}

template <class Seg> __global__ void __launch_bounds__(256) tile_count_kernel(TileArgs T, const Seg *seg, unsigned long long n) {
  extern __shared__ unsigned int s_hist[];
  for (int t = threadIdx.x; t < T.n_tiles; t += blockDim.x) s_hist[t] = 0u;
  __syncthreads();
  unsigned long long lo, hi;
  cta_range(n, blockIdx.x, gridDim.x, lo, hi);
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 ab = SegIO<Seg>::load_ab(seg, i);
    const RasterSetup S = raster_setup(T.A, ab.x, ab.y, ab.z, ab.w);
    for_each_tile(S, T.A.W, T.A.H, T.tiles_x, [&](int tile) { atomicAdd(&s_hist[tile], 1u); });
  }
  __syncthreads();
  unsigned int *out = T.hist + (size_t)blockIdx.x * T.n_tiles;
  for (int t = threadIdx.x; t < T.n_tiles; t += blockDim.x) out[t] = s_hist[t];
}

// thread = tile: exclusive scan down the CTA dimension; hist[c][t] becomes CTA c's offset inside tile t's list
__global__ void __launch_bounds__(256) tile_rowscan_kernel(TileArgs T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T.n_tiles) return;
  unsigned int run = 0;
  for (int c = 0; c < T.n_ctas; ++c) {
    unsigned int *p = T.hist + (size_t)c * T.n_tiles + t; // coalesced across the warp's tiles
    const unsigned int v = *p;
    *p = run;
    run += v;
  }
  T.tile_count[t] = run;
}

template <class Seg> __global__ void __launch_bounds__(256) tile_fill_kernel(TileArgs T, const Seg *seg, unsigned long long n) {
  extern __shared__ unsigned int s_cur[]; // this CTA's next slot in every tile's list, relative to the tile's offset
  const unsigned int *mine = T.hist + (size_t)blockIdx.x * T.n_tiles;
  for (int t = threadIdx.x; t < T.n_tiles; t += blockDim.x) s_cur[t] = mine[t];
  __syncthreads();
  unsigned long long lo, hi;
  cta_range(n, blockIdx.x, gridDim.x, lo, hi);
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 ab = SegIO<Seg>::load_ab(seg, i);
    const RasterSetup S = raster_setup(T.A, ab.x, ab.y, ab.z, ab.w);
    for_each_tile(S, T.A.W, T.A.H, T.tiles_x, [&](int tile) {
      const unsigned int pos = atomicAdd(&s_cur[tile], 1u);
      T.list[T.tile_offset[tile] + pos] = (unsigned int)i;
    });
  }
}

// one block: exclusive scans of the per-tile counts (list offsets) and of the per-tile work items
__global__ void __launch_bounds__(1024) tile_scan_kernel(TileArgs T) {
  __shared__ unsigned long long s_pairs[1024];
  __shared__ unsigned int s_items[1024];
  __shared__ unsigned long long carry_pairs;
  __shared__ unsigned int carry_items;
  if (threadIdx.x == 0) carry_pairs = 0, carry_items = 0;
  __syncthreads();
  for (int base = 0; base < T.n_tiles; base += 1024) {
    const int t = base + threadIdx.x;
    const unsigned int c = t < T.n_tiles ? T.tile_count[t] : 0u;
    const unsigned int it = (c + kChunk - 1) / kChunk;
    s_pairs[threadIdx.x] = c;
    s_items[threadIdx.x] = it;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) { // Hillis-Steele inclusive scan
      unsigned long long vp = 0;
      unsigned int vi = 0;
      if ((int)threadIdx.x >= off) vp = s_pairs[threadIdx.x - off], vi = s_items[threadIdx.x - off];
      __syncthreads();
      s_pairs[threadIdx.x] += vp, s_items[threadIdx.x] += vi;
      __syncthreads();
    }
    if (t < T.n_tiles) {
      T.tile_offset[t] = carry_pairs + s_pairs[threadIdx.x] - c;
      T.item_prefix[t] = carry_items + s_items[threadIdx.x] - it;
      T.tile_cursor[t] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_pairs += s_pairs[1023], carry_items += s_items[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    T.tile_offset[T.n_tiles] = carry_pairs;
    T.item_prefix[T.n_tiles] = carry_items;
    T.totals[0] = carry_pairs;
    T.totals[1] = carry_items;
    *T.item_counter = 0u;
  }
}

struct RasterScratch { // per warp
  RasterSetup s[32];
  float4 ca[32];
  float4 dc[32];
};

template <class Seg> __global__ void __launch_bounds__(kRasterWarps * 32) tile_raster_kernel(TileArgs T, const Seg *seg) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  const int warp_in_block = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  float4 *tile = reinterpret_cast<float4 *>(tile_smem_raw) + (size_t)warp_in_block * kTileFloat4;
  RasterScratch &P = reinterpret_cast<RasterScratch *>(reinterpret_cast<float4 *>(tile_smem_raw) +
                                                       (size_t)kRasterWarps * kTileFloat4)[warp_in_block];
  constexpr bool kLerp = SegIO<Seg>::kLerp;
  const unsigned int n_items = (unsigned int)T.totals[1];
  unsigned long long cnt = 0;
  while (true) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(T.item_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    // item -> (tile, chunk): last tile whose item prefix is <= item
    int lo = 0, hi = T.n_tiles;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (T.item_prefix[mid] <= item) lo = mid; else hi = mid;
    }
    const int t = lo;
    const unsigned int chunk = item - T.item_prefix[t];
    const unsigned int count = T.tile_count[t];
    const unsigned int first = chunk * kChunk, last = min(count, first + kChunk);
    const unsigned int *lst = T.list + T.tile_offset[t];
    const int tx = t % T.tiles_x, ty = t / T.tiles_x;
    for (int k = lane; k < kTileFloat4; k += 32) tile[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    // the gather of the NEXT 32 list entries (index, then segment) is in flight while the current 32 are blended
    float4 g_ab = make_float4(0.f, 0.f, 0.f, 0.f), g_ca = g_ab, g_dc = g_ab;
    if (first + lane < last) SegIO<Seg>::load(seg, lst[first + lane], g_ab, g_ca, g_dc);
    for (unsigned int base = first; base < last; base += 32) {
      if (base + lane < last) { // lane = one list entry: park its raster setup
        P.s[lane] = raster_setup(T.A, g_ab.x, g_ab.y, g_ab.z, g_ab.w);
        P.ca[lane] = g_ca;
        if (kLerp) P.dc[lane] = g_dc;
      }
      if (base + 32 + lane < last) SegIO<Seg>::load(seg, lst[base + 32 + lane], g_ab, g_ca, g_dc);
      __syncwarp();
      const int m = (int)min(32u, last - base);
      // Two parked segments are in flight at a time: their coordinate math is independent (ILP hides the
      // shared-memory latency), the two read-modify-writes then happen in list order (a lane may hit the same
      // pixel in both).  The next two are fetched from shared memory before the current two are blended.
      RasterSetup Sa = P.s[0], Sb = P.s[m > 1 ? 1 : 0];
      float4 ca_a = P.ca[0], ca_b = P.ca[m > 1 ? 1 : 0];
      float4 dc_a = kLerp ? P.dc[0] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 dc_b = kLerp ? P.dc[m > 1 ? 1 : 0] : dc_a;
      for (int k = 0; k < m; k += 2) {
        const bool has_b = k + 1 < m;
        const int kn0 = k + 2 < m ? k + 2 : k, kn1 = k + 3 < m ? k + 3 : k;
        const RasterSetup Na = P.s[kn0], Nb = P.s[kn1];
        const float4 nca_a = P.ca[kn0], nca_b = P.ca[kn1];
        const float4 ndc_a = kLerp ? P.dc[kn0] : dc_a, ndc_b = kLerp ? P.dc[kn1] : dc_a;
        // this tile's stretch of each segment: one major-axis step per lane (arithmetic of raster_walk)
        int off_a, off_b;
        float s_a, s_b;
        bool act_a, act_b;
        {
          const int i = ((Sa.xmajor ? tx : ty) << kTileShift) + (int)lane;
          const float mc = (float)i + 0.5f;
          s_a = (mc - Sa.m0) * Sa.inv;
          const float fj = floorf(__fmaf_rn(s_a, Sa.dn, Sa.n0));
          const int Nmin = Sa.xmajor ? T.A.H : T.A.W;
          const int j = (int)fmaxf(fminf(fj, 1e9f), -1e9f);
          act_a = i >= Sa.i0 && i < Sa.i1 && fj >= 0.f && fj < (float)Nmin && (j >> kTileShift) == (Sa.xmajor ? ty : tx);
          const int jl = j & (kTile - 1);
          off_a = Sa.xmajor ? jl * kTilePitch + (int)lane : (int)lane * kTilePitch + jl;
        }
        {
          const int i = ((Sb.xmajor ? tx : ty) << kTileShift) + (int)lane;
          const float mc = (float)i + 0.5f;
          s_b = (mc - Sb.m0) * Sb.inv;
          const float fj = floorf(__fmaf_rn(s_b, Sb.dn, Sb.n0));
          const int Nmin = Sb.xmajor ? T.A.H : T.A.W;
          const int j = (int)fmaxf(fminf(fj, 1e9f), -1e9f);
          act_b = has_b && i >= Sb.i0 && i < Sb.i1 && fj >= 0.f && fj < (float)Nmin &&
                  (j >> kTileShift) == (Sb.xmajor ? ty : tx);
          const int jl = j & (kTile - 1);
          off_b = Sb.xmajor ? jl * kTilePitch + (int)lane : (int)lane * kTilePitch + jl;
        }
        if (act_a) {
          float c0 = ca_a.x, c1 = ca_a.y, c2 = ca_a.z, c3 = ca_a.w;
          if (kLerp) {
            c0 = __fmaf_rn(s_a, dc_a.x, ca_a.x), c1 = __fmaf_rn(s_a, dc_a.y, ca_a.y);
            c2 = __fmaf_rn(s_a, dc_a.z, ca_a.z), c3 = __fmaf_rn(s_a, dc_a.w, ca_a.w);
          }
          float4 v = tile[off_a]; // private tile, distinct pixel per lane: plain read-modify-write
          v.x += c0, v.y += c1, v.z += c2, v.w += c3 * c3; // mod.rs:57-73
          tile[off_a] = v;
          ++cnt;
        }
        if (act_b) {
          float c0 = ca_b.x, c1 = ca_b.y, c2 = ca_b.z, c3 = ca_b.w;
          if (kLerp) {
            c0 = __fmaf_rn(s_b, dc_b.x, ca_b.x), c1 = __fmaf_rn(s_b, dc_b.y, ca_b.y);
            c2 = __fmaf_rn(s_b, dc_b.z, ca_b.z), c3 = __fmaf_rn(s_b, dc_b.w, ca_b.w);
          }
          float4 v = tile[off_b];
          v.x += c0, v.y += c1, v.z += c2, v.w += c3 * c3;
          tile[off_b] = v;
          ++cnt;
        }
        Sa = Na, Sb = Nb, ca_a = nca_a, ca_b = nca_b, dc_a = ndc_a, dc_b = ndc_b;
      }
      __syncwarp();
    }
    // flush: one vector reduction per touched pixel (lanes sweep a row: coalesced 512 B)
    const int gx = (tx << kTileShift) + (int)lane;
    for (int row = 0; row < kTile; ++row) {
      const int gy = (ty << kTileShift) + row;
      const float4 v = tile[row * kTilePitch + lane];
      if (gx < T.A.W && gy < T.A.H && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f))
        red_add_v4(T.A.img + ((size_t)gy * T.A.W + gx) * 4, v.x, v.y, v.z, v.w);
    }
    __syncwarp();
  }
  flush_count(T.A, cnt, lane);
}

// host vertex pairs (f64 positions, `as f32`) -> Seg2
__global__ void pairs_to_seg2_kernel(const LgVertexPair *vp, Seg2 *out, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const LgVertexPair &s = vp[i];
  Seg2 o;
  o.ab = make_float4((float)s.a[0], (float)s.a[1], (float)s.b[0], (float)s.b[1]);
  o.ca = make_float4(s.color_a[0], s.color_a[1], s.color_a[2], s.color_a[3]);
  o.cb = make_float4(s.color_b[0], s.color_b[1], s.color_b[2], s.color_b[3]);
  out[i] = o;
}

// StringMod::draw into Seg2 (the tiled path needs the chords in memory; the direct path does not)
__global__ void string_mod_seg2_kernel(StringModArgs S, Seg2 *out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.count) return;
  const unsigned long long iix = S.first + i, ix = sm_target(S.sm, iix);
  Seg2 o;
  float ax, ay, bx, by, ca[4], cb[4];
  sm_point(S.sm, iix, ax, ay);
  sm_point(S.sm, ix, bx, by);
  sm_color(S, iix, ca);
  sm_color(S, ix, cb);
  o.ab = make_float4(ax, ay, bx, by);
  o.ca = make_float4(ca[0], ca[1], ca[2], ca[3]);
  o.cb = make_float4(cb[0], cb[1], cb[2], cb[3]);
  out[i] = o;
}

} // namespace lg
