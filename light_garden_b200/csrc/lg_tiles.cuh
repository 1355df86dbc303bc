// lg_tiles.cuh — K4', the tile-binned form of the line accumulation.
//
// Same semantics as lg_accum.cuh (ORACLE.md §8; reference: src/sub_render_pass.rs:188-212, src/shader.wgsl,
// blend src/light_garden/mod.rs:57-73): what changes is where the additive blend is resolved.  The direct kernels
// send one 16-byte red.global.add per fragment to L2; at 500..70 000 fragments per pixel (BASELINE configs C1..C5)
// that is the bottleneck.  Here fragments are summed in SHARED MEMORY first:
//
//   1. tile_count : lane = segment; walk the 32x32-pixel tiles the segment's fragments fall into and count them in
//                   a per-CTA SHARED-MEMORY histogram (hot tiles next to a light take millions of hits: global
//                   atomics on one address would serialise), written out as hist[cta][list].  Every tile has TWO
//                   lists, one for x-major and one for y-major segments (list = 2 tile + axis)
//   2. tile_rowscan + tile_scan: exclusive scans -> every CTA's write position in every tile's list; work items =
//                   (tile, chunk of <= kChunk list entries)
//   3. tile_fill  : same walk by the same CTA over the same segments, positions from a shared-memory cursor
//   4. tile_raster: persistent warps; a warp owns a PRIVATE 32x32 RGBA fp32 tile in shared memory, stored
//                   major-axis-fastest for the list it is working on (row-major for an x-major list, transposed for a
//                   y-major one): lane = major-axis step, so the 8 lanes of a quarter warp always hit 8 different
//                   16-byte bank groups whatever the slope — conflict-free LDS.128 / FADD / STS.128 (a 33-pixel pitch
//                   serving both axes at once cost 2.06 wavefronts per ideal one, ncu r01d).  Lanes are distinct
//                   major-axis steps of one segment, so there are no collisions and no shared-memory atomics (those
//                   are CAS spin loops for float).  Finally the touched pixels are flushed with one
//                   red.global.add.v4.f32 each.
//
// Fragment coordinates use exactly the arithmetic of raster_walk(); a fragment lands in the tile that contains its
// pixel, every (segment, tile) pair that can own a fragment is listed (the walk brackets the minor coordinate of a
// clip interval by its two end fragments: j(i) is monotone), so coverage is identical to the direct kernels.
#pragma once
#include "lg_accum.cuh"

namespace lg {

constexpr int kTile = 32;          // tile edge in pixels
constexpr int kTileShift = 5;
constexpr int kTilePitch = 32;     // float4 per major-axis line of the tile in shared memory
constexpr int kChunk = 2048;       // list entries per work item (larger items were measured: 8192 saves 0.5 % of the
                                   // C5 step -- fewer tile clears and flushes -- and costs accuracy in the hot pixels,
                                   // whose partial sums then grow four times as large before they reach the image)
constexpr int kRasterWarps = 4;    // warps per CTA of tile_raster_kernel
constexpr int kTileFloat4 = kTile * kTilePitch;

// device segment with two end colours (string-mod chords, host vertex pairs)
struct Seg2 {
  float4 ab; // a.x a.y b.x b.y
  float4 ca, cb;
};

struct TileArgs {
  AccumArgs A;
  int tiles_x, tiles_y;
  int n_tiles; // number of LISTS = 2 x tiles (x-major and y-major segments of a tile are binned apart)
  unsigned int *tile_count;   // [n_tiles]
  unsigned int *tile_cursor;  // [n_tiles]
  unsigned long long *tile_offset; // [n_tiles + 1]
  unsigned int *item_prefix;  // [n_tiles + 1]
  unsigned long long *totals; // [0] = pairs, [1] = items (0 when the list is too small), [2] = 1: list too small,
                              // [3] = items
  unsigned int *list;         // segment index per pair
  unsigned long long list_cap; // entries `list` can hold: sized from the previous call, checked by tile_scan_kernel
  unsigned int *item_counter;
  unsigned int *hist;         // [n_ctas][n_tiles] per-CTA counts, then per-CTA exclusive offsets within a tile
  int n_ctas;
  // segment count and "do nothing" flag that live on the device (the trace kernel's counters): the passes of a wave
  // are queued behind the trace kernel without the host learning the count first.  NULL: use the kernel argument.
  const unsigned long long *n_dev;
  const unsigned int *skip_dev;
};

// number of segments the count / fill passes walk
__device__ __forceinline__ unsigned long long pass_segments(const TileArgs &T, unsigned long long n) {
  if (T.n_dev) n = *T.n_dev < n ? *T.n_dev : n; // n = capacity of the segment buffer in that case
  if (T.skip_dev && *T.skip_dev) n = 0ull;      // the wave overflowed its buffer: nothing of it is drawn
  return n;
}

template <class Seg> struct SegIO;
// One 8-byte read-only load.  The raster's gather prefetch uses four of them per segment instead of two 16-byte
// ones: a vector load needs an aligned register quad, and the compiler copied one component out of that quad right
// behind the load (ncu r01f: 13 % of the raster kernel's samples on that one move, the prefetch waited out in full).
// Register pairs are placed without such a move; 4-byte loads also avoid it but double the sector requests of the
// gather again (C2: 555 -> 573 ms).
__device__ __forceinline__ void ldg_nc_v2(const float *p, float &a, float &b) {
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "l"(p));
}
template <> struct SegIO<LgSegment> {
  static constexpr bool kLerp = false;
  static __device__ __forceinline__ void load(const LgSegment *s, unsigned long long i, float4 &ab, float4 &ca, float4 &dc) {
    const float *p = reinterpret_cast<const float *>(s + i);
    ldg_nc_v2(p, ab.x, ab.y), ldg_nc_v2(p + 2, ab.z, ab.w);
    ldg_nc_v2(p + 4, ca.x, ca.y), ldg_nc_v2(p + 6, ca.z, ca.w);
    dc = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  static __device__ __forceinline__ float4 load_ab(const LgSegment *s, unsigned long long i) {
    return __ldg(reinterpret_cast<const float4 *>(s + i));
  }
};
template <> struct SegIO<Seg2> {
  static constexpr bool kLerp = true;
  // `cb` comes back raw: the colour difference is taken when the record is parked, a batch later -- taken here it
  // made the prefetch wait for its own loads
  static __device__ __forceinline__ void load(const Seg2 *s, unsigned long long i, float4 &ab, float4 &ca, float4 &cb) {
    const float *p = reinterpret_cast<const float *>(s + i);
    ldg_nc_v2(p, ab.x, ab.y), ldg_nc_v2(p + 2, ab.z, ab.w);
    ldg_nc_v2(p + 4, ca.x, ca.y), ldg_nc_v2(p + 6, ca.z, ca.w);
    ldg_nc_v2(p + 8, cb.x, cb.y), ldg_nc_v2(p + 10, cb.z, cb.w);
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

template <class Seg> __global__ void __launch_bounds__(1024) tile_count_kernel(TileArgs T, const Seg *seg, unsigned long long n) {
  extern __shared__ unsigned int s_hist[];
  for (int t = threadIdx.x; t < T.n_tiles; t += blockDim.x) s_hist[t] = 0u;
  __syncthreads();
  n = pass_segments(T, n);
  unsigned long long lo, hi;
  cta_range(n, blockIdx.x, gridDim.x, lo, hi);
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 ab = SegIO<Seg>::load_ab(seg, i);
    const RasterSetup S = raster_setup(T.A, ab.x, ab.y, ab.z, ab.w);
    const int axis = S.xmajor ? 0 : 1;
    for_each_tile(S, T.A.W, T.A.H, T.tiles_x, [&](int tile) { atomicAdd(&s_hist[2 * tile + axis], 1u); });
  }
  __syncthreads();
  unsigned int *out = T.hist + (size_t)blockIdx.x * T.n_tiles;
  for (int t = threadIdx.x; t < T.n_tiles; t += blockDim.x) out[t] = s_hist[t];
}

// Exclusive scan down the CTA dimension: hist[c][t] becomes CTA c's offset inside list t.  A block owns 32 lists
// (lane = list: every row access is one coalesced 128-byte line) and its 8 warps split the CTA rows between them:
// slice sums first, then the scan of each slice from its carry-in (the table is read twice, from L2).  One thread
// per list walking all rows took 2.1 ms of the 35 ms accumulate of C5 (r01f launch list): 64 blocks for 148 SMs.
constexpr int kRowscanWarps = 8;
__global__ void __launch_bounds__(32 * kRowscanWarps) tile_rowscan_kernel(TileArgs T) {
  __shared__ unsigned int part[kRowscanWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  const int c0 = (int)((long long)T.n_ctas * w / kRowscanWarps), c1 = (int)((long long)T.n_ctas * (w + 1) / kRowscanWarps);
  unsigned int sum = 0u;
  if (t < T.n_tiles) {
    const unsigned int *p = T.hist + (size_t)c0 * T.n_tiles + t;
    int c = c0;
    for (; c + 4 <= c1; c += 4, p += 4 * (size_t)T.n_tiles)
      sum += p[0] + p[T.n_tiles] + p[2 * (size_t)T.n_tiles] + p[3 * (size_t)T.n_tiles];
    for (; c < c1; ++c, p += T.n_tiles) sum += *p;
  }
  part[w][lane] = sum;
  __syncthreads();
  if (t >= T.n_tiles) return;
  unsigned int run = 0u;
  for (int k = 0; k < w; ++k) run += part[k][lane];
  unsigned int *p = T.hist + (size_t)c0 * T.n_tiles + t;
  for (int c = c0; c < c1; ++c, p += T.n_tiles) {
    const unsigned int v = *p;
    *p = run;
    run += v;
  }
  if (w == kRowscanWarps - 1) T.tile_count[t] = run;
}

// The cursors in shared memory give ABSOLUTE positions in the pair list, so an entry costs one shared-memory atomic and
// one store (with cursors relative to the list's start the store's address also waits for a global load of
// tile_offset[list] per entry).  kMode, chosen by the host:
//   0  the list's capacity is below 2^32: a cursor is the 32-bit absolute position;
//   1  larger lists (C2 at full size: 5.7e9 pairs): 32-bit cursors relative to the list's start (64-bit atomics on shared
//      memory are CAS loops) and the 64-bit starts staged beside them, 12 bytes per list;
//   2  larger lists on frames whose lists do not fit 12 bytes each in shared memory: the starts are read from global memory.
// What bounds the pass is the scattered 4-byte store (profiles/r02_tile_fill_c2.txt): the same walk
// with the atomic's return value used but nothing stored takes 8.6 ms where this takes 34.
template <class Seg, int kMode>
__global__ void __launch_bounds__(1024) tile_fill_kernel(TileArgs T, const Seg *seg, unsigned long long n) {
  extern __shared__ __align__(8) unsigned char s_fill_raw[];
  unsigned long long *s_off = reinterpret_cast<unsigned long long *>(s_fill_raw);                          // mode 1 only
  unsigned int *s_cur = reinterpret_cast<unsigned int *>(s_fill_raw + (kMode == 1 ? (size_t)T.n_tiles * 8 : 0)); // next slot per list
  const unsigned int *mine = T.hist + (size_t)blockIdx.x * T.n_tiles;
  if (T.totals[2]) return; // the list is too small (tile_scan_kernel): the host grows it and runs this pass again
  for (int t = threadIdx.x; t < T.n_tiles; t += blockDim.x) {
    if (kMode == 0) s_cur[t] = (unsigned int)(T.tile_offset[t] + mine[t]);
    else s_cur[t] = mine[t];
    if (kMode == 1) s_off[t] = T.tile_offset[t];
  }
  __syncthreads();
  n = pass_segments(T, n);
  unsigned long long lo, hi;
  cta_range(n, blockIdx.x, gridDim.x, lo, hi);
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 ab = SegIO<Seg>::load_ab(seg, i);
    const RasterSetup S = raster_setup(T.A, ab.x, ab.y, ab.z, ab.w);
    const int axis = S.xmajor ? 0 : 1;
    for_each_tile(S, T.A.W, T.A.H, T.tiles_x, [&](int tile) {
      const int l = 2 * tile + axis;
      const unsigned int p = atomicAdd(&s_cur[l], 1u);
      if (kMode == 0) T.list[p] = (unsigned int)i;
      else T.list[(kMode == 1 ? s_off[l] : T.tile_offset[l]) + p] = (unsigned int)i;
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
  constexpr unsigned int chunk_sz = kChunk;
  for (int base = 0; base < T.n_tiles; base += 1024) {
    const int t = base + threadIdx.x;
    const unsigned int c = t < T.n_tiles ? T.tile_count[t] : 0u;
    const unsigned int it = (c + chunk_sz - 1) / chunk_sz;
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
    const bool fits = carry_pairs <= T.list_cap;
    T.totals[0] = carry_pairs;
    T.totals[1] = fits ? carry_items : 0u; // nothing for the raster to do until the list has been grown and filled
    T.totals[2] = fits ? 0ull : 1ull;
    T.totals[3] = carry_items;
    *T.item_counter = 0u;
  }
}

// Per-warp parking area of the raster kernel: 32 list entries reduced to what the blend loop reads.
struct RasterScratch {
  float4 geo[32];        // m0, 1/(m1 - m0), minor delta, minor start (RasterSetup)
  float4 col[32];        // colour at a (single-colour segments: alpha already squared)
  float4 dc[32];         // colour delta (two-colour segments only)
  unsigned int mask[32]; // lanes of this tile inside the segment's major-axis range [i0, i1)
};

// the records of two parked entries, as the blend loop holds them in registers
struct PairRecs {
  float4 ga, gb, ca, cb, da, db;
  uint2 m;
};
template <bool kLerp> __device__ __forceinline__ void load_recs(const RasterScratch &P, int k, PairRecs &q) {
  q.ga = P.geo[k], q.gb = P.geo[k + 1], q.ca = P.col[k], q.cb = P.col[k + 1];
  if (kLerp) q.da = P.dc[k], q.db = P.dc[k + 1];
  q.m = *reinterpret_cast<const uint2 *>(&P.mask[k]);
}
// Blends two parked entries into the warp's private tile.  lane = major-axis step; the tile is stored
// major-axis-fastest, so a fragment at minor coordinate j sits at lane_base + j * row_bytes (lane_base is this lane's
// byte address of canvas minor coordinate 0).  The arithmetic per fragment is raster_walk's (ORACLE.md 8.2-8.4): the
// minor coordinate is floor(fma(s, dn, n0)), taken here with one float -> int conversion rounding down (the same
// integer for every value a tile can own), and "inside this tile" is ONE unsigned compare of j - bmin against the
// tile's extent; which lanes lie inside the segment's major-axis range is a bit mask made when the entry is parked.
// Both pixels are read before either is written; a lane that hits the same pixel in both carries the first sum into
// the second.  Straight-line predicated code: no branch.
template <bool kLerp>
__device__ __forceinline__ void blend_two(unsigned char *lane_base, int row_bytes, const PairRecs &q, float mc, unsigned lane_bit,
                                          int bmin, unsigned extent, unsigned &cnt) {
  const float sa = (mc - q.ga.x) * q.ga.y, sb = (mc - q.gb.x) * q.gb.y;
  const int ja = __float2int_rd(__fmaf_rn(sa, q.ga.z, q.ga.w)), jb = __float2int_rd(__fmaf_rn(sb, q.gb.z, q.gb.w));
  const bool act_a = (q.m.x & lane_bit) != 0u && (unsigned)(ja - bmin) < extent;
  const bool act_b = (q.m.y & lane_bit) != 0u && (unsigned)(jb - bmin) < extent;
  float4 *pa = reinterpret_cast<float4 *>(lane_base + ja * row_bytes);
  float4 *pb = reinterpret_cast<float4 *>(lane_base + jb * row_bytes);
  float a0 = q.ca.x, a1 = q.ca.y, a2 = q.ca.z, a3 = q.ca.w, b0 = q.cb.x, b1 = q.cb.y, b2 = q.cb.z, b3 = q.cb.w;
  if (kLerp) {
    a0 = __fmaf_rn(sa, q.da.x, q.ca.x), a1 = __fmaf_rn(sa, q.da.y, q.ca.y), a2 = __fmaf_rn(sa, q.da.z, q.ca.z);
    a3 = __fmaf_rn(sa, q.da.w, q.ca.w), a3 *= a3;
    b0 = __fmaf_rn(sb, q.db.x, q.cb.x), b1 = __fmaf_rn(sb, q.db.y, q.cb.y), b2 = __fmaf_rn(sb, q.db.z, q.cb.z);
    b3 = __fmaf_rn(sb, q.db.w, q.cb.w), b3 *= b3;
  }
  float4 va, vb; // an inactive lane adds into registers nobody reads: no need to clear them
  if (act_a) va = *pa;
  if (act_b) vb = *pb;
  va.x += a0, va.y += a1, va.z += a2, va.w += a3; // mod.rs:57-73
  const bool same = act_a && ja == jb;            // (only read when act_b)
  vb.x = same ? va.x : vb.x, vb.y = same ? va.y : vb.y, vb.z = same ? va.z : vb.z, vb.w = same ? va.w : vb.w;
  vb.x += b0, vb.y += b1, vb.z += b2, vb.w += b3;
  if (act_a) *pa = va, ++cnt;
  if (act_b) *pb = vb, ++cnt;
}
template <bool kLerp>
__device__ __forceinline__ unsigned blend_run(unsigned char *lane_base, int row_bytes, const RasterScratch &P, int m, float mc,
                                              unsigned lane_bit, int bmin, unsigned extent) {
  // entries [0, m) two at a time; when m is odd the caller has parked an empty entry (mask 0) at index m.
  // The records of the NEXT two entries are fetched before the current two touch the tile (the compiler cannot
  // move those loads across the tile's stores by itself: same shared-memory array); the two register sets swap
  // roles every half iteration, so nothing is copied.
  unsigned cnt = 0u;
  PairRecs A, B;
  A.da = A.db = B.da = B.db = make_float4(0.f, 0.f, 0.f, 0.f);
  load_recs<kLerp>(P, 0, A);
  for (int k = 0; k < m; k += 4) {
    load_recs<kLerp>(P, k + 2 < m ? k + 2 : k, B);
    blend_two<kLerp>(lane_base, row_bytes, A, mc, lane_bit, bmin, extent, cnt);
    if (k + 2 >= m) break;
    load_recs<kLerp>(P, k + 4 < m ? k + 4 : k, A);
    blend_two<kLerp>(lane_base, row_bytes, B, mc, lane_bit, bmin, extent, cnt);
  }
  return cnt;
}

template <class Seg> __global__ void __launch_bounds__(kRasterWarps * 32) tile_raster_kernel(TileArgs T, const Seg *seg) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  const int warp_in_block = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  float4 *tile = reinterpret_cast<float4 *>(tile_smem_raw) + (size_t)warp_in_block * kTileFloat4;
  RasterScratch &P = reinterpret_cast<RasterScratch *>(reinterpret_cast<float4 *>(tile_smem_raw) +
                                                       (size_t)kRasterWarps * kTileFloat4)[warp_in_block];
  constexpr bool kLerp = SegIO<Seg>::kLerp;
  const unsigned int n_items = (unsigned int)T.totals[1];
  constexpr unsigned int chunk_sz = kChunk;
  unsigned long long cnt = 0;
  while (true) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(T.item_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    // item -> (list, chunk): last list whose item prefix is <= item
    int lo = 0, hi = T.n_tiles;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (T.item_prefix[mid] <= item) lo = mid; else hi = mid;
    }
    const int list = lo, t = list >> 1;
    const bool xmajor = (list & 1) == 0;
    const unsigned int chunk = item - T.item_prefix[list];
    const unsigned int count = T.tile_count[list];
    const unsigned int first = chunk * chunk_sz, last = min(count, first + chunk_sz);
    const unsigned int *lst = T.list + T.tile_offset[list];
    const int tx = t % T.tiles_x, ty = t / T.tiles_x;
    const int bx = tx << kTileShift, by = ty << kTileShift;
    // loop constants of the blend loop: lane = column for an x-major list (tile stored row-major), lane = row for a
    // y-major one (tile stored transposed)
    const int bmaj = xmajor ? bx : by, bmin = xmajor ? by : bx;
    const float mc = (float)(bmaj + (int)lane) + 0.5f;
    const unsigned extent = (unsigned)(min(xmajor ? T.A.H : T.A.W, bmin + kTile) - bmin); // minor pixels of the tile on the canvas
    // byte address of this lane's pixel at minor offset 0 of the canvas (the tile starts at minor offset bmin)
    unsigned char *lane_base = reinterpret_cast<unsigned char *>(tile) + ((int)lane - bmin * kTilePitch) * (int)sizeof(float4);
    for (int k = lane; k < kTileFloat4; k += 32) tile[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    // gather pipeline, two batches deep: while batch b is blended, the SEGMENTS of batch b + 1 (their list indices
    // arrived during the previous batch) and the INDICES of batch b + 2 are in flight
    float4 g_ab = make_float4(0.f, 0.f, 0.f, 0.f), g_ca = g_ab, g_dc = g_ab;
    if (first + lane < last) SegIO<Seg>::load(seg, lst[first + lane], g_ab, g_ca, g_dc);
    unsigned int idx_next = first + 32 + lane < last ? lst[first + 32 + lane] : 0u;
    for (unsigned int base = first; base < last; base += 32) {
      // lane = one list entry: its raster setup, reduced to this tile, parked for the blend loop
      const int m = (int)min(32u, last - base);
      if ((int)lane < m) {
        const RasterSetup S = raster_setup(T.A, g_ab.x, g_ab.y, g_ab.z, g_ab.w);
        const int l0 = max(S.i0 - bmaj, 0), l1 = min(S.i1 - bmaj, kTile);
        const int nl = l1 - l0;
        const unsigned msk = nl <= 0 ? 0u : ((nl >= 32 ? 0xffffffffu : ((1u << nl) - 1u)) << l0);
        P.geo[lane] = make_float4(S.m0, S.inv, S.dn, S.n0);
        P.col[lane] = make_float4(g_ca.x, g_ca.y, g_ca.z, kLerp ? g_ca.w : g_ca.w * g_ca.w);
        if (kLerp) P.dc[lane] = make_float4(g_dc.x - g_ca.x, g_dc.y - g_ca.y, g_dc.z - g_ca.z, g_dc.w - g_ca.w);
        P.mask[lane] = msk;
      } else if ((int)lane == m) { // the empty partner of an odd last entry
        P.geo[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        P.col[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kLerp) P.dc[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        P.mask[lane] = 0u;
      }
      if (base + 32 + lane < last) SegIO<Seg>::load(seg, idx_next, g_ab, g_ca, g_dc);
      idx_next = base + 64 + lane < last ? lst[base + 64 + lane] : 0u;
      __syncwarp();
      cnt += blend_run<kLerp>(lane_base, kTilePitch * (int)sizeof(float4), P, m, mc, 1u << lane, bmin, extent);
      __syncwarp();
    }
    // flush: one vector reduction per touched pixel, lanes sweep an image row (coalesced 512 B).  A transposed
    // tile is read down its columns here: bank conflicts, but 1024 pixels against thousands of entries per item
    const int gx = bx + (int)lane;
    for (int row = 0; row < kTile; ++row) {
      const int gy = by + row;
      const float4 v = xmajor ? tile[row * kTilePitch + lane] : tile[lane * kTilePitch + row];
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
