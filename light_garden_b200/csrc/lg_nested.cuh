// lg_nested.cuh — StringMod with `nested: Some(inner)` (src/light_garden/string_mod.rs:87-101,152-158): the outer
// pattern's chords are intersected pairwise, the crossing points — in the reference's order — become the point set
// the inner pattern draws its chords between.  ORACLE.md §7.2 defines the segment-segment intersection
// (collision2d's LineSegment::intersect is not available).
//
//   lines     : string_mod_pairs_kernel  -> LgVertexPair per outer chord (f64 end points)
//   crossings : candidate q = (diff - 1) * L + ixa  (diff = 1..L-1, ixa = 0..L-1), partner (ixa + diff) % L;
//               count per block -> scan -> ordered write (two passes over the L (L - 1) candidates)
//   chords    : nested_pairs_kernel      -> LgVertexPair per inner chord, accumulated like host lines
#pragma once
#include "lg_accum.cuh"

namespace lg {

constexpr int kNestBlock = 256;  // threads per block of the crossing passes
constexpr int kNestPer = 16;     // consecutive candidates per thread

__global__ void string_mod_pairs_kernel(StringModArgs S, LgVertexPair *out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.count) return;
  const unsigned long long iix = S.first + i, ix = sm_target(S.sm, iix);
  LgVertexPair o;
  sm_point64(S.sm, iix, o.a[0], o.a[1]);
  sm_point64(S.sm, ix, o.b[0], o.b[1]);
  sm_color(S, iix, o.color_a);
  sm_color(S, ix, o.color_b);
  out[i] = o;
}

// ORACLE.md §7.2: p = a1 + t e1 = a2 + u e2 with t, u in [0, 1] (end points included), f64, explicit fma
__device__ __forceinline__ bool segments_cross(const LgVertexPair &p, const LgVertexPair &q, double &x, double &y) {
  const double e1x = __dsub_rn(p.b[0], p.a[0]), e1y = __dsub_rn(p.b[1], p.a[1]);
  const double e2x = __dsub_rn(q.b[0], q.a[0]), e2y = __dsub_rn(q.b[1], q.a[1]);
  const double denom = __fma_rn(e1x, e2y, -__dmul_rn(e1y, e2x));
  if (!(fabs(denom) > 1e-12)) return false;
  const double wx = __dsub_rn(q.a[0], p.a[0]), wy = __dsub_rn(q.a[1], p.a[1]);
  const double t = __ddiv_rn(__fma_rn(wx, e2y, -__dmul_rn(wy, e2x)), denom);
  const double u = __ddiv_rn(__fma_rn(wx, e1y, -__dmul_rn(wy, e1x)), denom);
  if (!(t >= 0.0) || !(t <= 1.0) || !(u >= 0.0) || !(u <= 1.0)) return false;
  x = __fma_rn(t, e1x, p.a[0]);
  y = __fma_rn(t, e1y, p.a[1]);
  return true;
}

// candidates [lo, hi) of this thread; n_cand = L (L - 1)
__device__ __forceinline__ void nest_range(unsigned long long n_cand, unsigned long long &lo, unsigned long long &hi) {
  lo = ((unsigned long long)blockIdx.x * kNestBlock + threadIdx.x) * kNestPer;
  hi = lo + kNestPer < n_cand ? lo + kNestPer : n_cand;
  if (lo > n_cand) lo = n_cand;
}
__device__ __forceinline__ bool nest_candidate(const LgVertexPair *lines, unsigned long long L, unsigned long long q,
                                               double &x, double &y) {
  const unsigned long long diff = q / L + 1ull, ixa = q % L;
  unsigned long long ixb = ixa + diff;
  if (ixb >= L) ixb -= L;
  return segments_cross(lines[ixa], lines[ixb], x, y);
}

__global__ void __launch_bounds__(kNestBlock) nested_count_kernel(const LgVertexPair *lines, unsigned long long L,
                                                                  unsigned long long n_cand, unsigned long long *block_count) {
  __shared__ unsigned int warp_sum[kNestBlock / 32];
  unsigned long long lo, hi;
  nest_range(n_cand, lo, hi);
  unsigned int n = 0;
  double x, y;
  for (unsigned long long q = lo; q < hi; ++q) n += nest_candidate(lines, L, q, x, y) ? 1u : 0u;
  for (int off = 16; off > 0; off >>= 1) n += __shfl_down_sync(0xffffffffu, n, off);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < kNestBlock / 32; ++w) t += warp_sum[w];
    block_count[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts in place (one block; the counts are few: L (L - 1) / 4096)
__global__ void __launch_bounds__(1024) nested_scan_kernel(unsigned long long *block_count, unsigned long long n_blocks,
                                                           unsigned long long *total) {
  __shared__ unsigned long long part[1024];
  const unsigned long long per = (n_blocks + 1023ull) / 1024ull;
  const unsigned long long lo = per * threadIdx.x < n_blocks ? per * threadIdx.x : n_blocks;
  const unsigned long long hi = lo + per < n_blocks ? lo + per : n_blocks;
  unsigned long long s = 0;
  for (unsigned long long i = lo; i < hi; ++i) s += block_count[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) {
      const unsigned long long v = part[i];
      part[i] = run;
      run += v;
    }
    *total = run;
  }
  __syncthreads();
  unsigned long long run = part[threadIdx.x];
  for (unsigned long long i = lo; i < hi; ++i) {
    const unsigned long long v = block_count[i];
    block_count[i] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(kNestBlock) nested_write_kernel(const LgVertexPair *lines, unsigned long long L,
                                                                  unsigned long long n_cand,
                                                                  const unsigned long long *block_offset, double2 *points) {
  __shared__ unsigned int warp_sum[kNestBlock / 32];
  unsigned long long lo, hi;
  nest_range(n_cand, lo, hi);
  double px[kNestPer], py[kNestPer];
  unsigned int hit = 0, n = 0;
#pragma unroll
  for (int k = 0; k < kNestPer; ++k) {
    px[k] = py[k] = 0.0;
    if (lo + k < hi && nest_candidate(lines, L, lo + k, px[k], py[k])) hit |= 1u << k, ++n;
  }
  // exclusive scan of n over the block: within the warp by shuffles, across warps through shared memory
  const unsigned lane = threadIdx.x & 31u;
  unsigned int incl = n;
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= (unsigned)off) incl += v;
  }
  if (lane == 31u) warp_sum[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned int before = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_sum[w];
  unsigned long long at = block_offset[blockIdx.x] + before + (incl - n);
#pragma unroll
  for (int k = 0; k < kNestPer; ++k)
    if (hit & (1u << k)) points[at++] = make_double2(px[k], py[k]);
}

// inner.draw_init_points(points) (string_mod.rs:103-122): chord iix joins points[iix % P] and points[f(iix) % P],
// coloured by the inner pattern's rules at iix and f(iix)
__global__ void nested_pairs_kernel(StringModArgs S, const double2 *points, unsigned long long P, LgVertexPair *out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.count) return;
  const unsigned long long iix = S.first + i, ix = sm_target(S.sm, iix);
  const double2 a = points[iix % P], b = points[ix % P];
  LgVertexPair o;
  o.a[0] = a.x, o.a[1] = a.y, o.b[0] = b.x, o.b[1] = b.y;
  sm_color(S, iix, o.color_a);
  sm_color(S, ix, o.color_b);
  out[i] = o;
}

} // namespace lg
