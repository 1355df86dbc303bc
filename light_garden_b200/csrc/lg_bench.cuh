// lg_bench.cuh — roofline denominators BASELINE.md §2 leaves to the builder:
// the FP32 CUDA-core FMA peak and the L2 vector-reduction throughput, measured on
// the same device, clocks and power state as the kernels they are compared with.
#pragma once
#include <cuda_runtime.h>

namespace lg {

// 8 independent FFMA chains per thread, 2 flops per FFMA.
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = __fmaf_rn(x0, a, b), x1 = __fmaf_rn(x1, a, b), x2 = __fmaf_rn(x2, a, b), x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b), x5 = __fmaf_rn(x5, a, b), x6 = __fmaf_rn(x6, a, b), x7 = __fmaf_rn(x7, a, b);
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s; // never true: keeps the chains alive
}
constexpr int kFmaPerIter = 64; // FFMAs per thread per outer iteration

// same with double: the f64 mode's denominator
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3.;
  double x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = __fma_rn(x0, a, b), x1 = __fma_rn(x1, a, b), x2 = __fma_rn(x2, a, b), x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b), x5 = __fma_rn(x5, a, b), x6 = __fma_rn(x6, a, b), x7 = __fma_rn(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}

// red.global.add.v4.f32 throughput.  pattern 0: every warp sweeps consecutive
// pixels of a `span_px`-pixel image (coalesced 512 B per warp instruction);
// pattern 1: every lane hits a pseudo-random pixel (16 B scattered).
__global__ void __launch_bounds__(256) red_peak_kernel(float *img, unsigned long long span_px, int iters, int pattern) {
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long x = tid * 0x9E3779B97F4A7C15ull + 12345ull;
  for (int i = 0; i < iters; ++i) {
    unsigned long long px;
    if (pattern == 0) {
      px = (tid + (unsigned long long)i * nth) % span_px;
    } else {
      x ^= x << 13, x ^= x >> 7, x ^= x << 17;
      px = x % span_px;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(img + px * 4), "f"(1.f), "f"(1.f), "f"(1.f),
                 "f"(1.f)
                 : "memory");
  }
}

// Ceiling of the tile-binned resolve (lg_tiles.cuh): every warp owns a 32x32 RGBA fp32 tile in shared memory (the
// raster kernel's launch shape: 4 warps and 72 KB per CTA) and does nothing but the blend's memory work -- four
// independent 16-byte read-modify-writes per iteration, lane = column (conflict-free), the row taken from a value the
// compiler cannot see through.  Fragments/s of this loop is what LDS.128 + 4 FADD + STS.128 allow when every lane
// has a fragment and no instruction is spent on finding it.
__global__ void __launch_bounds__(128) tile_rmw_peak_kernel(float *sink, int iters) {
  extern __shared__ __align__(16) unsigned char rmw_smem[];
  float4 *tile = reinterpret_cast<float4 *>(rmw_smem) + (size_t)(threadIdx.x >> 5) * 1024;
  const unsigned lane = threadIdx.x & 31u;
  for (int k = lane; k < 1024; k += 32) tile[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  unsigned r = (blockIdx.x * 7u + (threadIdx.x >> 5)) & 31u;
  const float4 c = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int i = 0; i < iters; ++i) {
    float4 *p0 = tile + ((r + 0u) & 31u) * 32u + lane, *p1 = tile + ((r + 7u) & 31u) * 32u + lane;
    float4 *p2 = tile + ((r + 13u) & 31u) * 32u + lane, *p3 = tile + ((r + 22u) & 31u) * 32u + lane;
    float4 a = *p0, b = *p1, d = *p2, e = *p3;
    a.x += c.x, a.y += c.y, a.z += c.z, a.w += c.w;
    b.x += c.x, b.y += c.y, b.z += c.z, b.w += c.w;
    d.x += c.x, d.y += c.y, d.z += c.z, d.w += c.w;
    e.x += c.x, e.y += c.y, e.z += c.z, e.w += c.w;
    *p0 = a, *p1 = b, *p2 = d, *p3 = e;
    r = (r * 5u + 1u) & 31u;
  }
  __syncwarp();
  const float4 v = tile[lane];
  if (v.x + v.y + v.z + v.w == -1.f) sink[0] = v.x; // never true: keeps the loop alive
}

} // namespace lg
