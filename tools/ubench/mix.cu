// Instruction-mix microbenchmark for the trace kernel's broad phase (sm_100a): cycles per "pair" (two ray x object
// tests of one slot) for candidate formulations, all operands in registers, 8 independent pairs per iteration.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned shf(unsigned d, unsigned m) {
  unsigned r;
  asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(d), "r"(m));
  return r;
}
__device__ __forceinline__ unsigned fmax_(unsigned a, unsigned b) {
  unsigned r;
  asm volatile("max.f32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ unsigned imax_(unsigned a, unsigned b) {
  unsigned r;
  asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ unsigned ior_(unsigned a, unsigned b) {
  unsigned r;
  asm volatile("or.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <int KIND> __global__ void k(unsigned *out, unsigned seed, long long *cycles) {
  u64 x[8], y[8], r2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = ((u64)(seed + i) << 32) | (threadIdx.x + i);
    y[i] = ((u64)(seed * 3 + i) << 32) | (threadIdx.x * 5 + i);
    r2[i] = ((u64)(seed * 7 + i) << 32) | (threadIdx.x * 9 + i);
  }
  u64 sdy = seed * 11ull, sdx = seed * 13ull, nk = seed * 17ull;
  unsigned m = seed, m2 = seed + 1, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      u64 t = ffma2(x[i], sdy, nk);
      u64 c = ffma2(y[i], sdx, t);
      u64 d = ffma2(c, c, r2[i]);
      unsigned d0 = (unsigned)d, d1 = (unsigned)(d >> 32);
      x[i] = d; // loop-carried: nothing is invariant
      if (KIND == 0) { // current: 2 SHF
        m = shf(d0, m);
        m = shf(d1, m);
      } else if (KIND == 1) { // FMNMX + 1 SHF
        m = shf(fmax_(d0, d1), m);
      } else if (KIND == 2) { // group of 4: 3 FMNMX + 1 SHF per two pairs
        acc = (i & 1) ? fmax_(acc, fmax_(d0, d1)) : fmax_(d0, d1);
        if (i & 1) m = shf(acc, m);
      } else if (KIND == 3) { // none
        m ^= d0 ^ d1;  // LOP3 (one)
      } else if (KIND == 4) { // integer max (sign-magnitude trick) + SHF
        m = shf(imax_(d0, d1), m);
      } else if (KIND == 5) { // two chains of SHF (two masks: even / odd objects)
        m = shf(d0, m);
        m2 = shf(d1, m2);
      } else if (KIND == 6) { // group of 8: 7 FMNMX + 1 SHF per four pairs
        acc = (i & 3) ? fmax_(acc, fmax_(d0, d1)) : fmax_(d0, d1);
        if ((i & 3) == 3) m = shf(acc, m);
      } else if (KIND == 7) { // pure FFMA2 (the sink keeps them alive through one LOP per 8 pairs)
        if (i == 7) m ^= d0 ^ d1;
      }
    }
  }
  long long t1 = clock64();
  unsigned s = m ^ m2 ^ acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= (unsigned)x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int KIND> void run(const char *name) {
  unsigned *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int ctas_per_sm[] = {1, 2, 3, 4, 6};
  printf("%-32s", name);
  for (int c : ctas_per_sm) {
    k<KIND><<<148 * c, 256>>>(out, 1, cyc);
    cudaEventRecord(e0);
    k<KIND><<<148 * c, 256>>>(out, 1, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double pairs_per_smsp = (double)ITERS * 8 * (c * 8) / 4.0; // warp-level pairs per SMSP
    printf("  %dw: %5.2f", c * 8, ms * 1e-3 * 1.965e9 / pairs_per_smsp);
  }
  printf("   cycles/pair (at 1965 MHz)\n");
  cudaFree(out);
  cudaFree(cyc);
}
int main() {
  run<7>("3 FFMA2 only");
  run<3>("3 FFMA2 + LOP3");
  run<0>("3 FFMA2 + 2 SHF (current)");
  run<5>("3 FFMA2 + 2 SHF two chains");
  run<1>("3 FFMA2 + FMNMX + SHF");
  run<4>("3 FFMA2 + IMNMX + SHF");
  run<2>("3 FFMA2 + 1.5 FMNMX + .5 SHF");
  run<6>("3 FFMA2 + 1.75 FMNMX + .25 SHF");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
