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
      } else if (KIND == 8) { // sign bits through the integer adder: m = 2 m + (d >> 31), which ptxas turns into LEA.HI
        unsigned e0, e1;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(e0), "=r"(e1) : "l"(d));
        m = (m << 1) + (e0 >> 31);
        m = (m << 1) + (e1 >> 31);
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
// cr with two packed FMAs, the sign from a scalar FADD (R - |cr|) per test, then the two shifts
__global__ void k_fadd(unsigned *out, unsigned seed, long long *cycles) {
  u64 x[8], y[8];
  float ra[8], rb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = ((u64)(seed + i) << 32) | (threadIdx.x + i);
    y[i] = ((u64)(seed * 3 + i) << 32) | (threadIdx.x * 5 + i);
    ra[i] = (float)(seed * 7 + i), rb[i] = (float)(threadIdx.x * 9 + i);
  }
  u64 sdy = seed * 11ull, sdx = seed * 13ull, nk = seed * 17ull;
  unsigned m = seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      u64 t = ffma2(x[i], sdy, nk);
      u64 c = ffma2(y[i], sdx, t);
      float c0 = __uint_as_float((unsigned)c), c1 = __uint_as_float((unsigned)(c >> 32));
      const float d0 = ra[i] - fabsf(c0), d1 = rb[i] - fabsf(c1); // SASS: FADD d, r, -|c| (abs is a free operand modifier)
      x[i] = c ^ (u64)__float_as_uint(d0); // loop-carried
      m = shf(__float_as_uint(d0), m);
      m = shf(__float_as_uint(d1), m);
    }
  }
  long long t1 = clock64();
  unsigned s = m;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= (unsigned)x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
// lane = object, the ray broadcast: two scalar FMAs, one compare and ONE vote per 32 tests.  A "pair"
// here is the same 64 tests per warp as above: two rays against the lane's object.
__global__ void k_vote(unsigned *out, unsigned seed, long long *cycles) {
  float x[8], y[8], r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (float)(threadIdx.x + i), y[i] = (float)(threadIdx.x * 5 + i), r[i] = (float)(threadIdx.x * 9 + i);
  float sdy0 = seed * 11.f, sdx0 = seed * 13.f, nk0 = seed * 17.f, sdy1 = seed * 3.f, sdx1 = seed * 5.f, nk1 = seed * 7.f;
  unsigned m = seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float c0 = fmaf(y[i], sdx0, fmaf(x[i], sdy0, nk0));
      float c1 = fmaf(y[i], sdx1, fmaf(x[i], sdy1, nk1));
      const unsigned b0 = __ballot_sync(0xffffffffu, fabsf(c0) <= r[i]);
      const unsigned b1 = __ballot_sync(0xffffffffu, fabsf(c1) <= r[i]);
      x[i] = c0 + c1;  // loop-carried (one FADD more than the real thing)
      m ^= b0 + b1;    // the masks are warp-uniform: one uniform-datapath op
    }
  }
  long long t1 = clock64();
  unsigned s = m;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= __float_as_uint(x[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <class K> void run_kernel(const char *name, K kern) {
  unsigned *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int ctas_per_sm[] = {1, 2, 3, 4, 6};
  printf("%-32s", name);
  for (int c : ctas_per_sm) {
    kern<<<148 * c, 256>>>(out, 1, cyc);
    cudaEventRecord(e0);
    kern<<<148 * c, 256>>>(out, 1, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double pairs_per_smsp = (double)ITERS * 8 * (c * 8) / 4.0;
    printf("  %dw: %5.2f", c * 8, ms * 1e-3 * 1.965e9 / pairs_per_smsp);
  }
  printf("   cycles/pair (at 1965 MHz)\n");
  cudaFree(out);
  cudaFree(cyc);
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
  run<8>("3 FFMA2 + 2 LEA.HI (int adder)");
  run_kernel("(a) 2 FFMA2 + 2 FADD|.| + 2 SHF", k_fadd);
  run_kernel("(b) lane=object: 4 FFMA+2 FSETP+2 VOTE", k_vote);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
