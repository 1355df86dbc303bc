// Co-issue microbenchmark (sm_100a): NF packed FMAs and NA ops of another kind per iteration, independent
// self-dependent chains, interleaved; reports cycles per iteration per SM sub-partition warp.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
enum { SHF, LOP, FMNMX, IADD, VIMNMX, PRMT, FADD, FFMA1, NONE };
template <int OP> __device__ __forceinline__ unsigned alu(unsigned s, unsigned b) {
  if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(s) : "r"(b));
  if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %1, 0x6a;" : "+r"(s) : "r"(b));
  if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+r"(s) : "r"(b));
  if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(s) : "r"(b));
  if (OP == VIMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(s) : "r"(b));
  if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0xb9b9;" : "+r"(s) : "r"(b));
  if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(s) : "r"(b));
  if (OP == FFMA1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(s) : "r"(b));
  return s;
}
template <int NF, int NA, int OP> __global__ void k(unsigned *out, unsigned seed) {
  u64 a2[8], bb = ((u64)(seed * 3) << 32) | (seed * 5);
  unsigned s[8], b = seed * 7 + 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) a2[i] = ((u64)(seed + i) << 32) | (threadIdx.x + i), s[i] = seed * 11 + threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
    // NF + NA slots, ALU ops spread evenly among the packed FMAs
    int fa = 0, aa = 0;
#pragma unroll
    for (int q = 0; q < NF + NA; ++q) {
      // Bresenham interleave
      const bool do_alu = NA > 0 && ((q + 1) * NA / (NF + NA) > q * NA / (NF + NA));
      if (do_alu) {
        s[aa & 7] = alu<OP>(s[aa & 7], b);
        ++aa;
      } else {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a2[fa & 7]) : "l"(bb));
        ++fa;
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= s[i] ^ (unsigned)a2[i] ^ (unsigned)(a2[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NF, int NA, int OP> void run(const char *name) {
  unsigned *out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  printf("%-8s NF=%2d NA=%2d :", name, NF, NA);
  for (int c : {2, 3, 4, 6}) {
    k<NF, NA, OP><<<148 * c, 256>>>(out, 1);
    cudaEventRecord(e0);
    k<NF, NA, OP><<<148 * c, 256>>>(out, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double iters_per_smsp = (double)ITERS * (c * 8) / 4.0;
    printf("  %2dw: %6.2f", c * 8, ms * 1e-3 * 1.965e9 / iters_per_smsp);
  }
  printf("   cycles/iter\n");
  cudaFree(out);
}
int main() {
  run<24, 0, NONE>("FFMA2");
  run<24, 16, SHF>("SHF");
  run<24, 8, SHF>("SHF");
  run<24, 4, SHF>("SHF");
  run<24, 16, LOP>("LOP3");
  run<24, 8, LOP>("LOP3");
  run<24, 16, PRMT>("PRMT");
  run<24, 16, FMNMX>("FMNMX");
  run<24, 8, FMNMX>("FMNMX");
  run<24, 16, IADD>("IADD");
  run<24, 8, IADD>("IADD");
  run<24, 16, VIMNMX>("VIMNMX");
  run<24, 16, FADD>("FADD");
  run<24, 16, FFMA1>("FFMA");
  run<0, 16, SHF>("SHF");
  run<0, 16, FMNMX>("FMNMX");
  run<0, 16, IADD>("IADD");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
