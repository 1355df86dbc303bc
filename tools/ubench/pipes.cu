// Pipe-rate microbenchmark (sm_100a): warp-instructions per clock per SM sub-partition for the instruction kinds the
// trace kernel's broad phase is built from.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
template <int KIND> __global__ void k(unsigned *out, unsigned seed, long long *cycles) {
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
  unsigned b = seed * 3 + 1, c = seed * 5 + 2;
  unsigned a3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a3[i] = seed * 11 + threadIdx.x + i;
  unsigned long long a2[8], bb = ((unsigned long long)b << 32) | c;
#pragma unroll
  for (int i = 0; i < 8; ++i) a2[i] = ((unsigned long long)a[i] << 32) | (a[i] * 7u);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) { // FFMA
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      } else if (KIND == 1) { // FFMA2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a2[i]) : "l"(bb));
      } else if (KIND == 2) { // HFMA2
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      } else if (KIND == 3) { // SHF
        asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 4) { // PRMT
        asm volatile("prmt.b32 %0, %0, %1, 0xb9b9;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 5) { // LOP3
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xea;" : "+r"(a[i]) : "r"(b), "r"(c));
      } else if (KIND == 6) { // HFMA2 + SHF interleaved (dual issue across pipes?)
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(b) : "r"(a[i]));
      } else if (KIND == 7) { // HMNMX2
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 9) { // FADD
        asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 10) { // FMNMX
        asm volatile("max.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 11) { // IADD3
        asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 12) { // FSETP + SEL
        asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a[i]) : "r"(b), "r"(c));
      } else if (KIND == 13) { // FSET (set.gt.u32.f32 -> 0xffffffff / 0)
        asm volatile("set.gt.u32.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      } else if (KIND == 14) { // IMAD
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      } else if (KIND == 15) { // FMNMX3
        asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      } else if (KIND == 16) { // FFMA + SHF independent streams
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a3[i]) : "r"(b));
      } else if (KIND == 17) { // 3 FFMA2 + 2 SHF independent (the broad phase mix)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a2[i]) : "l"(bb));
        if (i % 3 != 2) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a3[i]) : "r"(b));
      } else if (KIND == 18) { // FMNMX + FFMA independent
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        asm volatile("max.f32 %0, %0, %1;" : "+r"(a3[i]) : "r"(b));
      } else if (KIND == 19) { // VHMNMX + FFMA2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a2[i]) : "l"(bb));
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a3[i]) : "r"(b));
      } else if (KIND == 8) { // bf16x2 fma
        asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      }
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i] ^ a3[i] ^ (unsigned)a2[i] ^ (unsigned)(a2[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s ^ b;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int KIND> void run(const char *name, int per_iter) {
  unsigned *out;
  long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<KIND><<<148, warps * 32>>>(out, 1, cyc);
    k<KIND><<<148, warps * 32>>>(out, 1, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double inst = (double)ITERS * 8 * per_iter * warps / 4.0; // warp-instructions per SMSP
    printf("%-10s warps/SM %2d : %.3f warp-inst/clk/SMSP\n", name, warps, inst / (double)h);
  }
  cudaFree(out);
  cudaFree(cyc);
}
int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2", 1);
  run<2>("HFMA2", 1);
  run<8>("BFMA2", 1);
  run<3>("SHF", 1);
  run<4>("PRMT", 1);
  run<5>("LOP3", 1);
  run<7>("HMNMX2", 1);
  run<9>("FADD", 1);
  run<10>("FMNMX", 1);
  run<15>("FMNMX3", 1);
  run<11>("IADD", 1);
  run<12>("FSETP+SEL", 2);
  run<13>("FSET", 1);
  run<14>("IMAD", 1);
  run<16>("FFMA|SHF", 2);
  run<18>("FFMA|FMNMX", 2);
  run<19>("FFMA2|VHMNMX", 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
