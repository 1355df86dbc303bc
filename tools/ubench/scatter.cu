// Scattered-store microbenchmark behind DESIGN.md's note on the fill pass: how many store REQUESTS per second does L2
// take when every lane of a warp writes somewhere else -- 4 bytes (a pair-list entry), 16 bytes, or a full 32-byte
// sector (a segment record) -- over a buffer much larger than L2, and how much does it help when groups of G
// neighbouring lanes write neighbouring entries (what a spatially sorted segment stream would give the fill pass).
#include <cstdio>
#include <cuda_runtime.h>
template <int BYTES, int G> __global__ void k(unsigned *buf, unsigned long long n_slots, int iters) {
  unsigned long long x = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) / G * 0x9E3779B97F4A7C15ull + 777;
  const unsigned sub = threadIdx.x % G;
  for (int i = 0; i < iters; ++i) {
    x ^= x << 13, x ^= x >> 7, x ^= x << 17;
    const unsigned long long slot = (x % (n_slots - G)) + sub; // G lanes share a base: adjacent slots
    if (BYTES == 4) buf[slot] = (unsigned)i;
    else if (BYTES == 16) reinterpret_cast<uint4 *>(buf)[slot] = make_uint4(i, i, i, i);
    else { uint4 *p = reinterpret_cast<uint4 *>(buf) + 2 * slot; p[0] = make_uint4(i, i, i, i); p[1] = make_uint4(i, i, i, i); }
  }
}
template <int BYTES, int G> void run(unsigned *buf, size_t bytes) {
  const int iters = 256, grid = 148 * 8, block = 256;
  const unsigned long long slots = bytes / BYTES;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<BYTES, G><<<grid, block>>>(buf, slots, iters);
  cudaEventRecord(e0);
  k<BYTES, G><<<grid, block>>>(buf, slots, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double stores = (double)iters * grid * block;
  printf("%2d-byte stores, %2d adjacent lanes per group: %7.1f G stores/s  %7.1f GB/s\n", BYTES, G, stores / ms / 1e6,
         stores * BYTES / ms / 1e6);
}
int main() {
  unsigned *buf;
  const size_t bytes = 8ull << 30;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  run<4, 1>(buf, bytes), run<4, 2>(buf, bytes), run<4, 4>(buf, bytes), run<4, 8>(buf, bytes), run<4, 32>(buf, bytes);
  run<16, 1>(buf, bytes), run<16, 2>(buf, bytes);
  run<32, 1>(buf, bytes), run<32, 4>(buf, bytes);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
