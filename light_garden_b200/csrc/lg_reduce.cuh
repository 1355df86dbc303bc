// lg_reduce.cuh — the image reduce of the multi-GPU path fused with K5 (fp16 finalize) over NVLink peer memory.
//
// Every rank holds a full-frame partial fp32 RGBA image.  Instead of ncclReduce to the root followed
// by a separate finalize kernel on the root, every rank runs ONE kernel over its own band of rows: it loads that
// band from every peer's image directly (P2P loads over NVLink / NVSwitch), adds the partial sums in rank order
// (deterministic, unlike a ring), and stores both the fp32 sum and the Rgba16Float pixel straight into the root's
// buffers (P2P stores).  Reduce-scatter, finalize and gather in one pass; every GPU pulls from all its peers at once,
// so no link carries more than 1/N of the frame.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace lg {

struct PeerPtrs {
  const float4 *img[16]; // rank p's partial image (mapped into this process / device)
  int n;
  float4 *root_img;      // where the fp32 sum goes
  uint2 *root_img16;     // where the Rgba16Float frame goes
};

__global__ void __launch_bounds__(256) reduce_finalize_peer_kernel(PeerPtrs P, size_t px0, size_t px1) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = px0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < px1; i += stride) {
    float4 acc = P.img[0][i];
#pragma unroll 4
    for (int p = 1; p < P.n; ++p) {
      const float4 v = P.img[p][i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    P.root_img[i] = acc;
    const __half2 lo = __floats2half2_rn(acc.x, acc.y), hi = __floats2half2_rn(acc.z, acc.w);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned int *>(&lo);
    o.y = *reinterpret_cast<const unsigned int *>(&hi);
    P.root_img16[i] = o;
  }
}

} // namespace lg
