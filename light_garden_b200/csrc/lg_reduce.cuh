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

// `skip`: set by the barrier in front of it when some rank asked for the handle exchange (its buffers moved: the peer
// pointers here may be stale) or did not arrive -- the kernel then does nothing and the host repeats the reduce.
__global__ void __launch_bounds__(256) reduce_finalize_peer_kernel(PeerPtrs P, size_t px0, size_t px1, const unsigned int *skip) {
  if (skip && *skip) return;
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

// ---- barriers between the ranks without NCCL and without the host: flag words in peer-mapped device memory ----------
// Every rank owns `flags[phase][rank]` (two phases x 16 ranks, 64-bit epochs) and maps every peer's array.  A barrier
// is one tiny kernel on the rank's stream: thread p publishes this rank's value for the current epoch into PEER p's
// array (system-scope release store: everything this rank queued before it on the stream -- its accumulation, or its
// band of the reduce -- is visible first) and then waits until peer p's value for this epoch has arrived in its OWN
// array (acquire loads).  The value carries one payload bit ("this rank needs the handle exchange again"), ORed over
// the ranks into *status.  A peer that never arrives (it is not in this protocol: a context that was re-created) ends
// the wait after `timeout_cycles` with status bit 1, which the host treats like a request for the exchange.
struct PeerFlags {
  unsigned long long *peer[16]; // rank p's flag array (mapped here)
  unsigned long long *mine;     // this rank's array
  int n, rank;
};
constexpr int kFlagPhases = 2;
__global__ void peer_barrier_kernel(PeerFlags F, int phase, unsigned long long epoch, unsigned payload_bit,
                                    unsigned int *status, long long timeout_cycles) {
  const int p = threadIdx.x;
  if (p >= F.n) return;
  const unsigned long long word = (epoch << 1) | (payload_bit & 1u);
  __threadfence_system();
  unsigned long long *dst = F.peer[p] + (size_t)phase * 16 + F.rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
  const unsigned long long *src = F.mine + (size_t)phase * 16 + p;
  const long long t0 = clock64();
  unsigned long long got = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
    if ((got >> 1) >= epoch) break;
    if (clock64() - t0 > timeout_cycles) {
      atomicOr(status, 2u);
      return;
    }
    __nanosleep(200);
  }
  if ((got >> 1) == epoch && (got & 1ull)) atomicOr(status, 1u);
}

} // namespace lg
