// Instantiations of the trace kernel for the f32 throughput mode (R = 1, 2, 4 slots).
#include "lg_trace.cuh"
namespace lg {
const void *trace_kernel_f32(int slots, bool smem) {
  switch (slots) {
  case 1: return smem ? (const void *)trace_kernel<float, 1, true> : (const void *)trace_kernel<float, 1, false>;
  case 4: return smem ? (const void *)trace_kernel<float, 4, true> : (const void *)trace_kernel<float, 4, false>;
  default: return smem ? (const void *)trace_kernel<float, 2, true> : (const void *)trace_kernel<float, 2, false>;
  }
}
} // namespace lg
