// Instantiations of the trace kernel for the f64 reference-width mode (R = 1, 2 slots).
#include "lg_trace.cuh"
namespace lg {
const void *trace_kernel_f64(int slots, bool smem) {
  switch (slots) {
  case 2: return smem ? (const void *)trace_kernel<double, 2, true> : (const void *)trace_kernel<double, 2, false>;
  default: return smem ? (const void *)trace_kernel<double, 1, true> : (const void *)trace_kernel<double, 1, false>;
  }
}
} // namespace lg
