// The f64 trace kernel in its large-scene form (shared narrow phase, kLarge table loads; R = 1, 2 slots).
#include "lg_trace.cuh"
namespace lg {
const void *trace_kernel_f64_large(int slots, bool smem) {
  if (slots == 2)
    return smem ? (const void *)trace_kernel<double, 2, true, false, true, true>
                : (const void *)trace_kernel<double, 2, false, false, true, true>;
  return smem ? (const void *)trace_kernel<double, 1, true, false, true, true>
              : (const void *)trace_kernel<double, 1, false, false, true, true>;
}
} // namespace lg
