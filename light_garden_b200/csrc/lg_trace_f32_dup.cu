// The f32 trace kernel with one narrow-phase instance per ray slot (R = 2): the variant for large scenes.
#include "lg_trace.cuh"
namespace lg {
const void *trace_kernel_f32_per_slot(bool smem) {
  return smem ? (const void *)trace_kernel<float, 2, true, false, false> : (const void *)trace_kernel<float, 2, false, false, false>;
}
} // namespace lg
