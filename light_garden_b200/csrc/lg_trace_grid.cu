// Instantiations of the trace kernel that find the nearest hit through the uniform grid (lg_tile_map_enable).
#include "lg_trace.cuh"
namespace lg {
const void *trace_kernel_grid_f32(int slots) {
  return slots == 1 ? (const void *)trace_kernel<float, 1, false, true> : (const void *)trace_kernel<float, 2, false, true>;
}
const void *trace_kernel_grid_f64(int) { return (const void *)trace_kernel<double, 1, false, true>; }
} // namespace lg
