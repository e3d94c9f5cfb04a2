// tcgen05/TMEM assignment kernel (3xTF32, TMA-fed) — placeholder until the kernel lands.
#include "common.cuh"

namespace vqgnn {
int launch_assign_tc(const float*, int64_t, const float*, int64_t, const float*, const float*, const float*,
                     int64_t, int, int, int, int, int, const int32_t*, int16_t*, int64_t, int16_t*, float*,
                     cudaStream_t) {
  set_error("vq_assign: impl=1 (tcgen05) is not available in this build");
  return VQGNN_ERR_ARG;
}
}  // namespace vqgnn
