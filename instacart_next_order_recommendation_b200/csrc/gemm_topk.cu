// K2 placeholder (replaced by the tcgen05 implementation).
#include "common.cuh"
namespace icr {
bool gemm_topk_supported(int64_t, int64_t, int64_t, int, int, const uint8_t*) { return false; }
size_t gemm_topk_workspace_bytes(int64_t, int64_t, int64_t, int, int, int) { return 256; }
int launch_gemm_topk(const void*, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, const uint16_t*, const uint8_t*,
                     int, int64_t, float*, int64_t*, void*, size_t, cudaStream_t) {
  set_error("GEMM path not built");
  return ICR_ERR_ARG;
}
}  // namespace icr
