// gemm_tc.cu — bf16 tensor-core GEMM (tcgen05 + TMEM + TMA) for the MLP tower.
// Placeholder until the tcgen05 kernel lands: reports RS_ERR_UNSUPPORTED so that
// callers fail loudly instead of silently falling back.
#include "common.cuh"
namespace rs {
int gemm_bf16_tc(const void*, int64_t, int, const void*, int64_t, int, void*, int64_t, const float*,
                 const void*, int64_t, int, int, int, int, int, cudaStream_t) {
  set_error("gemm: bf16 tensor-core path not built yet");
  return RS_ERR_UNSUPPORTED;
}
}  // namespace rs
