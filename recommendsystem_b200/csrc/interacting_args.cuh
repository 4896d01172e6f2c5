// interacting_args.cuh — argument packs shared by the K4 dispatcher and the
// per-shape instantiations.
#pragma once
#include "common.cuh"
namespace rs {
struct IFwdArgs {
  const void* x; int64_t x_ld; const float* W; const float* b; const float* gm; const float* bt;
  float eps; void* y; int64_t y_ld; void* saved; int B, F, L, use_res, dtype; cudaStream_t st;
};
struct IBwdArgs {
  const void* x; int64_t x_ld; const void* saved; const float* W; const float* b; const float* gm;
  const float* bt; float eps; const void* dy; int64_t dy_ld; void* dx; int64_t dx_ld;
  float* dparams; int B, F, L, use_res, dtype; void* ws; size_t ws_bytes; cudaStream_t st;
};
}  // namespace rs
