// interacting_args.cuh — argument packs shared by the K4 dispatcher and the
// per-shape instantiations.
#pragma once
#include "common.cuh"
namespace rs {
struct IFwdArgs {
  const void* x; int64_t x_ld, x_bs; const float* W; const float* b; const float* gm; const float* bt;
  float eps; void* y; int64_t y_ld, y_bs; void* saved; int B, F, L, use_res, dtype; cudaStream_t st;
  float drop_rate = 0.f; unsigned long long drop_seed = 0; const unsigned long long* drop_step = nullptr;
};
struct IBwdArgs {
  const void* x; int64_t x_ld, x_bs; const void* saved; const float* W; const float* b; const float* gm;
  const float* bt; float eps; const void* dy; int64_t dy_ld, dy_bs; void* dx; int64_t dx_ld, dx_bs;
  float* dparams; int B, F, L, use_res, dtype; void* ws; size_t ws_bytes; cudaStream_t st;
  float drop_rate = 0.f; unsigned long long drop_seed = 0; const unsigned long long* drop_step = nullptr;
};
}  // namespace rs
