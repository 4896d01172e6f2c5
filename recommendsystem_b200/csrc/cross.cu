// cross.cu — DCN-v1 cross network, forward and backward, as HBM-bound row kernels.
//
// Reference: CrossNet.call (rough_rank/layer.py:256-264) and DeepCrossLayer.call (staytime/layer.py:66-72) — the
// same recurrence written twice:        x_{l+1} = x0 * (x_l . w_l) + b_l + x_l ,   x_0 = x0 .
//
// The recurrence never leaves span{x0, b_0..b_{l-1}}:   x_l = x0 (1 + S_l) + bsum_l ,  S_l = sum_{k<l} s_k ,
// bsum_l = sum_{k<l} b_k ,  s_l = x_l . w_l = (1 + S_l) a_l + beta_l  with  a_l = x0 . w_l ,  beta_l = bsum_l . w_l .
// So a sample needs L dot products over its row and a scalar recursion; TensorFlow materialises L tensors of the
// row's size.  One warp owns a row: pass 1 = the dots (the row streams from HBM once), pass 2 = the output (the row
// comes back from L1/L2).  Algorithmic bytes per sample: dim * (in + out).
//
// Backward with g_l = d x_l (g_L = dout):  ds_l = x0 . g_{l+1} = p + sum_{k>l} a_k ds_k  (p = x0 . dout),
//   dx0 = (1 + S_L) dout + sum_k gamma_k w_k ,  gamma_k = (1 + S_k) ds_k ,
//   dW[:, l] = sum_b gamma_l[b] x0[b, :] + bsum_l sum_b ds_l[b] ,   db_l = colsum(dout) + sum_{k>l} w_k sum_b ds_k[b] .
// Kernel A (warp per row) writes dx0 and the per-sample scalars gamma, ds; kernel B reduces the weighted column
// sums over fixed batch chunks; kernel C folds the chunks in a fixed order and adds the rank-one terms:
// deterministic, no atomics.
#include "common.cuh"
#include <algorithm>

namespace rs {

constexpr int CROSS_MAX_L = 8;
constexpr int CROSS_CHUNK = 512;        // batch rows per partial of the weight-gradient reduction

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ws layout (floats): beta[CROSS_MAX_L] | bsum[(L+1)][dim]
__global__ void cross_prep_kernel(const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ ws,
                                  int dim, int L) {
  float* beta = ws;
  float* bsum = ws + CROSS_MAX_L;
  __shared__ float red[32];
  float acc[CROSS_MAX_L];
#pragma unroll
  for (int l = 0; l < CROSS_MAX_L; ++l) acc[l] = 0.f;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float run = 0.f;
    for (int l = 0; l < L; ++l) {
      bsum[(int64_t)l * dim + c] = run;
      acc[l] = fmaf(run, W[(int64_t)l * dim + c], acc[l]);
      run += b[(int64_t)l * dim + c];
    }
    bsum[(int64_t)L * dim + c] = run;
  }
  for (int l = 0; l < L; ++l) {
    float v = warp_sum(acc[l]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) beta[l] = t;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cross_fwd_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ W, const float* __restrict__ ws,
                 T* __restrict__ out, int64_t out_ld, int B, int dim, int L) {
  const float* beta = ws;
  const float* bsumL = ws + CROSS_MAX_L + (int64_t)L * dim;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += (int64_t)gridDim.x * wpb) {
    const T* xr = x + r * x_ld;
    // dot products: 4-term fp32 groups added into fp64 accumulators (the dots cancel heavily for wide rows; the
    // kernel is HBM bound, the few DADDs are free) ; scalar recursion in fp64
    double a[CROSS_MAX_L];
#pragma unroll
    for (int l = 0; l < CROSS_MAX_L; ++l) a[l] = 0.0;
    for (int c = lane * 4; c < dim; c += 128) {
      const float4 v = load4<T>(xr + c);
#pragma unroll
      for (int l = 0; l < CROSS_MAX_L; ++l)
        if (l < L) {
          const float4 w = *reinterpret_cast<const float4*>(W + (int64_t)l * dim + c);
          a[l] += (double)fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, v.w * w.w)));
        }
    }
    double S = 0.0;
#pragma unroll
    for (int l = 0; l < CROSS_MAX_L; ++l)
      if (l < L) S += (1.0 + S) * warp_sum_d(a[l]) + (double)beta[l];
    const float k = (float)(1.0 + S);
    T* orow = out + r * out_ld;
    for (int c = lane * 4; c < dim; c += 128) {
      const float4 v = load4<T>(xr + c);
      const float4 bs = *reinterpret_cast<const float4*>(bsumL + c);
      store4<T>(orow + c, make_float4(fmaf(v.x, k, bs.x), fmaf(v.y, k, bs.y), fmaf(v.z, k, bs.z), fmaf(v.w, k, bs.w)));
    }
  }
}

// kernel A of the backward: dx0 rows and the per-sample scalars gamma[B][L], ds[B][L]
template <typename T>
__global__ void __launch_bounds__(256)
cross_bwd_rows_kernel(const T* __restrict__ x, int64_t x_ld, const T* __restrict__ dout, int64_t do_ld,
                      const float* __restrict__ W, const float* __restrict__ ws, T* __restrict__ dx, int64_t dx_ld,
                      float* __restrict__ gam, float* __restrict__ dsv, int B, int dim, int L) {
  const float* beta = ws;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += (int64_t)gridDim.x * wpb) {
    const T* xr = x + r * x_ld;
    const T* gr = dout + r * do_ld;
    double a[CROSS_MAX_L], p = 0.0;
#pragma unroll
    for (int l = 0; l < CROSS_MAX_L; ++l) a[l] = 0.0;
    for (int c = lane * 4; c < dim; c += 128) {
      const float4 v = load4<T>(xr + c);
      const float4 g = load4<T>(gr + c);
      p += (double)fmaf(v.x, g.x, fmaf(v.y, g.y, fmaf(v.z, g.z, v.w * g.w)));
#pragma unroll
      for (int l = 0; l < CROSS_MAX_L; ++l)
        if (l < L) {
          const float4 w = *reinterpret_cast<const float4*>(W + (int64_t)l * dim + c);
          a[l] += (double)fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, v.w * w.w)));
        }
    }
    p = warp_sum_d(p);
    double Sk[CROSS_MAX_L + 1];
    float ds[CROSS_MAX_L], gm[CROSS_MAX_L];
    Sk[0] = 0.0;
#pragma unroll
    for (int l = 0; l < CROSS_MAX_L; ++l) {
      a[l] = l < L ? warp_sum_d(a[l]) : 0.0;
      Sk[l + 1] = l < L ? Sk[l] + (1.0 + Sk[l]) * a[l] + (double)beta[l] : Sk[l];
    }
    double tail = 0.0;                      // sum_{k>l} a_k ds_k
#pragma unroll
    for (int l = CROSS_MAX_L - 1; l >= 0; --l) {
      const double d = l < L ? p + tail : 0.0;
      ds[l] = (float)d;
      gm[l] = (float)((1.0 + Sk[l]) * d);
      tail += a[l] * d;
    }
    const float alpha = (float)(1.0 + Sk[CROSS_MAX_L]);
    if (lane < L) {
      float gv = 0.f, dv = 0.f;
#pragma unroll
      for (int l = 0; l < CROSS_MAX_L; ++l)
        if (l == lane) { gv = gm[l]; dv = ds[l]; }
      gam[r * L + lane] = gv;
      dsv[r * L + lane] = dv;
    }
    T* drow = dx + r * dx_ld;
    for (int c = lane * 4; c < dim; c += 128) {
      const float4 g = load4<T>(gr + c);
      float4 o = make_float4(alpha * g.x, alpha * g.y, alpha * g.z, alpha * g.w);
#pragma unroll
      for (int l = 0; l < CROSS_MAX_L; ++l)
        if (l < L) {
          const float4 w = *reinterpret_cast<const float4*>(W + (int64_t)l * dim + c);
          o.x = fmaf(gm[l], w.x, o.x); o.y = fmaf(gm[l], w.y, o.y); o.z = fmaf(gm[l], w.z, o.z); o.w = fmaf(gm[l], w.w, o.w);
        }
      store4<T>(drow + c, o);
    }
  }
}

// kernel B: per (column, batch chunk): sum_b gamma_l[b] x0[b, col] (l < L) and sum_b dout[b, col]; thread = column
template <typename T>
__global__ void __launch_bounds__(128)
cross_wgrad_partial_kernel(const T* __restrict__ x, int64_t x_ld, const T* __restrict__ dout, int64_t do_ld,
                           const float* __restrict__ gam, float* __restrict__ part /*[chunks][L+1][dim]*/, int B,
                           int dim, int L) {
  __shared__ float gs[CROSS_CHUNK * CROSS_MAX_L];
  const int chunk = blockIdx.y;
  const int r0 = chunk * CROSS_CHUNK, r1 = min(B, r0 + CROSS_CHUNK);
  for (int i = threadIdx.x; i < (r1 - r0) * L; i += blockDim.x) gs[i] = gam[(int64_t)r0 * L + i];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dim) return;
  double acc[CROSS_MAX_L], accd = 0.0;
#pragma unroll
  for (int l = 0; l < CROSS_MAX_L; ++l) acc[l] = 0.0;
  for (int r = r0; r < r1; ++r) {
    const float v = to_f<T>(x[(int64_t)r * x_ld + c]);
    accd += (double)to_f<T>(dout[(int64_t)r * do_ld + c]);
#pragma unroll
    for (int l = 0; l < CROSS_MAX_L; ++l)
      if (l < L) acc[l] += (double)(gs[(r - r0) * L + l] * v);
  }
  float* p = part + (int64_t)chunk * (L + 1) * dim;
  for (int l = 0; l < L; ++l) p[(int64_t)l * dim + c] = (float)acc[l];
  p[(int64_t)L * dim + c] = (float)accd;
}

// kernel C: fold the chunks in order; sums of ds; rank-one terms
__global__ void cross_wgrad_final_kernel(const float* __restrict__ part, const float* __restrict__ dsv,
                                         const float* __restrict__ W, const float* __restrict__ ws,
                                         float* __restrict__ dW, float* __restrict__ db, int B, int dim, int L,
                                         int chunks) {
  const float* bsum = ws + CROSS_MAX_L;
  __shared__ float sds[CROSS_MAX_L];
  __shared__ float red[32];
  // sum_b ds_l[b], fixed order: thread-strided partials then a fixed tree
  for (int l = 0; l < L; ++l) {
    float s = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) s += dsv[(int64_t)r * L + l];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) sds[l] = t;
    }
    __syncthreads();
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += gridDim.x * blockDim.x) {
    float cd = 0.f;
    for (int k = 0; k < chunks; ++k) cd += part[((int64_t)k * (L + 1) + L) * dim + c];
    float tail = 0.f;                      // sum_{k>l} w_k[c] sum_b ds_k[b]
    for (int l = L - 1; l >= 0; --l) {
      float g = 0.f;
      for (int k = 0; k < chunks; ++k) g += part[((int64_t)k * (L + 1) + l) * dim + c];
      dW[(int64_t)l * dim + c] = fmaf(bsum[(int64_t)l * dim + c], sds[l], g);
      db[(int64_t)l * dim + c] = cd + tail;
      tail = fmaf(W[(int64_t)l * dim + c], sds[l], tail);
    }
  }
}

}  // namespace rs

using namespace rs;

extern "C" {

size_t rs_cross_workspace_bytes(int B, int dim, int L) {
  const size_t chunks = (size_t)cdiv(B, CROSS_CHUNK);
  return sizeof(float) * ((size_t)CROSS_MAX_L + (size_t)(L + 1) * dim        /* beta | bsum */
                          + 2 * (size_t)B * L                                  /* gamma | ds */
                          + chunks * (size_t)(L + 1) * dim);                   /* weight-gradient partials */
}

static int cross_check(const char* what, int B, int dim, int L, int dtype, int64_t ld0, int64_t ld1, size_t ws_bytes) {
  RS_REQUIRE(B > 0 && dim > 0 && dim % 4 == 0, "%s: B=%d dim=%d (dim must be a multiple of 4)", what, B, dim);
  RS_REQUIRE(L >= 1 && L <= CROSS_MAX_L, "%s: %d cross layers (1..%d built)", what, L, CROSS_MAX_L);
  RS_REQUIRE(dtype == RS_F32 || dtype == RS_BF16, "%s: bad dtype", what);
  RS_REQUIRE(ld0 % 4 == 0 && ld1 % 4 == 0 && ld0 >= dim && ld1 >= dim, "%s: leading dims must be multiples of 4 >= dim", what);
  if (ws_bytes < rs_cross_workspace_bytes(B, dim, L)) {
    set_error("%s: workspace %zu < %zu", what, ws_bytes, rs_cross_workspace_bytes(B, dim, L));
    return RS_ERR_WORKSPACE;
  }
  return 0;
}

int rs_cross_fwd(const void* x, int64_t x_ld, int dtype, const float* W, const float* b, void* out, int64_t out_ld,
                 int B, int dim, int L, void* ws, size_t ws_bytes, void* stream) {
  if (int e = cross_check("cross_fwd", B, dim, L, dtype, x_ld, out_ld, ws_bytes)) return e;
  cudaStream_t st = as_stream(stream);
  float* wsf = (float*)ws;
  cross_prep_kernel<<<1, 256, 0, st>>>(W, b, wsf, dim, L);
  if (int e = check_launch("cross_prep")) return e;
  const int grid = (int)std::min<int64_t>(cdiv(B, 8), (int64_t)sm_count() * 8);
  if (dtype == RS_F32)
    cross_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, x_ld, W, wsf, (float*)out, out_ld, B, dim, L);
  else
    cross_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, x_ld, W, wsf, (__nv_bfloat16*)out, out_ld,
                                                        B, dim, L);
  return check_launch("cross_fwd");
}

int rs_cross_bwd(const void* x, int64_t x_ld, const void* dout, int64_t do_ld, int dtype, const float* W,
                 const float* b, void* dx, int64_t dx_ld, float* dW, float* db, int B, int dim, int L, void* ws,
                 size_t ws_bytes, void* stream) {
  if (int e = cross_check("cross_bwd", B, dim, L, dtype, x_ld, do_ld, ws_bytes)) return e;
  RS_REQUIRE(dx_ld % 4 == 0 && dx_ld >= dim, "cross_bwd: dx_ld");
  cudaStream_t st = as_stream(stream);
  float* wsf = (float*)ws;
  float* gam = wsf + CROSS_MAX_L + (size_t)(L + 1) * dim;
  float* dsv = gam + (size_t)B * L;
  float* part = dsv + (size_t)B * L;
  const int chunks = (int)cdiv(B, CROSS_CHUNK);
  cross_prep_kernel<<<1, 256, 0, st>>>(W, b, wsf, dim, L);
  if (int e = check_launch("cross_prep")) return e;
  const int grid = (int)std::min<int64_t>(cdiv(B, 8), (int64_t)sm_count() * 8);
  const dim3 g2((unsigned)cdiv(dim, 128), (unsigned)chunks);
  if (dtype == RS_F32) {
    cross_bwd_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)x, x_ld, (const float*)dout, do_ld, W, wsf, (float*)dx,
                                                       dx_ld, gam, dsv, B, dim, L);
    if (int e = check_launch("cross_bwd_rows")) return e;
    cross_wgrad_partial_kernel<float><<<g2, 128, 0, st>>>((const float*)x, x_ld, (const float*)dout, do_ld, gam, part, B,
                                                          dim, L);
  } else {
    using H = __nv_bfloat16;
    cross_bwd_rows_kernel<H><<<grid, 256, 0, st>>>((const H*)x, x_ld, (const H*)dout, do_ld, W, wsf, (H*)dx, dx_ld, gam,
                                                   dsv, B, dim, L);
    if (int e = check_launch("cross_bwd_rows")) return e;
    cross_wgrad_partial_kernel<H><<<g2, 128, 0, st>>>((const H*)x, x_ld, (const H*)dout, do_ld, gam, part, B, dim, L);
  }
  if (int e = check_launch("cross_wgrad_partial")) return e;
  cross_wgrad_final_kernel<<<(unsigned)std::min<int64_t>(cdiv(dim, 256), 64), 256, 0, st>>>(part, dsv, W, wsf, dW, db, B, dim,
                                                                                          L, chunks);
  return check_launch("cross_wgrad_final");
}

}  // extern "C"
