// interacting_kernels.cuh — K4: fused InteractingLayer forward / backward
// (InteractingLayer.py:37-61 of the reference).
//
// Mapping: one THREAD owns one (sample, field) row for the whole layer.  A CTA
// of NT threads holds SPT = NT / F whole samples (F=39 -> 3 samples, 117 of
// 128 threads busy).  The row's x / q / r / o / y live in registers; the only
// cross-row traffic is K and V (and, in the backward, Q and dO), staged in
// shared memory with a padded row stride so that 16-B row writes are
// conflict-free and row reads are warp broadcasts.  The F x F softmax is done
// per row with a chunked online softmax (8 keys per chunk), so nothing of size
// F*F is ever materialised.  All `layer_num` iterations run inside the kernel
// (the reference re-applies the SAME weights, InteractingLayer.py:24-31,41).
//
// Backward recomputes the forward of every iteration from its saved input
// (flash-attention style: row stats m, 1/l and delta = dO.O), then
//   row pass    (thread = query i): dq_i  = scale * sum_j dS_ij k_j
//   column pass (thread = key j):   dk_j  = scale * sum_i dS_ij q_i ,
//                                    dv_j  = sum_i P_ij dO_i
// and finally dx = dZ W^T per row, while dW = X^T dZ, db, dgamma, dbeta are
// accumulated per CTA in registers over all tiles in a fixed order and reduced
// over CTAs by a second kernel (deterministic, no atomics).
#pragma once
#include "interacting_args.cuh"

namespace rs {

constexpr float LOG2E = 1.4426950408889634f;

template <int D, int U>
struct ISmem {
  static constexpr int N4 = 4 * U;
  static constexpr int SA = 4 * U + 4;   // row stride of region A (K|V|Q|dO, later dZ)
  static constexpr int SB = 2 * U + 4;   // row stride of region B (dgamma | dbeta terms)
  static constexpr int SX = D + 4;       // row stride of the x stash
};

// z[4U] = b + x W   (W is [D][4U] in smem)
template <int D, int U>
__device__ __forceinline__ void project_row(const float* __restrict__ Ws,
                                            const float* __restrict__ bs, const float (&x)[D],
                                            float (&z)[4 * U]) {
  constexpr int N4 = 4 * U;
#pragma unroll
  for (int c = 0; c < N4 / 4; ++c) {
    float4 acc = *reinterpret_cast<const float4*>(bs + c * 4);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float4 w = *reinterpret_cast<const float4*>(Ws + d * N4 + c * 4);
      acc.x = fmaf(x[d], w.x, acc.x);
      acc.y = fmaf(x[d], w.y, acc.y);
      acc.z = fmaf(x[d], w.z, acc.z);
      acc.w = fmaf(x[d], w.w, acc.w);
    }
    z[c * 4 + 0] = acc.x; z[c * 4 + 1] = acc.y; z[c * 4 + 2] = acc.z; z[c * 4 + 3] = acc.w;
  }
}

// Attention-weight dropout (InteractingLayer.py:53-54, tf.keras Dropout = inverted dropout): element
// (iteration, sample, head, query i, key j) is kept iff the top 24 bits of a splitmix64 hash of its linear
// index (+ seed) are >= rate * 2^24; kept weights are scaled by 1 / (1 - rate).  Counter-based, so the
// backward (row pass AND column pass) regenerates any element's mask from its indices, and
// oracle/oracle_np.py::dropout_scale reproduces it bit for bit.
struct DropCfg { float rate; float inv_keep; unsigned long long seed; const unsigned long long* step; };
// effective seed of this launch: seed + golden-ratio * (*step) when a device-side step counter is given (a captured
// CUDA graph bakes scalar arguments in; the counter lives in device memory and advances between replays)
__device__ __forceinline__ void drop_resolve(DropCfg& dc) {
  if (dc.step) dc.seed += 0x9E3779B97F4A7C15ULL * (*dc.step);
}

__device__ __forceinline__ float drop_scale(const DropCfg& dc, unsigned long long idx) {
  unsigned long long z = idx + dc.seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  const float u = (float)(unsigned int)(z >> 40) * (1.f / 16777216.f);
  return u >= dc.rate ? dc.inv_keep : 0.f;
}
// linear index of attention element (it, b, h, i, j = 0): add j
__device__ __forceinline__ unsigned long long drop_row_base(int it, long long B, long long b, int H, int h, int F, int i) {
  return ((((unsigned long long)it * (unsigned long long)B + (unsigned long long)b) * H + h) * F + i) * (unsigned long long)F;
}

// Chunked online softmax attention for one query row and one head.
// Krow0 points at K of the sample's first field (head offset applied), rows are
// `stride` floats apart; V sits `voff` floats after K in the same row.
// Returns m (running max of raw*scale in log2 domain), l (sum) and o = P V.
template <int DH>
__device__ __forceinline__ void attn_row_fwd(const float (&q)[DH], const float* __restrict__ Krow0,
                                             int stride, int voff, int F, float scale_log2,
                                             float& m_out, float& l_out, float (&o)[DH],
                                             const DropCfg& dc, unsigned long long drop_base) {
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int e = 0; e < DH; ++e) o[e] = 0.f;
  for (int j0 = 0; j0 < F; j0 += 8) {
    float s[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = min(j0 + c, F - 1);
      const float* kr = Krow0 + j * stride;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(kr + e);
        acc = fmaf(q[e], k4.x, acc);
        acc = fmaf(q[e + 1], k4.y, acc);
        acc = fmaf(q[e + 2], k4.z, acc);
        acc = fmaf(q[e + 3], k4.w, acc);
      }
      s[c] = (j0 + c < F) ? acc * scale_log2 : -INFINITY;
    }
    float cm = s[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) cm = fmaxf(cm, s[c]);
    const float mn = fmaxf(m, cm);
    const float alpha = exp2f(m - mn);
    l *= alpha;
#pragma unroll
    for (int e = 0; e < DH; ++e) o[e] *= alpha;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = min(j0 + c, F - 1);
      float p = exp2f(s[c] - mn);
      l += p;                                                       // the softmax normaliser ignores the mask
      if (dc.rate > 0.f) p *= drop_scale(dc, drop_base + (unsigned long long)j);
      const float* vr = Krow0 + j * stride + voff;
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(vr + e);
        o[e] = fmaf(p, v4.x, o[e]);
        o[e + 1] = fmaf(p, v4.y, o[e + 1]);
        o[e + 2] = fmaf(p, v4.z, o[e + 2]);
        o[e + 3] = fmaf(p, v4.w, o[e + 3]);
      }
    }
    m = mn;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int e = 0; e < DH; ++e) o[e] *= inv;
  m_out = m;
  l_out = inv;
}

template <int N, typename T>
__device__ __forceinline__ void load_row(const T* p, float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N; c += 4) {
    const float4 t = load4<T>(p + c);
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
  }
}
template <int N, typename T>
__device__ __forceinline__ void store_row(T* p, const float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N; c += 4) store4<T>(p + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
}
template <int N, typename T>
__device__ __forceinline__ void round_row(float (&v)[N]) {
  if (sizeof(T) == 2) {
#pragma unroll
    for (int c = 0; c < N; ++c) v[c] = bf16_round(v[c]);
  }
}

// relu(o [+ r]) -> LayerNorm; returns y, and (for the backward) xhat, rstd, t>0 mask.
template <int U>
__device__ __forceinline__ void res_relu_ln(const float (&o)[U], const float (&r)[U], int use_res,
                                            const float* __restrict__ gs, const float* __restrict__ be,
                                            float eps, float (&y)[U], float (&xhat)[U], float& rstd,
                                            uint32_t& tmask) {
  float a[U];
  float mean = 0.f;
  tmask = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float t = use_res ? o[u] + r[u] : o[u];
    if (t > 0.f) tmask |= (1u << u);
    a[u] = fmaxf(t, 0.f);
    mean += a[u];
  }
  mean *= (1.f / U);
  float var = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float c = a[u] - mean;
    var = fmaf(c, c, var);
  }
  var *= (1.f / U);
  rstd = rsqrtf(var + eps);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    xhat[u] = (a[u] - mean) * rstd;
    y[u] = fmaf(xhat[u], gs[u], be[u]);
  }
}

// ------------------------------------------------------------------ forward
template <int D, int U, int H, int NT, typename T>
__global__ void __launch_bounds__(NT)
interacting_fwd_kernel(const T* __restrict__ x, int64_t x_ld, int64_t x_bs, const float* __restrict__ W,
                       const float* __restrict__ bias, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, T* __restrict__ y, int64_t y_ld, int64_t y_bs,
                       float* __restrict__ saved, int B, int F, int L, int use_res, DropCfg dc) {
  static_assert(U <= 32, "tmask is 32 bits");
  drop_resolve(dc);
  constexpr int DH = U / H;
  constexpr int N4 = 4 * U;
  constexpr int SK = 2 * U + 4;  // K|V row stride (forward needs only K and V)
  extern __shared__ float4 smem4[];
  float* Ws = reinterpret_cast<float*>(smem4);
  float* bs = Ws + D * N4;
  float* gs = bs + N4;
  float* be = gs + U;
  float* KV = be + U;  // [NT][SK]

  const int tid = threadIdx.x;
  for (int i = tid; i < D * N4; i += NT) Ws[i] = W[i];
  for (int i = tid; i < N4; i += NT) bs[i] = bias[i];
  for (int i = tid; i < U; i += NT) { gs[i] = gamma[i]; be[i] = beta[i]; }
  __syncthreads();

  const int SPT = NT / F;
  const int rows_per_tile = SPT * F;
  const int64_t total_rows = (int64_t)B * F;
  const int ntiles = (B + SPT - 1) / SPT;
  const int ls = tid / F;
  const float* Ksample = KV + (ls * F) * SK;
  const float scale_log2 = LOG2E / sqrtf((float)DH);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = (int64_t)tile * rows_per_tile + tid;
    const bool active = tid < rows_per_tile && row < total_rows;
    float xr[D];
    // element (b, f, c) lives at ptr[b * bs + f * ld + c]
    const int64_t smp = (int64_t)tile * SPT + ls;
    const int fld = tid - ls * F;
    if (active) load_row<D, T>(x + smp * x_bs + fld * x_ld, xr);
    else {
#pragma unroll
      for (int d = 0; d < D; ++d) xr[d] = 0.f;
    }
    float yv[U];
    for (int it = 0; it < L; ++it) {
      float q[U], r[U];
      {
        float z[N4];
        project_row<D, U>(Ws, bs, xr, z);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          q[u] = fmaxf(z[u], 0.f);
          r[u] = fmaxf(z[3 * U + u], 0.f);
        }
        float* kv = KV + tid * SK;
#pragma unroll
        for (int u = 0; u < 2 * U; u += 4)
          *reinterpret_cast<float4*>(kv + u) =
              make_float4(fmaxf(z[U + u], 0.f), fmaxf(z[U + u + 1], 0.f),
                          fmaxf(z[U + u + 2], 0.f), fmaxf(z[U + u + 3], 0.f));
      }
      __syncthreads();
      if (active) {
        float o[U];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float qh[DH], oh[DH], m, linv;
#pragma unroll
          for (int e = 0; e < DH; ++e) qh[e] = q[h * DH + e];
          attn_row_fwd<DH>(qh, Ksample + h * DH, SK, U, F, scale_log2, m, linv, oh, dc,
                           drop_row_base(it, B, smp, H, h, F, fld));
#pragma unroll
          for (int e = 0; e < DH; ++e) o[h * DH + e] = oh[e];
        }
        float xhat[U], rstd;
        uint32_t tmask;
        res_relu_ln<U>(o, r, use_res, gs, be, eps, yv, xhat, rstd, tmask);
      }
      __syncthreads();
      if (it + 1 < L) {
        if (active && saved) store_row<U, float>(saved + ((int64_t)it * total_rows + row) * U, yv);
        if constexpr (D == U) {
#pragma unroll
          for (int u = 0; u < U; ++u) xr[u] = yv[u];
        }
      }
    }
    if (active) store_row<U, T>(y + smp * y_bs + fld * y_ld, yv);
  }
}

// ----------------------------------------------------------------- backward
template <int D, int U, int H, int NT, typename T>
__global__ void __launch_bounds__(NT)
interacting_bwd_kernel(const T* __restrict__ x, int64_t x_ld, int64_t x_bs, const float* __restrict__ saved,
                       const float* __restrict__ W, const float* __restrict__ bias,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       const T* __restrict__ dy, int64_t dy_ld, int64_t dy_bs, T* __restrict__ dx, int64_t dx_ld,
                       int64_t dx_bs,
                       float* __restrict__ part, int B, int F, int L, int use_res, DropCfg dc) {
  drop_resolve(dc);
  constexpr int DH = U / H;
  constexpr int N4 = 4 * U;
  using S = ISmem<D, U>;
  constexpr int SA = S::SA, SB = S::SB, SX = S::SX;
  constexpr int OPT = (D * N4 + NT - 1) / NT;          // dW outputs per thread
  static_assert((D * N4) % NT == 0, "dW outputs must tile the CTA");
  static_assert(N4 % OPT == 0, "a thread's dW outputs stay in one W row");
  constexpr int NCS = (6 * U + NT - 1) / NT;           // column sums per thread
  extern __shared__ float4 smem4[];
  float* Ws = reinterpret_cast<float*>(smem4);
  float* bs = Ws + D * N4;
  float* gs = bs + N4;
  float* be = gs + U;
  float* A = be + U;             // [NT][SA]
  float* Bt = A + NT * SA;       // [NT][SB]
  float* Xs = Bt + NT * SB;      // [NT][SX]
  float* St = Xs + NT * SX;      // [NT][3H] m, 1/l, delta

  const int tid = threadIdx.x;
  for (int i = tid; i < D * N4; i += NT) Ws[i] = W[i];
  for (int i = tid; i < N4; i += NT) bs[i] = bias[i];
  for (int i = tid; i < U; i += NT) { gs[i] = gamma[i]; be[i] = beta[i]; }
  __syncthreads();

  const int SPT = NT / F;
  const int rows_per_tile = SPT * F;
  const int64_t total_rows = (int64_t)B * F;
  const int ntiles = (B + SPT - 1) / SPT;
  const int ls = tid / F;
  const int sbase = ls * F;
  const float scale = 1.f / sqrtf((float)DH);
  const float scale_log2 = LOG2E * scale;

  float dWacc[OPT];
#pragma unroll
  for (int k = 0; k < OPT; ++k) dWacc[k] = 0.f;
  float csacc[NCS];
#pragma unroll
  for (int k = 0; k < NCS; ++k) csacc[k] = 0.f;
  const int w_d = (tid * OPT) / N4;   // dW row (input dim) this thread owns
  const int w_u0 = (tid * OPT) % N4;  // first output column

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = (int64_t)tile * rows_per_tile + tid;
    const bool active = tid < rows_per_tile && row < total_rows;
    const int rows_here = (int)min((int64_t)rows_per_tile, total_rows - (int64_t)tile * rows_per_tile);
    float g[U];  // gradient wrt the output of the current iteration
    const int64_t smp = (int64_t)tile * SPT + ls;
    const int fld = tid - ls * F;
    if (active) load_row<U, T>(dy + smp * dy_bs + fld * dy_ld, g);
    else {
#pragma unroll
      for (int u = 0; u < U; ++u) g[u] = 0.f;
    }

    for (int it = L - 1; it >= 0; --it) {
      float xr[D];
      if (active) {
        if (it == 0) load_row<D, T>(x + smp * x_bs + fld * x_ld, xr);
        else {
          if constexpr (D == U) load_row<U, float>(saved + ((int64_t)(it - 1) * total_rows + row) * U, xr);
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) xr[d] = 0.f;
      }
      float* arow = A + tid * SA;
      float q[U];
      uint32_t rmask = 0, tmask = 0;
      float dT[U];
      {
        float r[U];
        {
          float z[N4];
          project_row<D, U>(Ws, bs, xr, z);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            q[u] = fmaxf(z[u], 0.f);
            r[u] = fmaxf(z[3 * U + u], 0.f);
            if (z[3 * U + u] > 0.f) rmask |= (1u << u);
          }
#pragma unroll
          for (int u = 0; u < 2 * U; u += 4)
            *reinterpret_cast<float4*>(arow + u) =
                make_float4(fmaxf(z[U + u], 0.f), fmaxf(z[U + u + 1], 0.f),
                            fmaxf(z[U + u + 2], 0.f), fmaxf(z[U + u + 3], 0.f));
        }
        __syncthreads();  // K, V visible
        float o[U];
        float mh[H], lh[H];
        if (active) {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            float qh[DH], oh[DH];
#pragma unroll
            for (int e = 0; e < DH; ++e) qh[e] = q[h * DH + e];
            attn_row_fwd<DH>(qh, A + sbase * SA + h * DH, SA, U, F, scale_log2, mh[h], lh[h], oh, dc,
                             drop_row_base(it, B, smp, H, h, F, fld));
#pragma unroll
            for (int e = 0; e < DH; ++e) o[h * DH + e] = oh[e];
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) o[u] = 0.f;
#pragma unroll
          for (int h = 0; h < H; ++h) { mh[h] = 0.f; lh[h] = 0.f; }
        }
        // LayerNorm backward
        float yv[U], xhat[U], rstd;
        res_relu_ln<U>(o, r, use_res, gs, be, eps, yv, xhat, rstd, tmask);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float gg = g[u] * gs[u];
          s1 += gg;
          s2 = fmaf(gg, xhat[u], s2);
        }
        s1 *= (1.f / U);
        s2 *= (1.f / U);
        float* brow = Bt + tid * SB;
#pragma unroll
        for (int u = 0; u < U; u += 4) {
          *reinterpret_cast<float4*>(brow + u) =
              active ? make_float4(g[u] * xhat[u], g[u + 1] * xhat[u + 1], g[u + 2] * xhat[u + 2],
                                   g[u + 3] * xhat[u + 3])
                     : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(brow + U + u) =
              active ? make_float4(g[u], g[u + 1], g[u + 2], g[u + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float dA = (g[u] * gs[u] - s1 - xhat[u] * s2) * rstd;
          dT[u] = (active && ((tmask >> u) & 1u)) ? dA : 0.f;
        }
        // delta_h = dO_h . o_h ; publish Q, dO and the row stats
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float dl = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) dl = fmaf(dT[h * DH + e], o[h * DH + e], dl);
          St[tid * 3 * H + h * 3 + 0] = mh[h];
          St[tid * 3 * H + h * 3 + 1] = lh[h];
          St[tid * 3 * H + h * 3 + 2] = dl;
        }
#pragma unroll
        for (int u = 0; u < U; u += 4) {
          *reinterpret_cast<float4*>(arow + 2 * U + u) = make_float4(q[u], q[u + 1], q[u + 2], q[u + 3]);
          *reinterpret_cast<float4*>(arow + 3 * U + u) = make_float4(dT[u], dT[u + 1], dT[u + 2], dT[u + 3]);
        }
      }
      __syncthreads();  // Q, dO, stats visible

      float dz[N4];  // dq | dk | dv | dr
#pragma unroll
      for (int u = 0; u < N4; ++u) dz[u] = 0.f;
      if (active) {
        const float* Asmp = A + sbase * SA;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          // ---- row pass: this thread is query i
          {
            float qh[DH], doh[DH], dq[DH];
#pragma unroll
            for (int e = 0; e < DH; ++e) { qh[e] = q[h * DH + e]; doh[e] = dT[h * DH + e]; dq[e] = 0.f; }
            const float m = St[tid * 3 * H + h * 3 + 0];
            const float linv = St[tid * 3 * H + h * 3 + 1];
            const float delta = St[tid * 3 * H + h * 3 + 2];
            for (int j = 0; j < F; ++j) {
              const float* kr = Asmp + j * SA + h * DH;
              float s = 0.f, dp = 0.f;
              float kk[DH];
#pragma unroll
              for (int e = 0; e < DH; e += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kr + e);
                const float4 v4 = *reinterpret_cast<const float4*>(kr + U + e);
                kk[e] = k4.x; kk[e + 1] = k4.y; kk[e + 2] = k4.z; kk[e + 3] = k4.w;
                s = fmaf(qh[e], k4.x, s); s = fmaf(qh[e + 1], k4.y, s);
                s = fmaf(qh[e + 2], k4.z, s); s = fmaf(qh[e + 3], k4.w, s);
                dp = fmaf(doh[e], v4.x, dp); dp = fmaf(doh[e + 1], v4.y, dp);
                dp = fmaf(doh[e + 2], v4.z, dp); dp = fmaf(doh[e + 3], v4.w, dp);
              }
              const float p = exp2f(s * scale_log2 - m) * linv;
              if (dc.rate > 0.f) dp *= drop_scale(dc, drop_row_base(it, B, smp, H, h, F, fld) + (unsigned long long)j);
              const float ds = p * (dp - delta);
#pragma unroll
              for (int e = 0; e < DH; ++e) dq[e] = fmaf(ds, kk[e], dq[e]);
            }
#pragma unroll
            for (int e = 0; e < DH; ++e) dz[h * DH + e] = dq[e] * scale;
          }
          // ---- column pass: this thread is key j
          {
            float kh[DH], vh[DH], dk[DH], dv[DH];
            const float* own = A + tid * SA + h * DH;
#pragma unroll
            for (int e = 0; e < DH; ++e) { kh[e] = own[e]; vh[e] = own[U + e]; dk[e] = 0.f; dv[e] = 0.f; }
            for (int i = 0; i < F; ++i) {
              const float* qr = Asmp + i * SA + 2 * U + h * DH;
              const float* st = St + (sbase + i) * 3 * H + h * 3;
              float s = 0.f, dp = 0.f;
              float qq[DH], dd[DH];
#pragma unroll
              for (int e = 0; e < DH; e += 4) {
                const float4 q4 = *reinterpret_cast<const float4*>(qr + e);
                const float4 d4 = *reinterpret_cast<const float4*>(qr + U + e);
                qq[e] = q4.x; qq[e + 1] = q4.y; qq[e + 2] = q4.z; qq[e + 3] = q4.w;
                dd[e] = d4.x; dd[e + 1] = d4.y; dd[e + 2] = d4.z; dd[e + 3] = d4.w;
                s = fmaf(q4.x, kh[e], s); s = fmaf(q4.y, kh[e + 1], s);
                s = fmaf(q4.z, kh[e + 2], s); s = fmaf(q4.w, kh[e + 3], s);
                dp = fmaf(d4.x, vh[e], dp); dp = fmaf(d4.y, vh[e + 1], dp);
                dp = fmaf(d4.z, vh[e + 2], dp); dp = fmaf(d4.w, vh[e + 3], dp);
              }
              const float p = exp2f(s * scale_log2 - st[0]) * st[1];
              float ks = 1.f;                                          // mask of element (query i, key = this row)
              if (dc.rate > 0.f) ks = drop_scale(dc, drop_row_base(it, B, smp, H, h, F, i) + (unsigned long long)fld);
              const float ds = p * (ks * dp - st[2]);
              const float pk = p * ks;
#pragma unroll
              for (int e = 0; e < DH; ++e) {
                dk[e] = fmaf(ds, qq[e], dk[e]);
                dv[e] = fmaf(pk, dd[e], dv[e]);
              }
            }
#pragma unroll
            for (int e = 0; e < DH; ++e) {
              dz[U + h * DH + e] = kh[e] > 0.f ? dk[e] * scale : 0.f;
              dz[2 * U + h * DH + e] = vh[e] > 0.f ? dv[e] : 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          dz[u] = q[u] > 0.f ? dz[u] : 0.f;
          dz[3 * U + u] = (use_res && ((rmask >> u) & 1u)) ? dT[u] : 0.f;
        }
      }
      __syncthreads();  // all reads of K|V|Q|dO done -> region A becomes dZ
#pragma unroll
      for (int u = 0; u < N4; u += 4)
        *reinterpret_cast<float4*>(arow + u) = make_float4(dz[u], dz[u + 1], dz[u + 2], dz[u + 3]);
      {
        float* xrow = Xs + tid * SX;
#pragma unroll
        for (int d = 0; d < D; d += 4)
          *reinterpret_cast<float4*>(xrow + d) = make_float4(xr[d], xr[d + 1], xr[d + 2], xr[d + 3]);
      }
      // dx = dZ W^T  (per row, registers)
      float dxr[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < N4; u += 4) {
          const float4 w = *reinterpret_cast<const float4*>(Ws + d * N4 + u);
          acc = fmaf(dz[u], w.x, acc); acc = fmaf(dz[u + 1], w.y, acc);
          acc = fmaf(dz[u + 2], w.z, acc); acc = fmaf(dz[u + 3], w.w, acc);
        }
        dxr[d] = acc;
      }
      __syncthreads();  // dZ, x stash, LN terms visible
      // dW[w_d][w_u0 .. +OPT) += sum_rows x[row][w_d] * dZ[row][..]  (fixed row order)
      for (int rr = 0; rr < rows_here; ++rr) {
        const float xv = Xs[rr * SX + w_d];
        const float* zr = A + rr * SA + w_u0;
#pragma unroll
        for (int k = 0; k < OPT; ++k) dWacc[k] = fmaf(xv, zr[k], dWacc[k]);
      }
#pragma unroll
      for (int k = 0; k < NCS; ++k) {
        const int c = tid + k * NT;
        if (c < 6 * U) {
          const float* base = c < N4 ? (A + c) : (Bt + (c - N4));
          const int stride = c < N4 ? SA : SB;
          float acc = csacc[k];
          for (int rr = 0; rr < rows_here; ++rr) acc += base[rr * stride];
          csacc[k] = acc;
        }
      }
      __syncthreads();  // smem free for the next iteration
      if (it > 0) {
        if constexpr (D == U) {
#pragma unroll
          for (int u = 0; u < U; ++u) g[u] = dxr[u];   // stays fp32 between iterations
        }
      } else if (active) {
        store_row<D, T>(dx + smp * dx_bs + fld * dx_ld, dxr);
      }
    }
  }
  // per-CTA partials: dW | db | dgamma | dbeta
  float* mine = part + (int64_t)blockIdx.x * (D * N4 + 6 * U);
#pragma unroll
  for (int k = 0; k < OPT; ++k) mine[tid * OPT + k] = dWacc[k];
#pragma unroll
  for (int k = 0; k < NCS; ++k) {
    const int c = tid + k * NT;
    if (c < 6 * U) mine[D * N4 + c] = csacc[k];
  }
}

// out[i] = sum over CTAs of part[cta][i]; one warp per output (ordered => deterministic)
static __global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out,
                                              int nparts, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const float s = warp_ordered_sum(part + i, nparts, n);
  if ((threadIdx.x & 31) == 0) out[i] = s;
}

// ------------------------------------------------------------ host dispatch
static inline DropCfg drop_cfg(float rate, unsigned long long seed, const unsigned long long* step) {
  DropCfg dc;
  dc.step = step;
  dc.rate = rate;
  dc.inv_keep = rate > 0.f ? 1.f / (1.f - rate) : 1.f;
  dc.seed = seed;
  return dc;
}

template <int D, int U, int NT>
static size_t fwd_smem_bytes() { return (size_t)(D * 4 * U + 4 * U + 2 * U + NT * (2 * U + 4)) * 4; }
template <int D, int U, int H, int NT>
static size_t bwd_smem_bytes() {
  using S = ISmem<D, U>;
  return (size_t)(D * 4 * U + 4 * U + 2 * U + NT * (S::SA + S::SB + S::SX) + NT * 3 * H) * 4;
}

// CTAs of `kern` that are resident on one SM at once, clamped to [1, cap]: the persistent
// grids below are sized to exactly one full wave.
static inline int resident_ctas(const void* kern, int threads, size_t smem, int cap) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess) occ = 1;
  return occ < 1 ? 1 : (occ > cap ? cap : occ);
}

inline int interacting_bwd_grid(int B, int F, int NT, int per_sm = 2) {
  const int SPT = NT / F;
  const int ntiles = (B + SPT - 1) / SPT;
  int g = sm_count() * per_sm;
  if (g > ntiles) g = ntiles;
  if (g < 1) g = 1;
  return g;
}

template <int D, int U, int H, int NT, typename T>
static int launch_fwd(const IFwdArgs& a) {
  const size_t smem = fwd_smem_bytes<D, U, NT>();
  auto kern = interacting_fwd_kernel<D, U, H, NT, T>;
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int SPT = NT / a.F;
  const int ntiles = (a.B + SPT - 1) / SPT;
  int grid = sm_count() * resident_ctas((const void*)kern, NT, smem, 8);
  if (grid > ntiles) grid = ntiles;
  kern<<<grid, NT, smem, a.st>>>((const T*)a.x, a.x_ld, a.x_bs, a.W, a.b, a.gm, a.bt, a.eps, (T*)a.y, a.y_ld, a.y_bs,
                                 (float*)a.saved, a.B, a.F, a.L, a.use_res, drop_cfg(a.drop_rate, a.drop_seed, a.drop_step));
  return check_launch("interacting_fwd");
}

template <int D, int U, int H, int NT, typename T>
static int launch_bwd(const IBwdArgs& a) {
  const size_t smem = bwd_smem_bytes<D, U, H, NT>();
  auto kern = interacting_bwd_kernel<D, U, H, NT, T>;
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = interacting_bwd_grid(a.B, a.F, NT, resident_ctas((const void*)kern, NT, smem, 2));
  const int np = D * 4 * U + 6 * U;
  if (a.ws_bytes < (size_t)grid * np * sizeof(float)) {
    set_error("interacting_bwd: workspace %zu < %zu", a.ws_bytes, (size_t)grid * np * sizeof(float));
    return RS_ERR_WORKSPACE;
  }
  kern<<<grid, NT, smem, a.st>>>((const T*)a.x, a.x_ld, a.x_bs, (const float*)a.saved, a.W, a.b, a.gm, a.bt, a.eps,
                                 (const T*)a.dy, a.dy_ld, a.dy_bs, (T*)a.dx, a.dx_ld, a.dx_bs, (float*)a.ws, a.B, a.F,
                                 a.L, a.use_res, drop_cfg(a.drop_rate, a.drop_seed, a.drop_step));
  if (int e = check_launch("interacting_bwd")) return e;
  reduce_partials_kernel<<<(np * 32 + 255) / 256, 256, 0, a.st>>>((const float*)a.ws, a.dparams, grid, np);
  return check_launch("interacting_bwd_reduce");
}

template <int D, int U, int H>
static int fwd_shape(const IFwdArgs& a) {
  if (a.F <= 128) {
    if (a.dtype == RS_F32) return launch_fwd<D, U, H, 128, float>(a);
    return launch_fwd<D, U, H, 128, __nv_bfloat16>(a);
  }
  if (a.dtype == RS_F32) return launch_fwd<D, U, H, 256, float>(a);
  return launch_fwd<D, U, H, 256, __nv_bfloat16>(a);
}
template <int D, int U, int H>
static int bwd_shape(const IBwdArgs& a) {
  if (a.F <= 128) {
    if (a.dtype == RS_F32) return launch_bwd<D, U, H, 128, float>(a);
    return launch_bwd<D, U, H, 128, __nv_bfloat16>(a);
  }
  if (a.dtype == RS_F32) return launch_bwd<D, U, H, 256, float>(a);
  return launch_bwd<D, U, H, 256, __nv_bfloat16>(a);
}

}  // namespace rs
