// din.cu — K6: fused DIN local-activation unit, forward and backward.
//   mode A  din.py:18-47            s = relu(relu([q,k,q*k] W1 + b1) W2 + b2), zero mask, sum pool
//   mode B  staytime/layer.py:16-41 s = sigmoid([q,f,q-f,q*f] W1 + b1) W2 + b2, -2^32+1 mask,
//                                   softmax over T, weighted sum of facts
//
// The concat is never materialised.  With W1 split by input block,
//   z W1 = q WQ + k WK + (q o k) WP        A: WQ,WK,WP = W1[0:H], W1[H:2H], W1[2H:3H]
//                                           B: WQ = Wa+Wc, WK = Wb-Wc, WP = Wd
// so per SAMPLE   cq = b1 + q WQ   and   Weff[c][j] = WK[c][j] + q[c] WP[c][j]
// are formed once (H*HD + HD values in shared memory), and each behaviour
// position costs one [1,H]x[H,HD] product h = cq + k_t Weff instead of the
// 3H/4H-wide Dense of the reference: 3-4x fewer FLOPs and HBM traffic of exactly
// one pass over the keys (6.5 KB/sample at T=100, H=16, fp32).
//
// Mapping: one warp per sample (grid-stride over samples), lane l owns positions
// t = l, l+32, ...  Weff rows are read as warp-broadcast float4s; two positions
// are processed per Weff read.
// Backward: per-warp shared stash of k and dh1 rows turns the per-sample
// outer-product sums G = k^T dh1 (H x HD) into a lane-parallel reduction; the
// parameter gradients accumulate in registers over all of a warp's samples in
// sample order and are reduced over warps/CTAs in a fixed order (deterministic,
// no atomics).
#include "common.cuh"

namespace rs {

constexpr int DIN_WARPS = 4;
constexpr float DIN_LOG2E = 1.4426950408889634f;

template <int H, int HD>
struct DinShape {
  static constexpr int NW = H * HD;                 // entries of one W block
  static constexpr int NP = 3 * NW + 2 * HD + 1;    // WQ | WK | WP | b1 | W2 | b2
  static constexpr int EPL = NW / 32;               // G entries per lane
  static constexpr int KS = H + 4;                  // padded row strides (conflict-free float4 rows)
  static constexpr int DS = HD + 4;
  static_assert(NW % 32 == 0 && HD % EPL == 0 && EPL % 4 == 0, "unsupported H/HD");
};

// Shared weights of the CTA: WQ | WK | WP | b1 | W2 | b2 (mode mapping applied).
template <int MODE, int H, int HD>
__device__ __forceinline__ void din_load_weights(float* Ws, const float* __restrict__ W1,
                                                 const float* __restrict__ b1,
                                                 const float* __restrict__ W2,
                                                 const float* __restrict__ b2) {
  using S = DinShape<H, HD>;
  for (int e = threadIdx.x; e < S::NW; e += blockDim.x) {
    if (MODE == RS_DIN_A) {
      Ws[e] = W1[e];
      Ws[S::NW + e] = W1[S::NW + e];
      Ws[2 * S::NW + e] = W1[2 * S::NW + e];
    } else {
      const float wa = W1[e], wb = W1[S::NW + e], wc = W1[2 * S::NW + e], wd = W1[3 * S::NW + e];
      Ws[e] = wa + wc;
      Ws[S::NW + e] = wb - wc;
      Ws[2 * S::NW + e] = wd;
    }
  }
  for (int e = threadIdx.x; e < HD; e += blockDim.x) {
    Ws[3 * S::NW + e] = b1[e];
    Ws[3 * S::NW + HD + e] = W2[e];
  }
  if (threadIdx.x == 0) Ws[3 * S::NW + 2 * HD] = b2[0];
}

// Per-sample setup by one warp: Weff (H*HD) and cq (HD) into the warp's smem.
template <int H, int HD, typename T>
__device__ __forceinline__ float din_setup_sample(const float* Ws, float* Weff, float* cq,
                                                  const T* __restrict__ q, int64_t b, int lane) {
  using S = DinShape<H, HD>;
  const float qc = to_f<T>(q[b * H + (lane % H)]);   // lane c (< H) holds q[c]
#pragma unroll
  for (int i = 0; i < S::NW / 32; ++i) {
    const int e = i * 32 + lane;
    const float qv = __shfl_sync(0xffffffffu, qc, e / HD);
    Weff[e] = fmaf(qv, Ws[2 * S::NW + e], Ws[S::NW + e]);
  }
  {
    const int j = lane % HD;
    float acc = Ws[3 * S::NW + j];
#pragma unroll
    for (int c = 0; c < H; ++c) acc = fmaf(__shfl_sync(0xffffffffu, qc, c), Ws[c * HD + j], acc);
    if (lane < HD) cq[lane] = acc;
  }
  __syncwarp();
  return qc;
}

template <int N, typename T>
__device__ __forceinline__ void din_load_row(const T* p, bool ok, float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N; c += 4) {
    const float4 t = ok ? load4<T>(p + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
  }
}
template <int N, typename T>
__device__ __forceinline__ void din_store_row(T* p, const float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N; c += 4) store4<T>(p + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
}

// h[r][:] = cq + k[r] Weff for two positions at once (one Weff read serves both).
template <int H, int HD>
__device__ __forceinline__ void din_hidden2(const float* Weff, const float* cq, const float (&k0)[H],
                                            const float (&k1)[H], float (&h0)[HD], float (&h1)[HD]) {
#pragma unroll
  for (int j = 0; j < HD; j += 4) {
    const float4 c4 = *reinterpret_cast<const float4*>(cq + j);
    h0[j] = c4.x; h0[j + 1] = c4.y; h0[j + 2] = c4.z; h0[j + 3] = c4.w;
    h1[j] = c4.x; h1[j + 1] = c4.y; h1[j + 2] = c4.z; h1[j + 3] = c4.w;
  }
#pragma unroll
  for (int c = 0; c < H; ++c) {
    // keep the Weff reads in program order: without the barrier the compiler batches
    // all H*HD/4 shared loads up front and spills
    if ((c & 1) == 0) asm volatile("" ::: "memory");
#pragma unroll
    for (int j = 0; j < HD; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(Weff + c * HD + j);
      h0[j] = fmaf(k0[c], w.x, h0[j]); h0[j + 1] = fmaf(k0[c], w.y, h0[j + 1]);
      h0[j + 2] = fmaf(k0[c], w.z, h0[j + 2]); h0[j + 3] = fmaf(k0[c], w.w, h0[j + 3]);
      h1[j] = fmaf(k1[c], w.x, h1[j]); h1[j + 1] = fmaf(k1[c], w.y, h1[j + 1]);
      h1[j + 2] = fmaf(k1[c], w.z, h1[j + 2]); h1[j + 3] = fmaf(k1[c], w.w, h1[j + 3]);
    }
  }
}

__device__ __forceinline__ float din_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }

// Raw score of one position from its hidden pre-activation (before masking).
template <int MODE, int HD>
__device__ __forceinline__ float din_score(const float (&h)[HD], const float* W2s, float b2) {
  float s = b2;
#pragma unroll
  for (int j = 0; j < HD; ++j) {
    const float a = MODE == RS_DIN_A ? fmaxf(h[j], 0.f) : din_sigmoid(h[j]);
    s = fmaf(a, W2s[j], s);
  }
  return MODE == RS_DIN_A ? fmaxf(s, 0.f) : s;
}

constexpr float DIN_PAD_SCORE = -4294967296.0f;  // float32(-2**32 + 1), staytime/layer.py:32

// ------------------------------------------------------------------ forward
// Positions are walked 64 at a time (two per lane per step, any T); mode B keeps
// a running (max, sum, weighted sum) so the softmax needs no second pass.
template <int MODE, int H, int HD, typename T>
#ifndef RS_DIN_MINB
#define RS_DIN_MINB 4
#endif
__global__ void __launch_bounds__(DIN_WARPS * 32, RS_DIN_MINB)
din_fwd_kernel(const T* __restrict__ q, const T* __restrict__ keys, const T* __restrict__ values,
               int64_t kv_ld, const int32_t* __restrict__ seq_len, const uint8_t* __restrict__ mask,
               const float* __restrict__ W1, const float* __restrict__ b1,
               const float* __restrict__ W2, const float* __restrict__ b2, T* __restrict__ out,
               int B, int Tn) {
  using S = DinShape<H, HD>;
  extern __shared__ float4 din_smem4[];
  float* Ws = reinterpret_cast<float*>(din_smem4);
  float* warp_base = Ws + ((S::NP + 3) & ~3);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* Weff = warp_base + wid * (S::NW + HD);
  float* cq = Weff + S::NW;
  din_load_weights<MODE, H, HD>(Ws, W1, b1, W2, b2);
  __syncthreads();
  const float* W2s = Ws + 3 * S::NW + HD;
  const float b2v = Ws[3 * S::NW + 2 * HD];

  for (int64_t b = (int64_t)blockIdx.x * DIN_WARPS + wid; b < B; b += (int64_t)gridDim.x * DIN_WARPS) {
    __syncwarp();
    din_setup_sample<H, HD, T>(Ws, Weff, cq, q, b, lane);
    const int slen = (MODE == RS_DIN_A) ? seq_len[b] : 0;
    float acc[H];
#pragma unroll
    for (int c = 0; c < H; ++c) acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
#pragma unroll 1
    for (int tb = 0; tb < Tn; tb += 64) {
      const int t0 = tb + lane, t1 = tb + 32 + lane;
      const bool in0 = t0 < Tn, in1 = t1 < Tn;
      float k0[H], k1[H];
      din_load_row<H, T>(keys + (b * Tn + t0) * kv_ld, in0, k0);
      din_load_row<H, T>(keys + (b * Tn + t1) * kv_ld, in1, k1);
      float h0[HD], h1[HD];
      din_hidden2<H, HD>(Weff, cq, k0, k1, h0, h1);
      float s0 = din_score<MODE, HD>(h0, W2s, b2v);
      float s1 = din_score<MODE, HD>(h1, W2s, b2v);
      if (MODE == RS_DIN_A) {
        s0 = (in0 && t0 < slen) ? s0 : 0.f;
        s1 = (in1 && t1 < slen) ? s1 : 0.f;
        din_load_row<H, T>(values + (b * Tn + t0) * kv_ld, in0, k0);
        din_load_row<H, T>(values + (b * Tn + t1) * kv_ld, in1, k1);
#pragma unroll
        for (int c = 0; c < H; ++c) acc[c] = fmaf(s0, k0[c], fmaf(s1, k1[c], acc[c]));
      } else {
        if (mask) {
          if (in0 && !mask[b * Tn + t0]) s0 = DIN_PAD_SCORE;
          if (in1 && !mask[b * Tn + t1]) s1 = DIN_PAD_SCORE;
        }
        s0 = in0 ? s0 : -INFINITY;
        s1 = in1 ? s1 : -INFINITY;
        const float mn = fmaxf(m, warp_max(fmaxf(s0, s1)));   // finite: t = tb is in range
        const float alpha = exp2f((m - mn) * DIN_LOG2E);
        const float p0 = exp2f((s0 - mn) * DIN_LOG2E), p1 = exp2f((s1 - mn) * DIN_LOG2E);
        l = fmaf(l, alpha, p0 + p1);
#pragma unroll
        for (int c = 0; c < H; ++c) acc[c] = fmaf(acc[c], alpha, fmaf(p0, k0[c], p1 * k1[c]));
        m = mn;
      }
    }
    const float linv = (MODE == RS_DIN_B) ? 1.f / warp_sum(l) : 1.f;
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < H; ++c) {
      const float s = warp_sum(acc[c]);
      if (lane == c) mine = s;
    }
    if (lane < H) out[b * H + lane] = from_f<T>(mine * linv);
  }
}

// ----------------------------------------------------------------- backward
// Per-warp shared stash (Tpad = T rounded up to 64 rows): h1 -> dh1 rows and three
// per-position scalars (score -> softmax weight, dS, live flag).  The k rows are not
// stashed for the whole sequence: phase 2 re-reads the row of its position (L2-resident
// since phase 1) into a 32-row chunk and the outer products of phase 3 are accumulated
// chunk by chunk, in the same position order.  15.5 KB per warp for T = 100: three
// CTAs per SM instead of two.
template <int MODE, int H, int HD, typename T>
#ifndef RS_DIN_BWD_MINB
#define RS_DIN_BWD_MINB 3
#endif
__global__ void __launch_bounds__(DIN_WARPS * 32, RS_DIN_BWD_MINB)
din_bwd_kernel(const T* __restrict__ q, const T* __restrict__ keys, const T* __restrict__ values,
               int64_t kv_ld, const int32_t* __restrict__ seq_len, const uint8_t* __restrict__ mask,
               const float* __restrict__ W1, const float* __restrict__ b1,
               const float* __restrict__ W2, const float* __restrict__ b2,
               const T* __restrict__ dout, T* __restrict__ dq, T* __restrict__ dkeys,
               T* __restrict__ dvalues, int64_t dkv_ld, float* __restrict__ part, int B, int Tn,
               int Tpad) {
  using S = DinShape<H, HD>;
  constexpr int EPL = S::EPL;
  extern __shared__ float4 din_smem4[];
  float* Ws = reinterpret_cast<float*>(din_smem4);
  float* warp_base = Ws + ((S::NP + 3) & ~3);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per_warp = S::NW + HD + HD + H + 32 * S::KS + Tpad * (S::DS + 3);
  float* Weff = warp_base + wid * per_warp;
  float* cq = Weff + S::NW;
  float* gs = cq + HD;            // g[j] = sum_t dh1_t[j]
  float* dos = gs + HD;           // dout of the sample
  float* Kc = dos + H;            // [32][KS]    k rows of the current chunk (phase 2/3)
  float* Ds = Kc + 32 * S::KS;    // [Tpad][DS]  h1, then dh1
  float* Sc = Ds + Tpad * S::DS;  // [Tpad] score, then softmax weight (B)
  float* Dp = Sc + Tpad;          // [Tpad] A: dout.v   B: dout.f, then dS
  float* Lv = Dp + Tpad;          // [Tpad] 1 if gradient reaches the score
  din_load_weights<MODE, H, HD>(Ws, W1, b1, W2, b2);
  __syncthreads();
  const float* W2s = Ws + 3 * S::NW + HD;
  const float b2v = Ws[3 * S::NW + 2 * HD];

  // lane-owned slice of the H x HD outer products
  const int gc = (lane * EPL) / HD;
  const int gj0 = (lane * EPL) % HD;
  float aWQ[EPL], aWK[EPL], aWP[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) { aWQ[e] = 0.f; aWK[e] = 0.f; aWP[e] = 0.f; }
  float ab1 = 0.f;            // lane j < HD
  float aW2[HD];              // per-lane partials (own positions), reduced at the end
#pragma unroll
  for (int j = 0; j < HD; ++j) aW2[j] = 0.f;
  float ab2 = 0.f;

  for (int64_t b = (int64_t)blockIdx.x * DIN_WARPS + wid; b < B; b += (int64_t)gridDim.x * DIN_WARPS) {
    __syncwarp();
    const float qc = din_setup_sample<H, HD, T>(Ws, Weff, cq, q, b, lane);
    if (lane < H) dos[lane] = to_f<T>(dout[b * H + lane]);
    __syncwarp();
    float dov[H];
#pragma unroll
    for (int c = 0; c < H; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(dos + c);
      dov[c] = t.x; dov[c + 1] = t.y; dov[c + 2] = t.z; dov[c + 3] = t.w;
    }
    const int slen = (MODE == RS_DIN_A) ? seq_len[b] : 0;

    // ---- phase 1: forward recompute; stash k and h1 rows and per-position scalars
    float mx = -INFINITY;
#pragma unroll 1
    for (int tb = 0; tb < Tn; tb += 64) {
      const int t0 = tb + lane, t1 = tb + 32 + lane;
      const bool in0 = t0 < Tn, in1 = t1 < Tn;
      float k0[H], k1[H];
      din_load_row<H, T>(keys + (b * Tn + t0) * kv_ld, in0, k0);
      din_load_row<H, T>(keys + (b * Tn + t1) * kv_ld, in1, k1);
      float h0[HD], h1[HD];
      din_hidden2<H, HD>(Weff, cq, k0, k1, h0, h1);
#pragma unroll
      for (int j = 0; j < HD; j += 4) {
        *reinterpret_cast<float4*>(Ds + t0 * S::DS + j) = make_float4(h0[j], h0[j + 1], h0[j + 2], h0[j + 3]);
        *reinterpret_cast<float4*>(Ds + t1 * S::DS + j) = make_float4(h1[j], h1[j + 1], h1[j + 2], h1[j + 3]);
      }
      float s0 = din_score<MODE, HD>(h0, W2s, b2v);
      float s1 = din_score<MODE, HD>(h1, W2s, b2v);
      float d0 = 0.f, d1 = 0.f;
      bool lv0, lv1;
      if (MODE == RS_DIN_A) {
        const bool ok0 = in0 && t0 < slen, ok1 = in1 && t1 < slen;
        s0 = ok0 ? s0 : 0.f;
        s1 = ok1 ? s1 : 0.f;
        din_load_row<H, T>(values + (b * Tn + t0) * kv_ld, in0, k0);
        din_load_row<H, T>(values + (b * Tn + t1) * kv_ld, in1, k1);
#pragma unroll
        for (int c = 0; c < H; ++c) { d0 = fmaf(dov[c], k0[c], d0); d1 = fmaf(dov[c], k1[c], d1); }
#pragma unroll
        for (int c = 0; c < H; ++c) { k0[c] = s0 * dov[c]; k1[c] = s1 * dov[c]; }   // dvalues_t = s_t dout
        if (in0) din_store_row<H, T>(dvalues + (b * Tn + t0) * dkv_ld, k0);
        if (in1) din_store_row<H, T>(dvalues + (b * Tn + t1) * dkv_ld, k1);
        lv0 = ok0 && s0 > 0.f;        // relu output > 0 <=> pre-activation > 0
        lv1 = ok1 && s1 > 0.f;
      } else {
        lv0 = in0; lv1 = in1;
        if (mask) {
          lv0 = in0 && mask[b * Tn + t0];
          lv1 = in1 && mask[b * Tn + t1];
        }
        if (in0 && !lv0) s0 = DIN_PAD_SCORE;
        if (in1 && !lv1) s1 = DIN_PAD_SCORE;
        s0 = in0 ? s0 : -INFINITY;
        s1 = in1 ? s1 : -INFINITY;
#pragma unroll
        for (int c = 0; c < H; ++c) { d0 = fmaf(dov[c], k0[c], d0); d1 = fmaf(dov[c], k1[c], d1); }
        mx = fmaxf(mx, fmaxf(s0, s1));
      }
      Sc[t0] = s0; Sc[t1] = s1;
      Dp[t0] = d0; Dp[t1] = d1;
      Lv[t0] = lv0 ? 1.f : 0.f; Lv[t1] = lv1 ? 1.f : 0.f;
    }
    // ---- softmax statistics (mode B): p_t and dS_t = p_t (dp_t - sum_u p_u dp_u)
    if (MODE == RS_DIN_B) {
      mx = warp_max(mx);
      float l = 0.f, dsum = 0.f;
#pragma unroll 1
      for (int t = lane; t < Tpad; t += 32) {
        const float p = exp2f((Sc[t] - mx) * DIN_LOG2E);
        Sc[t] = p;
        l += p;
        dsum = fmaf(p, Dp[t], dsum);
      }
      const float linv = 1.f / warp_sum(l);
      dsum = warp_sum(dsum) * linv;
#pragma unroll 1
      for (int t = lane; t < Tpad; t += 32) {
        const float p = Sc[t] * linv;
        Sc[t] = p;
        Dp[t] = p * (Dp[t] - dsum);
      }
    }
    // ---- phase 2 + 3: per position dh1 and dk (rows of dkeys / dfacts); per 32-position
    //      chunk G[c][j] += sum_t k_t[c] dh1_t[j] (lane slice), then g[j] and dq
    float G[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) G[e] = 0.f;
#pragma unroll 1
    for (int tb = 0; tb < Tpad; tb += 32) {
      const int t = tb + lane;
      const bool in = t < Tn;
      float kr[H];
      din_load_row<H, T>(keys + (b * Tn + t) * kv_ld, in, kr);
      const float ds = Lv[t] != 0.f ? Dp[t] : 0.f;
      const float pw = Sc[t];
      float dh[HD];
#pragma unroll
      for (int j = 0; j < HD; j += 4) {
        const float4 h4 = *reinterpret_cast<const float4*>(Ds + t * S::DS + j);
        const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float a, da;
          if (MODE == RS_DIN_A) { a = fmaxf(hv[u], 0.f); da = hv[u] > 0.f ? 1.f : 0.f; }
          else { a = din_sigmoid(hv[u]); da = a * (1.f - a); }
          aW2[j + u] = fmaf(a, ds, aW2[j + u]);
          dh[j + u] = ds * W2s[j + u] * da;
        }
        *reinterpret_cast<float4*>(Ds + t * S::DS + j) = make_float4(dh[j], dh[j + 1], dh[j + 2], dh[j + 3]);
      }
      ab2 += ds;
      float dk[H];
#pragma unroll
      for (int c = 0; c < H; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < HD; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(Weff + c * HD + j);
          acc = fmaf(dh[j], w.x, acc); acc = fmaf(dh[j + 1], w.y, acc);
          acc = fmaf(dh[j + 2], w.z, acc); acc = fmaf(dh[j + 3], w.w, acc);
        }
        dk[c] = (MODE == RS_DIN_B) ? fmaf(pw, dov[c], acc) : acc;
      }
      if (in) din_store_row<H, T>(dkeys + (b * Tn + t) * dkv_ld, dk);
#pragma unroll
      for (int c = 0; c < H; c += 4)
        *reinterpret_cast<float4*>(Kc + lane * S::KS + c) = make_float4(kr[c], kr[c + 1], kr[c + 2], kr[c + 3]);
      __syncwarp();
      const int nt = min(32, Tn - tb);
#pragma unroll 4
      for (int u = 0; u < nt; ++u) {
        const float kc = Kc[u * S::KS + gc];
#pragma unroll
        for (int e = 0; e < EPL; e += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(Ds + (tb + u) * S::DS + gj0 + e);
          G[e] = fmaf(kc, d4.x, G[e]); G[e + 1] = fmaf(kc, d4.y, G[e + 1]);
          G[e + 2] = fmaf(kc, d4.z, G[e + 2]); G[e + 3] = fmaf(kc, d4.w, G[e + 3]);
        }
      }
      __syncwarp();
    }
    {
      const int j = lane % HD, half = lane / HD;      // HD == 16: two lanes per column
      float g = 0.f;
      for (int t = half; t < Tn; t += 32 / HD) g += Ds[t * S::DS + j];
#pragma unroll
      for (int o = HD; o < 32; o <<= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if (lane < HD) { gs[lane] = g; ab1 += g; }
    }
    __syncwarp();
    const float qv = __shfl_sync(0xffffffffu, qc, gc);
    float dqp = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const float gj = gs[gj0 + e];
      aWQ[e] = fmaf(qv, gj, aWQ[e]);
      aWK[e] += G[e];
      aWP[e] = fmaf(qv, G[e], aWP[e]);
      dqp = fmaf(gj, Ws[gc * HD + gj0 + e], dqp);
      dqp = fmaf(G[e], Ws[2 * S::NW + gc * HD + gj0 + e], dqp);
    }
#pragma unroll
    for (int o = 1; o < HD / EPL; o <<= 1) dqp += __shfl_xor_sync(0xffffffffu, dqp, o);
    if ((lane % (HD / EPL)) == 0) dq[b * H + gc] = from_f<T>(dqp);
  }

  // ---- reduce parameter-gradient partials: lanes -> warp -> CTA -> workspace
  __syncthreads();
  float* red = warp_base;      // reuse the per-warp regions (per_warp >= NP floats each)
  float* mine = red + wid * per_warp;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    mine[lane * EPL + e] = aWQ[e];
    mine[S::NW + lane * EPL + e] = aWK[e];
    mine[2 * S::NW + lane * EPL + e] = aWP[e];
  }
  if (lane < HD) mine[3 * S::NW + lane] = ab1;
#pragma unroll
  for (int j = 0; j < HD; ++j) {
    const float s = warp_sum(aW2[j]);
    if (lane == 0) mine[3 * S::NW + HD + j] = s;
  }
  {
    const float s = warp_sum(ab2);
    if (lane == 0) mine[3 * S::NW + 2 * HD] = s;
  }
  __syncthreads();
  float* dst = part + (int64_t)blockIdx.x * S::NP;
  for (int i = threadIdx.x; i < S::NP; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DIN_WARPS; ++w) s += red[w * per_warp + i];
    dst[i] = s;
  }
}

// Sum CTA partials in CTA order and map (WQ, WK, WP) back to the W1 layout.
template <int MODE, int H, int HD>
__global__ void din_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dparams) {
  using S = DinShape<H, HD>;
  constexpr int NIN = (MODE == RS_DIN_A ? 3 : 4) * H;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per output
  const int total = NIN * HD + 2 * HD + 1;
  if (i >= total) return;
  auto sum_of = [&](int idx) { return warp_ordered_sum(part + idx, nparts, S::NP); };
  float v;
  if (i < NIN * HD) {
    const int blk = i / S::NW, e = i % S::NW;
    if (MODE == RS_DIN_A) v = sum_of(blk * S::NW + e);
    else if (blk == 0) v = sum_of(e);
    else if (blk == 1) v = sum_of(S::NW + e);
    else if (blk == 2) v = sum_of(e) - sum_of(S::NW + e);
    else v = sum_of(2 * S::NW + e);
  } else {
    v = sum_of(3 * S::NW + (i - NIN * HD));
  }
  if ((threadIdx.x & 31) == 0) dparams[i] = v;
}

template <int H, int HD>
static size_t din_fwd_smem() {
  using S = DinShape<H, HD>;
  return (size_t)(((S::NP + 3) & ~3) + DIN_WARPS * (S::NW + HD)) * sizeof(float);
}
template <int H, int HD>
static size_t din_bwd_smem(int Tpad) {
  using S = DinShape<H, HD>;
  return (size_t)(((S::NP + 3) & ~3) +
                  DIN_WARPS * (S::NW + HD + HD + H + 32 * S::KS + Tpad * (S::DS + 3))) * sizeof(float);
}

// Persistent grid: one warp per sample, at most `per_sm` CTAs per SM.  per_sm is the
// number of CTAs that are resident at once (a larger grid runs a second, partly empty wave).
constexpr int DIN_MAX_CTAS_PER_SM = 4;
static int din_grid(int B, int per_sm = DIN_MAX_CTAS_PER_SM) {
  int64_t g = cdiv(B, DIN_WARPS);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int din_resident(const void* kern, size_t smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DIN_WARPS * 32, smem) != cudaSuccess) occ = 1;
  return occ < 1 ? 1 : (occ > DIN_MAX_CTAS_PER_SM ? DIN_MAX_CTAS_PER_SM : occ);
}

struct DinArgs {
  const void *q, *keys, *values; int64_t kv_ld; const int32_t* seq_len; const uint8_t* mask;
  const float *W1, *b1, *W2, *b2; void* out; const void* dout; void *dq, *dkeys, *dvalues;
  int64_t dkv_ld; float* dparams; int B, T; void* ws; size_t ws_bytes; cudaStream_t st;
};

template <int MODE, int H, int HD, typename T>
static int din_launch_fwd(const DinArgs& a) {
  auto kern = din_fwd_kernel<MODE, H, HD, T>;
  const size_t smem = din_fwd_smem<H, HD>();
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<din_grid(a.B, din_resident((const void*)kern, smem)), DIN_WARPS * 32, smem, a.st>>>(
      (const T*)a.q, (const T*)a.keys, (const T*)a.values, a.kv_ld, a.seq_len, a.mask, a.W1, a.b1, a.W2,
      a.b2, (T*)a.out, a.B, a.T);
  return check_launch("din_fwd");
}

template <int MODE, int H, int HD, typename T>
static int din_launch_bwd(const DinArgs& a) {
  using S = DinShape<H, HD>;
  auto kern = din_bwd_kernel<MODE, H, HD, T>;
  const int Tpad = (a.T + 63) / 64 * 64;
  const size_t smem = din_bwd_smem<H, HD>(Tpad);
  if (smem > 200 * 1024) {
    set_error("din_bwd: T=%d needs %zu B of shared memory per CTA (> 200 KB)", a.T, smem);
    return RS_ERR_UNSUPPORTED;
  }
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = din_grid(a.B, din_resident((const void*)kern, smem));
  if (a.ws_bytes < (size_t)grid * S::NP * sizeof(float)) {
    set_error("din_bwd: workspace %zu < %zu", a.ws_bytes, (size_t)grid * S::NP * sizeof(float));
    return RS_ERR_WORKSPACE;
  }
  kern<<<grid, DIN_WARPS * 32, smem, a.st>>>(
      (const T*)a.q, (const T*)a.keys, (const T*)a.values, a.kv_ld, a.seq_len, a.mask, a.W1, a.b1, a.W2,
      a.b2, (const T*)a.dout, (T*)a.dq, (T*)a.dkeys, (T*)a.dvalues, a.dkv_ld, (float*)a.ws, a.B, a.T, Tpad);
  if (int e = check_launch("din_bwd")) return e;
  constexpr int NIN = (MODE == RS_DIN_A ? 3 : 4) * H;
  const int total = NIN * HD + 2 * HD + 1;
  din_reduce_kernel<MODE, H, HD><<<(total * 32 + 255) / 256, 256, 0, a.st>>>((const float*)a.ws, grid, a.dparams);
  return check_launch("din_bwd_reduce");
}

template <bool BWD, int MODE, int H, int HD>
static int din_dispatch_t(const DinArgs& a, int dtype) {
  if (dtype == RS_F32)
    return BWD ? din_launch_bwd<MODE, H, HD, float>(a) : din_launch_fwd<MODE, H, HD, float>(a);
  return BWD ? din_launch_bwd<MODE, H, HD, __nv_bfloat16>(a)
             : din_launch_fwd<MODE, H, HD, __nv_bfloat16>(a);
}

template <bool BWD>
static int din_dispatch(int mode, int H, int Hd, int dtype, const DinArgs& a) {
  RS_REQUIRE(mode == RS_DIN_A || mode == RS_DIN_B, "din: mode=%d", mode);
  RS_REQUIRE(a.B > 0 && a.T > 0, "din: B=%d T=%d", a.B, a.T);
  RS_REQUIRE(dtype == RS_F32 || dtype == RS_BF16, "din: bad dtype");
  RS_REQUIRE(a.kv_ld % 4 == 0 && a.kv_ld >= H, "din: kv_ld=%lld must be a multiple of 4 and >= H", (long long)a.kv_ld);
  RS_REQUIRE(mode != RS_DIN_A || a.seq_len != nullptr, "din: mode A needs seq_len (din.py:24)");
  if (Hd == 16 && H == 16) {
    if (mode == RS_DIN_A) return din_dispatch_t<BWD, RS_DIN_A, 16, 16>(a, dtype);
    return din_dispatch_t<BWD, RS_DIN_B, 16, 16>(a, dtype);
  }
  if (Hd == 16 && H == 8) {
    if (mode == RS_DIN_A) return din_dispatch_t<BWD, RS_DIN_A, 8, 16>(a, dtype);
    return din_dispatch_t<BWD, RS_DIN_B, 8, 16>(a, dtype);
  }
  set_error("din: (H=%d, Hd=%d) not built (H in {8,16}, Hd = 16)", H, Hd);
  return RS_ERR_UNSUPPORTED;
}

}  // namespace rs

using namespace rs;

extern "C" {

int rs_din_fwd(int mode, const void* q, const void* keys, const void* values, int64_t kv_ld,
               int dtype, const int32_t* seq_len, const uint8_t* mask, const float* W1,
               const float* b1, const float* W2, const float* b2, void* out, int B, int T, int H,
               int Hd, void* stream) {
  DinArgs a{};
  a.q = q; a.keys = keys; a.values = values ? values : keys; a.kv_ld = kv_ld; a.seq_len = seq_len;
  a.mask = mask; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.out = out; a.B = B; a.T = T;
  a.st = as_stream(stream);
  return din_dispatch<false>(mode, H, Hd, dtype, a);
}

size_t rs_din_workspace_bytes(int mode, int B, int T, int H, int Hd) {
  (void)mode; (void)T;
  return (size_t)din_grid(B > 0 ? B : 1) * (size_t)(3 * H * Hd + 2 * Hd + 1) * sizeof(float);
}

int rs_din_bwd(int mode, const void* q, const void* keys, const void* values, int64_t kv_ld,
               int dtype, const int32_t* seq_len, const uint8_t* mask, const float* W1,
               const float* b1, const float* W2, const float* b2, const void* dout, void* dq,
               void* dkeys, void* dvalues, int64_t dkv_ld, float* dparams, int B, int T, int H,
               int Hd, void* ws, size_t ws_bytes, void* stream) {
  RS_REQUIRE(dkv_ld % 4 == 0 && dkv_ld >= H, "din_bwd: dkv_ld=%lld", (long long)dkv_ld);
  RS_REQUIRE(mode != RS_DIN_A || dvalues != nullptr, "din_bwd: mode A needs dvalues");
  DinArgs a{};
  a.q = q; a.keys = keys; a.values = values ? values : keys; a.kv_ld = kv_ld; a.seq_len = seq_len;
  a.mask = mask; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.dout = dout; a.dq = dq;
  a.dkeys = dkeys; a.dvalues = dvalues; a.dkv_ld = dkv_ld; a.dparams = dparams; a.B = B; a.T = T;
  a.ws = ws; a.ws_bytes = ws_bytes; a.st = as_stream(stream);
  return din_dispatch<true>(mode, H, Hd, dtype, a);
}

}  // extern "C"
