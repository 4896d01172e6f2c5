// interacting_tc.cu — K4 on the 5th-gen tensor cores (bf16 mode): fused InteractingLayer
// forward (InteractingLayer.py:37-61) with every contraction issued as tcgen05.mma.
//
// A CTA of 128 threads owns a tile of SPT whole samples; sample s occupies the tile rows
// [s*FP, s*FP + F) with FP = F rounded up to 8 (F = 39 -> 3 samples at rows 0/40/80).  Thread t
// IS tile row t and TMEM lane t, so every accumulator row comes back to the thread that owns
// the (sample, field) row and the row-wise softmax / residual / ReLU / LayerNorm need no
// cross-thread traffic at all.  Per iteration of the layer_num loop (weights shared):
//
//   1. Z[128,4U]  = X[128,D] Wqkvr[D,4U]        kind::tf32, 3xTF32 split (fp32-grade), six MMAs of K = 8
//      thread: tcgen05.ld its Z row, +bias, ReLU -> q, k, v, r
//   2. S_h[128,128] = Q_h K_h^T  per head        kind::tf32, K = U/H = 8, one MMA per head
//      (block diagonal in effect: a thread reads only the FP columns of its own sample)
//      thread: masked softmax over its F keys (scale folded into exp2), unnormalised p -> bf16
//   3. O_h[128,U]  = P_h[128,128] V[128,U]       kind::f16, 8 MMAs of K = 16 per head; P rows are
//      zero outside the sample's own key window, so other samples never contribute
//      thread: O/l + r -> ReLU -> LayerNorm -> y (next iteration's X)
//
// Shared-memory operand tiles are written by the owning threads directly in the canonical UMMA
// layouts (no-swizzle 8x16B core matrices for X/W/Q/K/V, SWIZZLE_128B for P); the P buffers are
// zero-filled once per CTA because a row's key window never moves.  TMEM: 256 columns
// (S_0 | S_1, with Z and O_h aliased onto dead S columns) => 2 CTAs per SM.
#include "tc_common.cuh"
#include "interacting_args.cuh"

namespace rs {


template <int D, int U, int H, int NCHF, typename T>
__global__ void __launch_bounds__(128, 2)
interacting_tc_fwd_kernel(const T* __restrict__ x, int64_t x_ld, int64_t x_bs, const float* __restrict__ W,
                          const float* __restrict__ bias, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, T* __restrict__ y, int64_t y_ld,
                          int64_t y_bs, float* __restrict__ saved, int B, int F, int L, int use_res) {
  static_assert(D == 16 && U == 16 && H == 2, "tensor-core path is built for D = U = 16, 2 heads");
  constexpr int DH = U / H;          // 8 -> one tf32 K step
  constexpr int N4 = 4 * U;          // 64
  constexpr int FP = NCHF * 8;       // padded fields per sample (key-window width)
  constexpr int SPT = 128 / FP;      // samples per tile
  // ---- shared memory carve-up (1024-byte aligned base for the swizzled P tiles)
  constexpr int P_BYTES = 128 * 256;             // [128 rows][128 keys] bf16
  constexpr int OFF_P = 0;                       // H buffers
  constexpr int OFF_X = OFF_P + H * P_BYTES;     // [128][x_hi(16) | x_lo(16)] tf32, 8 chunks/row  16 KB
  constexpr int OFF_V = OFF_X + 16384;           // [128 keys][16] bf16            4 KB
  constexpr int OFF_Q = OFF_V + 4096;            // H x [128][8] tf32, 2 chunks/row
  constexpr int OFF_K = OFF_Q + H * 4096;
  constexpr int OFF_W = OFF_K + H * 4096;        // W_hi | W_lo, each [64 n][16 k] tf32   8 KB
  constexpr int OFF_F = OFF_W + 8192;            // bias[64] gamma[16] beta[16] fp32
  constexpr int OFF_BAR = OFF_F + (N4 + 2 * U) * 4;
  extern __shared__ uint8_t itc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(itc_smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bs = reinterpret_cast<float*>(smem + OFF_F);
  float* gs = bs + N4;
  float* be = gs + U;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  // ---- one-time setup: barrier, TMEM, weights (bf16 W^T in UMMA layout), zeroed P buffers
  if (tid == 0) {
    mbar_init(bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < H * P_BYTES / 16; i += 128) reinterpret_cast<uint4*>(smem + OFF_P)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < N4 + 2 * U; i += 128) bs[i] = i < N4 ? bias[i] : (i < N4 + U ? gamma[i - N4] : beta[i - N4 - U]);
  stage_w_3xtf32(smem + OFF_W, W, tid, 128);     // B operands of the projection
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;        // this warp's TMEM lanes
  constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_Z = 128, TM_O0 = 0, TM_O1 = 128;

  // descriptors that never change
  const uint32_t sbase = smem_u32(smem);
  constexpr uint32_t ID_Z = make_idesc(2, 128, N4, 0, 0);               // tf32, N = 64
  constexpr uint32_t ID_S = make_idesc(2, 128, 128, 0, 0);              // tf32, N = 128
  constexpr uint32_t ID_O = make_idesc(1, 128, U, 0, 1);                // bf16, B = V MN-major, N = 16

  const int s_loc = tid / FP, f_loc = tid - s_loc * FP;
  const int issuer = tid == 0 ? 0 : (tid == 32 ? 1 : -1);   // one MMA issuer per head; both commit every phase
  const int ntiles = (B + SPT - 1) / SPT;
  const float scale_log2 = ITC_LOG2E / sqrtf((float)DH);
  uint32_t phase = 0;
  // samples whose rows intersect this warp's 32 lanes (warp-uniform loop bounds)
  const int ws_lo = (warp * 32) / FP, ws_hi = min((warp * 32 + 31) / FP, SPT - 1);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t smp = (int64_t)tile * SPT + s_loc;
    const bool active = s_loc < SPT && f_loc < F && smp < B;
    float xr[D];
    if (active) {
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        const float4 t4 = load4<T>(x + smp * x_bs + (int64_t)f_loc * x_ld + c);
        xr[c] = t4.x; xr[c + 1] = t4.y; xr[c + 2] = t4.z; xr[c + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) xr[c] = 0.f;
    }
    float yv[U];
    for (int it = 0; it < L; ++it) {
      // ---- 1. X tile (3xTF32 split of the fp32 row) -> Z = X W, fp32-grade
#pragma unroll
      for (int c = 0; c < 4; ++c)
        stage_x4_3xtf32(smem + OFF_X, tid, c, xr[c * 4], xr[c * 4 + 1], xr[c * 4 + 2], xr[c * 4 + 3]);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (issuer >= 0) {
        tc_fence_after();
        if (issuer == 0) issue_proj_3xtf32(tmem + TM_Z, sbase + OFF_X, sbase + OFF_W, ID_Z);
        tc_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1u;
      tc_fence_after();
      float q[U], r[U];
      {
        uint32_t z[32];
        tc_ld_32x32(tmem + lane_base + TM_Z, z);                       // q | k pre-activations
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = fmaxf(__uint_as_float(z[u]) + bs[u], 0.f);
        float kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) kk[u] = fmaxf(__uint_as_float(z[U + u]) + bs[U + u], 0.f);
#pragma unroll
        for (int h = 0; h < H; ++h) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {                                // tf32 = fp32 bits, 4 per chunk
            *reinterpret_cast<float4*>(smem + OFF_Q + h * 4096 + nosw_off<2>(tid, c)) =
                make_float4(q[h * DH + c * 4], q[h * DH + c * 4 + 1], q[h * DH + c * 4 + 2], q[h * DH + c * 4 + 3]);
            *reinterpret_cast<float4*>(smem + OFF_K + h * 4096 + nosw_off<2>(tid, c)) =
                make_float4(kk[h * DH + c * 4], kk[h * DH + c * 4 + 1], kk[h * DH + c * 4 + 2], kk[h * DH + c * 4 + 3]);
          }
        }
      }
      uint32_t zvr[32];                                                // v | r pre-activations: read now (S_1 aliases
      tc_ld_32x32(tmem + lane_base + TM_Z + 32, zvr);                  // these columns), used under the S MMA
      // ---- 2. S_h = Q_h K_h^T (both heads), Z columns are dead now
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (issuer >= 0) {
        tc_fence_after();
        const int h = issuer;
        tc_mma_tf32(tmem + (h == 0 ? TM_S0 : TM_S1), make_nosw_desc(sbase + OFF_Q + h * 4096, 128, 256),
                    make_nosw_desc(sbase + OFF_K + h * 4096, 128, 256), ID_S, 0u);
        tc_commit(bar);
      }
      {
        float vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          vv[u] = fmaxf(__uint_as_float(zvr[u]) + bs[2 * U + u], 0.f);
          r[u] = fmaxf(__uint_as_float(zvr[U + u]) + bs[3 * U + u], 0.f);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {                                  // V: MN-major B, row = key (read by P.V, phase 3)
          uint4 v;
          v.x = pack_bf16x2(vv[c * 8 + 0], vv[c * 8 + 1]); v.y = pack_bf16x2(vv[c * 8 + 2], vv[c * 8 + 3]);
          v.z = pack_bf16x2(vv[c * 8 + 4], vv[c * 8 + 5]); v.w = pack_bf16x2(vv[c * 8 + 6], vv[c * 8 + 7]);
          *reinterpret_cast<uint4*>(smem + OFF_V + nosw_off<2>(tid, c)) = v;
        }
      }
      mbar_wait(bar, phase); phase ^= 1u;
      tc_fence_after();
      float linv[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float p[FP];
#pragma unroll
        for (int j = 0; j < FP; ++j) p[j] = -INFINITY;
        for (int s = ws_lo; s <= ws_hi; ++s) {                          // warp-uniform; 1 or 2 trips
          const uint32_t col = (h == 0 ? TM_S0 : TM_S1) + (uint32_t)(s * FP);
          const bool mine = s == s_loc;
#pragma unroll
          for (int c0 = 0; c0 < FP; c0 += 8) {
            uint32_t t8[8];
            tc_ld_32x8(tmem + lane_base + col + c0, t8);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (mine && c0 + j < F) p[c0 + j] = __uint_as_float(t8[j]);
          }
        }
        float m4[4] = {p[0], p[1], p[2], p[3]};
#pragma unroll
        for (int j = 4; j < FP; ++j) m4[j & 3] = fmaxf(m4[j & 3], p[j]);
        float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (!active) m = 0.f;
        const float mb = m * scale_log2;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < FP; j += 2) {
          // -inf -> 0 for padded keys; round to the bf16 value the MMA will see, so that the
          // normaliser l is the sum of exactly the weights used (a true convex combination).
          // (same expression as the backward's recomputation: interacting_tc_bwd.cu)
          p[j] = ex2_approx(fmaf(p[j], scale_log2, -mb));
          p[j + 1] = ex2_approx(fmaf(p[j + 1], scale_log2, -mb));
          bf16_round2(p[j], p[j + 1]);
          l4[j & 3] += p[j];
          l4[(j + 1) & 3] += p[j + 1];
        }
        const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        linv[h] = active ? 1.f / l : 0.f;
        // P row: chunks [s_loc*NCHF, +NCHF) of the 16 chunks, SWIZZLE_128B (chunk ^= row & 7)
        if (s_loc < SPT) {
#pragma unroll
          for (int c = 0; c < NCHF; ++c) {
            const int ch = s_loc * NCHF + c;                           // 0..15
            uint4 v;
            v.x = pack_bf16x2(p[c * 8 + 0], p[c * 8 + 1]); v.y = pack_bf16x2(p[c * 8 + 2], p[c * 8 + 3]);
            v.z = pack_bf16x2(p[c * 8 + 4], p[c * 8 + 5]); v.w = pack_bf16x2(p[c * 8 + 6], p[c * 8 + 7]);
            if (!active) v = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(smem + OFF_P + h * P_BYTES + (ch >> 3) * 16384 + tid * 128 +
                                      (((ch & 7) ^ (tid & 7)) << 4)) = v;
          }
        }
      }
      // ---- 3. O_h = P_h V  (K = 128 keys = 8 steps of 16), S columns are dead now
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (issuer >= 0) {
        tc_fence_after();
        const int h = issuer;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t pa = sbase + OFF_P + h * P_BYTES + (ks >> 2) * 16384 + (ks & 3) * 32;
          // V advances 16 keys = 2 K-groups of 256 B per step
          tc_mma_bf16(tmem + (h == 0 ? TM_O0 : TM_O1), make_sw128_kmajor_desc(pa),
                      make_nosw_desc(sbase + OFF_V + ks * 512, 256, 128), ID_O, ks ? 1u : 0u);
        }
        tc_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1u;
      tc_fence_after();
      {
        float o[U];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          uint32_t t16[16];
          tc_ld_32x16(tmem + lane_base + (h == 0 ? TM_O0 : TM_O1), t16);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < DH; ++e) o[h * DH + e] = __uint_as_float(t16[h * DH + e]) * linv[h];
        }
        // residual, ReLU, LayerNorm (InteractingLayer.py:57-60)
        float a[U], mean, rstd;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = fmaxf(use_res ? o[u] + r[u] : o[u], 0.f);
        ln_row_stats<U>(a, eps, mean, rstd);
#pragma unroll
        for (int u = 0; u < U; ++u) yv[u] = ln_apply(a[u], mean, rstd, gs[u], be[u]);
        // saved for the backward: the pre-LayerNorm activations of EVERY iteration (the backward
        // re-derives each iteration's input as LayerNorm(a) and differentiates ReLU/LayerNorm at a)
        if (active && saved) {
          float* sp = saved + ((int64_t)it * B * F + smp * F + f_loc) * U;
#pragma unroll
          for (int u = 0; u < U; u += 4) *reinterpret_cast<float4*>(sp + u) = make_float4(a[u], a[u + 1], a[u + 2], a[u + 3]);
        }
      }
      if (it + 1 < L) {
#pragma unroll
        for (int u = 0; u < U; ++u) xr[u] = active ? yv[u] : 0.f;
      }
    }
    if (active) {
#pragma unroll
      for (int u = 0; u < U; u += 4)
        store4<T>(y + smp * y_bs + (int64_t)f_loc * y_ld + u, make_float4(yv[u], yv[u + 1], yv[u + 2], yv[u + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}

template <int NCHF, typename T>
static int launch_itc_fwd(const IFwdArgs& a) {
  auto kern = interacting_tc_fwd_kernel<16, 16, 2, NCHF, T>;
  constexpr int smem = 2 * 128 * 256 + 16384 + 4096 + 2 * 4096 + 2 * 4096 + 8192 + 96 * 4 + 64 + 1024;
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  constexpr int SPT = 128 / (NCHF * 8);
  const int ntiles = (a.B + SPT - 1) / SPT;
  int grid = sm_count() * 2;
  if (grid > ntiles) grid = ntiles;
  kern<<<grid, 128, smem, a.st>>>((const T*)a.x, a.x_ld, a.x_bs, a.W, a.b, a.gm, a.bt, a.eps, (T*)a.y, a.y_ld,
                                  a.y_bs, (float*)a.saved, a.B, a.F, a.L, a.use_res);
  return check_launch("interacting_tc_fwd");
}

// Returns RS_ERR_UNSUPPORTED (without setting an error) when the shape is not built for the
// tensor-core path; the caller then uses the FFMA kernels.
// Built for D = U = 16, 2 heads, bf16 activations and up to 48 fields: a tile holds floor(128 / FP) whole samples
// with FP = F rounded up to 8 (F = 39 -> 3 samples, 26 -> 4, 16 -> 8, 48 -> 2).
bool interacting_tc_supported(int F, int D, int U, int H, int dtype) {
  return dtype == RS_BF16 && D == 16 && U == 16 && H == 2 && F >= 1 && F <= 48;
}

int interacting_tc_fwd(const IFwdArgs& a) {
  switch ((a.F + 7) / 8) {
    case 1: return launch_itc_fwd<1, __nv_bfloat16>(a);
    case 2: return launch_itc_fwd<2, __nv_bfloat16>(a);
    case 3: return launch_itc_fwd<3, __nv_bfloat16>(a);
    case 4: return launch_itc_fwd<4, __nv_bfloat16>(a);
    case 5: return launch_itc_fwd<5, __nv_bfloat16>(a);
    default: return launch_itc_fwd<6, __nv_bfloat16>(a);
  }
}

}  // namespace rs
