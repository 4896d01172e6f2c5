// interacting_tc_bwd.cu — K4 backward on the 5th-gen tensor cores (bf16 mode): the gradient of
// InteractingLayer.call (InteractingLayer.py:37-61) with every contraction issued as tcgen05.mma.
//
// One CTA of 256 threads per SM owns a tile of SPT whole samples (sample s = tile rows
// [s*FP, s*FP+F), FP = F rounded up to 8; F = 39 -> 3 samples per 128-row tile).  Two warpgroups
// share the 128 TMEM lanes: thread (wg, row) IS tile row `row`; warpgroup h does the per-head work
// of head h (softmax row, dS row, its 8 columns of q/k/v and their gradients) and both do the
// cheap row-wise LayerNorm/ReLU algebra redundantly, so no value ever crosses threads except
// through an MMA operand.  Flash-style: the forward saved only the pre-LayerNorm activations a of
// each iteration (64 B per row); iteration inputs are re-derived as LayerNorm(a), the attention is
// recomputed with the forward's own arithmetic, and P.V is never needed again (the softmax
// backward takes delta = sum_j P_ij dP_ij).  Per iteration (last to first, weights shared):
//
//   1. Z = X Wqkvr                    3xTF32        -> q k v r (ReLU), masks; LayerNorm/ReLU backward
//                                                      at the stored a -> dT (= dO, dR)
//   2. S_h = Q_h K_h^T                tf32, K = 8   -> P = softmax row (bf16 tile in smem)
//   3. dP_h = dO_h V_h^T              tf32, K = 8
//      dV_h = P_h^T dO_h              bf16, A = the P tile read MN-major (transposed)
//                                      -> dS = P (dP - delta) / sqrt(dh)  overwrites P in place
//   4. dQ_h = dS_h K_h ; dK_h = dS_h^T Q_h          bf16 (dS tile read K-major and MN-major)
//   5. dX = dZ Wqkvr^T                bf16, K = 64  -> gradient of the previous iteration's output
//      [dW^T | db ; dgamma ; dbeta] += [dZ | g*xhat | g]^T [X | 1]      bf16, K = 128 rows,
//      accumulated in TMEM for the CTA's whole lifetime (fixed order => deterministic)
//
// Every shared-memory operand tile is the no-swizzle canonical layout [row/8][chunk][row%8][16 B]
// written by the thread that owns the row; the same bytes serve as K-major and MN-major operand.
// Rows outside a sample (padding fields, the 8 spare rows, samples past B) carry g = 0 and P = 0,
// which makes every gradient they could contribute exactly zero.
#include "tc_common.cuh"
#include "interacting_args.cuh"

namespace rs {

// own-sample window (FP columns) of a [128 x 128] TMEM accumulator; a warp whose 32 lanes span
// two samples loads both windows (tcgen05.ld is warp-wide) and each lane keeps its own.
template <int FP>
__device__ __forceinline__ void ld_window(uint32_t taddr, int ws_lo, int ws_hi, int s_loc, int F, float fill,
                                          float (&out)[FP]) {
#pragma unroll
  for (int j = 0; j < FP; ++j) out[j] = fill;
  for (int s = ws_lo; s <= ws_hi; ++s) {
    uint32_t t[FP];
#pragma unroll
    for (int c0 = 0; c0 < FP; c0 += 8) {
      uint32_t t8[8];
      tc_ld_32x8(taddr + (uint32_t)(s * FP + c0), t8);
#pragma unroll
      for (int j = 0; j < 8; ++j) t[c0 + j] = t8[j];
    }
    tc_wait_ld();
    const bool mine = s == s_loc;
#pragma unroll
    for (int j = 0; j < FP; ++j)
      if (mine && j < F) out[j] = __uint_as_float(t[j]);
  }
}

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// element e of this thread's head inside a 16-wide register row (compile-time indices + select:
// a runtime index would push the array to local memory)
#ifdef RS_ITB_PROFILE
// what-if timing experiments (WRONG results, profile builds only; tools/itb_whatif.py): bit 0 = no MMA issue,
// bit 1 = no fence.proxy.async, bit 2 = no exponentials in the softmax, bit 3 = no P / dS tile stores;
// bit 8 = clock64 phase hooks on (tools/itb_profile.py)
__device__ int itb_exp_mode = 0;
#define EXP(bit) (itb_exp_mode & (1 << (bit)))
__device__ unsigned long long itb_prof[32];
#define PROF(i)                                         \
  if (blockIdx.x == 0 && tid == 0 && EXP(8)) {          \
    const long long t_now = clock64();                  \
    itb_prof[i] += (unsigned long long)(t_now - t_last); \
    t_last = t_now;                                     \
  }
#define PROFW(i)                                                                     \
  if (blockIdx.x == 0 && (tid == 0 || tid == 32) && EXP(8)) {                         \
    const long long t_now = clock64();                                               \
    itb_prof[(tid == 0 ? 16 : 24) + (i)] += (unsigned long long)(t_now - t_w);        \
    t_w = t_now;                                                                     \
  }
#else
#define PROF(i)
#define PROFW(i)
#define EXP(bit) 0
#endif

#define HSEL(a, e) (wg ? (a)[8 + (e)] : (a)[(e)])
#define HSELF(a, e) __uint_as_float(HSEL(a, e))

template <int NCHF, typename T>
__global__ void __launch_bounds__(256, 1)
interacting_tc_bwd_kernel(const T* __restrict__ x, int64_t x_ld, int64_t x_bs, const float* __restrict__ saved,
                          const float* __restrict__ W, const float* __restrict__ bias,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                          const T* __restrict__ dy, int64_t dy_ld, int64_t dy_bs, T* __restrict__ dx, int64_t dx_ld,
                          int64_t dx_bs, float* __restrict__ part, int B, int F, int L, int use_res) {
  constexpr int D = 16, U = 16, H = 2, DH = 8, N4 = 64;
  constexpr int FP = NCHF * 8;
  constexpr int SPT = 128 / FP;
  // ---- shared memory map
  constexpr int OFF_P = 0;                         // H x [128][128] bf16, 16 chunks/row: P then dS
  constexpr int OFF_DZ = OFF_P + H * 32768;        // [128][128] bf16: dq|dk|dv|dr | g*xhat | g | 0
  constexpr int OFF_XB = OFF_DZ + 32768;           // [128][32] bf16: x | 1 0..0 | 0      (B of dW)
  constexpr int OFF_X = OFF_XB + 16384;            // [128][x_hi | x_lo] tf32, 8 chunks/row (A of Z)
  constexpr int OFF_Q32 = OFF_X + 16384;           // H x [128][8] tf32
  constexpr int OFF_K32 = OFF_Q32 + H * 4096;
  constexpr int OFF_V32 = OFF_K32 + H * 4096;
  constexpr int OFF_DO32 = OFF_V32 + H * 4096;
  constexpr int OFF_Q16 = OFF_DO32 + H * 4096;     // [128][16] bf16, MN-major B operands
  constexpr int OFF_K16 = OFF_Q16 + 4096;
  constexpr int OFF_DO16 = OFF_K16 + 4096;
  constexpr int OFF_W32 = OFF_DO16 + 4096;         // W_hi | W_lo, each [64 n][16 k] tf32  (B of Z)
  constexpr int OFF_WT = OFF_W32 + 8192;           // [16 n = d][64 k] bf16                (B of dX)
  constexpr int OFF_F = OFF_WT + 2048;             // bias[64] gamma[16] beta[16]
  constexpr int OFF_BAR = OFF_F + (N4 + 2 * U) * 4;
  extern __shared__ uint8_t itb_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(itb_smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bs = reinterpret_cast<float*>(smem + OFF_F);
  float* gs = bs + N4;
  float* be = gs + U;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int wg = tid >> 7;                  // warpgroup = head this thread works for
  const int row = tid & 127;                // tile row = TMEM lane
  if (tid == 0) {
    mbar_init(bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // zero P/dS, dZ and XB; then the constant ones column of XB (chunk 2, element 0)
  for (int i = tid; i < (OFF_X - OFF_P) / 16; i += 256) reinterpret_cast<uint4*>(smem + OFF_P)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < N4 + 2 * U; i += 256) bs[i] = i < N4 ? bias[i] : (i < N4 + U ? gamma[i - N4] : beta[i - N4 - U]);
  stage_w_3xtf32(smem + OFF_W32, W, tid, 256);     // B of Z
  for (int i = tid; i < D * 8; i += 256) {         // B of dX: row n = d, chunk c = 8 consecutive n4 (bf16)
    const int d = i >> 3, c = i & 7;
    float w8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w8[e] = W[d * N4 + c * 8 + e];
    *reinterpret_cast<uint4*>(smem + OFF_WT + nosw_off<8>(d, c)) = pack8_bf16(w8);
  }
  __syncthreads();
  if (tid < 128) *reinterpret_cast<uint4*>(smem + OFF_XB + nosw_off<4>(tid, 2)) = make_uint4(0x00003F80u, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);     // this warp's TMEM lanes
  constexpr uint32_t TM_S = 0;        // S_h / dP_h at h*128
  constexpr uint32_t TM_Z = 256;      // 64
  constexpr uint32_t TM_DV = 352;     // +h*16
  constexpr uint32_t TM_DQ = 384;
  constexpr uint32_t TM_DK = 416;
  constexpr uint32_t TM_DX = 448;     // 16
  constexpr uint32_t TM_DW = 464;     // 32, persistent

  const uint32_t sbase = smem_u32(smem);
  const uint32_t b16 = sbase >> 4;                              // descriptor start-address unit
  constexpr uint32_t ID_Z = make_idesc(2, 128, N4, 0, 0);       // tf32
  constexpr uint32_t ID_S = make_idesc(2, 128, 128, 0, 0);      // tf32 (S and dP)
  constexpr uint32_t ID_AK = make_idesc(1, 128, 16, 0, 1);      // bf16, A K-major, B MN-major (dQ)
  constexpr uint32_t ID_AT = make_idesc(1, 128, 16, 1, 1);      // bf16, A MN-major (transposed), B MN-major (dV, dK)
  constexpr uint32_t ID_DX = make_idesc(1, 128, 16, 0, 0);      // bf16, both K-major
  constexpr uint32_t ID_DW = make_idesc(1, 128, 32, 1, 1);      // bf16, A = dZ^T, B = [X|1]

  const int s_loc = row / FP, f_loc = row - s_loc * FP;
  const int ntiles = (B + SPT - 1) / SPT;
  const float scale = 1.f / sqrtf((float)DH);
  const float scale_log2 = ITC_LOG2E * scale;
  const int wq = warp & 3;
  const int ws_lo = (wq * 32) / FP, ws_hi = min((wq * 32 + 31) / FP, SPT - 1);
  const int64_t total_rows = (int64_t)B * F;
  const bool row_ok = s_loc < SPT && f_loc < F;
  const int hb = wg * DH;              // first column of this thread's head
  // Four MMA issuers (lane 0 of warps 0, 1, 4, 5): independent accumulator chains are issued
  // concurrently; every issuer commits to the one mbarrier each phase (count 4).
  const int issuer = tid == 0 ? 0 : (tid == 32 ? 1 : (tid == 128 ? 2 : (tid == 160 ? 3 : -1)));
#ifdef RS_ITB_PROFILE
  long long t_last = clock64();
  long long t_w = clock64();
#endif
  uint32_t phase = 0;
  uint32_t dw_acc = 0;                 // 0 until the first dW MMA of this CTA

  // x-source row of step (tile, it): the bf16 layer input (it == 0) or the stored activations of
  // iteration it-1 (LayerNorm is applied when the row is staged)
  auto load_xsrc = [&](int tile_, int it_, float (&dst)[U]) {
    const int64_t smp_ = (int64_t)tile_ * SPT + s_loc;
    if (row_ok && smp_ < B) {
      if (it_ == 0) {
#pragma unroll
        for (int c = 0; c < U; c += 4) {
          const float4 t4 = load4<T>(x + smp_ * x_bs + (int64_t)f_loc * x_ld + c);
          dst[c] = t4.x; dst[c + 1] = t4.y; dst[c + 2] = t4.z; dst[c + 3] = t4.w;
        }
      } else {
        const float* sp = saved + ((int64_t)(it_ - 1) * total_rows + smp_ * F + f_loc) * U;
#pragma unroll
        for (int c = 0; c < U; c += 4) {
          const float4 t4 = ldg_nc_f4(reinterpret_cast<const float4*>(sp + c));
          dst[c] = t4.x; dst[c + 1] = t4.y; dst[c + 2] = t4.z; dst[c + 3] = t4.w;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < U; ++c) dst[c] = 0.f;
    }
  };
  // first-step-of-a-tile rows: stored activations of the last iteration and the incoming gradient
  auto load_tile_head = [&](int tile_, float (&a_)[U], float (&g_)[U]) {
    const int64_t smp_ = (int64_t)tile_ * SPT + s_loc;
    if (row_ok && smp_ < B) {
      const float* sp = saved + ((int64_t)(L - 1) * total_rows + smp_ * F + f_loc) * U;
#pragma unroll
      for (int c = 0; c < U; c += 4) {
        const float4 t4 = ldg_nc_f4(reinterpret_cast<const float4*>(sp + c));
        a_[c] = t4.x; a_[c + 1] = t4.y; a_[c + 2] = t4.z; a_[c + 3] = t4.w;
        const float4 d4 = load4<T>(dy + smp_ * dy_bs + (int64_t)f_loc * dy_ld + c);
        g_[c] = d4.x; g_[c + 1] = d4.y; g_[c + 2] = d4.z; g_[c + 3] = d4.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < U; ++c) { a_[c] = 0.f; g_[c] = 0.f; }
    }
  };
  // stage this warpgroup's half of the step input: X tile (3xTF32 split); returns the half row
  auto stage_x = [&](const float (&src)[U], int it_, bool act_, float (&xh)[8]) {
    if (act_ && it_ > 0) {
      float mean, rstd;
      ln_row_stats<U>(src, eps, mean, rstd);      // bit-identical to the forward's LayerNorm
#pragma unroll
      for (int e = 0; e < 8; ++e) xh[e] = ln_apply(HSEL(src, e), mean, rstd, gs[hb + e], be[hb + e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) xh[e] = act_ ? HSEL(src, e) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
      stage_x4_3xtf32(smem + OFF_X, row, wg * 2 + c, xh[c * 4], xh[c * 4 + 1], xh[c * 4 + 2], xh[c * 4 + 3]);
  };

  // ---- prologue: first step's rows, X tile, Z MMA
  int tile = blockIdx.x, it = L - 1;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nsteps = my_tiles * L;
  bool active = row_ok && (int64_t)tile * SPT + s_loc < B;
  float a[U], g[U], xs[U];
  load_tile_head(tile, a, g);
  load_xsrc(tile, it, xs);
  {
    float xh[8];
    stage_x(xs, it, active, xh);
    *reinterpret_cast<uint4*>(smem + OFF_XB + nosw_off<4>(row, wg)) = pack8_bf16(xh);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (issuer >= 0) {
    tc_fence_after();
    if (issuer == 0) issue_proj_3xtf32(tmem + TM_Z, sbase + OFF_X, sbase + OFF_W32, ID_Z);
    tc_commit(bar);
  }
  mbar_wait(bar, phase); phase ^= 1u;
  tc_fence_after();

  for (int step = 0; step < nsteps; ++step) {
    const bool last = step + 1 == nsteps;
    const bool new_tile = it == 0;                 // the next step starts another tile
    const int ntile = new_tile ? tile + (int)gridDim.x : tile;
    const int nit = new_tile ? L - 1 : it - 1;
    const bool nactive = !last && row_ok && (int64_t)ntile * SPT + s_loc < B;
    const int64_t smp = (int64_t)tile * SPT + s_loc;
    // ================= E1a. Z -> q k (the operands of S) ; v, r pre-activations stay in registers
    uint32_t qmask = 0, kmask = 0, vmask = 0;
    uint32_t zv[8], zr[16];
    {
      uint32_t zq[8], zk[8];
      tc_ld_32x8(tl + TM_Z + hb, zq);
      tc_ld_32x8(tl + TM_Z + U + hb, zk);
      tc_ld_32x8(tl + TM_Z + 2 * U + hb, zv);
      tc_ld_32x16(tl + TM_Z + 3 * U, zr);
      tc_wait_ld();
      float q[DH], kk[DH];
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        q[e] = fmaxf(__uint_as_float(zq[e]) + bs[hb + e], 0.f);
        kk[e] = fmaxf(__uint_as_float(zk[e]) + bs[U + hb + e], 0.f);
        qmask |= (q[e] > 0.f ? 1u : 0u) << e;
        kmask |= (kk[e] > 0.f ? 1u : 0u) << e;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        *reinterpret_cast<float4*>(smem + OFF_Q32 + wg * 4096 + nosw_off<2>(row, c)) =
            make_float4(q[c * 4], q[c * 4 + 1], q[c * 4 + 2], q[c * 4 + 3]);
        *reinterpret_cast<float4*>(smem + OFF_K32 + wg * 4096 + nosw_off<2>(row, c)) =
            make_float4(kk[c * 4], kk[c * 4 + 1], kk[c * 4 + 2], kk[c * 4 + 3]);
      }
      *reinterpret_cast<uint4*>(smem + OFF_Q16 + nosw_off<2>(row, wg)) = pack8_bf16(q);
      *reinterpret_cast<uint4*>(smem + OFF_K16 + nosw_off<2>(row, wg)) = pack8_bf16(kk);
    }
    // ================= 2. S_h = Q_h K_h^T
    if (!EXP(1)) fence_async_smem();
    tc_fence_before();
    __syncthreads();
    PROF(0)
    if (issuer >= 0) {
      tc_fence_after();
      if (!EXP(0)) {
      if ((issuer & 1) == 0) {
        const int h = issuer >> 1;
        tc_mma_tf32(tmem + TM_S + h * 128, mk_desc(b16, OFF_Q32 + h * 4096, 128, 256),
                    mk_desc(b16, OFF_K32 + h * 4096, 128, 256), ID_S, 0u);
      }
      }
      tc_commit(bar);
    }
    PROF(1)
    // ================= E1b (under the S MMA). v ; LayerNorm / ReLU backward at the stored a
    {
      float vv[DH];
      uint32_t rmask = 0;
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        vv[e] = fmaxf(__uint_as_float(zv[e]) + bs[2 * U + hb + e], 0.f);
        vmask |= (vv[e] > 0.f ? 1u : 0u) << e;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) rmask |= (__uint_as_float(zr[u]) + bs[3 * U + u] > 0.f ? 1u : 0u) << u;
#pragma unroll
      for (int c = 0; c < 2; ++c)
        *reinterpret_cast<float4*>(smem + OFF_V32 + wg * 4096 + nosw_off<2>(row, c)) =
            make_float4(vv[c * 4], vv[c * 4 + 1], vv[c * 4 + 2], vv[c * 4 + 3]);
      // ---- LayerNorm + ReLU backward at the stored activations a (InteractingLayer.py:59-60)
      float mean, rstd;
      ln_row_stats<U>(a, eps, mean, rstd);
      float xhat[U], s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
      for (int u = 0; u < U; u += 2) {
        xhat[u] = (a[u] - mean) * rstd;
        xhat[u + 1] = (a[u + 1] - mean) * rstd;
        const float g0 = g[u] * gs[u], g1 = g[u + 1] * gs[u + 1];
        s1a += g0; s1b += g1;
        s2a = fmaf(g0, xhat[u], s2a); s2b = fmaf(g1, xhat[u + 1], s2b);
      }
      const float s1 = (s1a + s1b) * (1.f / U), s2 = (s2a + s2b) * (1.f / U);
      // this head's 8 columns of dT = dO = dR
      float dTh[8], t8[8];
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        const float ae = HSEL(a, e), xe = HSEL(xhat, e);
        const float dA = (HSEL(g, e) * gs[hb + e] - s1 - xe * s2) * rstd;
        dTh[e] = (active && ae > 0.f) ? dA : 0.f;
      }
      // operands: dO_h (tf32 A of dP; bf16 B of dV), and the dR, g*xhat, g columns of dZ
#pragma unroll
      for (int c = 0; c < 2; ++c)
        *reinterpret_cast<float4*>(smem + OFF_DO32 + wg * 4096 + nosw_off<2>(row, c)) =
            make_float4(dTh[c * 4], dTh[c * 4 + 1], dTh[c * 4 + 2], dTh[c * 4 + 3]);
      *reinterpret_cast<uint4*>(smem + OFF_DO16 + nosw_off<2>(row, wg)) = pack8_bf16(dTh);
      const uint32_t rm = rmask >> hb;
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = (use_res && ((rm >> e) & 1u)) ? dTh[e] : 0.f;
      *reinterpret_cast<uint4*>(smem + OFF_DZ + nosw_off<16>(row, 6 + wg)) = pack8_bf16(t8);
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = HSEL(g, e) * HSEL(xhat, e);
      *reinterpret_cast<uint4*>(smem + OFF_DZ + nosw_off<16>(row, 8 + wg)) = pack8_bf16(t8);
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = HSEL(g, e);
      *reinterpret_cast<uint4*>(smem + OFF_DZ + nosw_off<16>(row, 10 + wg)) = pack8_bf16(t8);
    }
    // ---- prefetch the next step's rows: a whole step of latency cover
    float xs_n[U], a_n[U], g_n[U];
    if (!last) load_xsrc(ntile, nit, xs_n);
    if (!last && new_tile) load_tile_head(ntile, a_n, g_n);
    mbar_wait(bar, phase); phase ^= 1u;
    PROF(2)
    tc_fence_after();
    float p[FP];                     // normalised attention row of this head
    {
#ifdef RS_ITB_PROFILE
      t_w = clock64();
#endif
      ld_window<FP>(tl + TM_S + wg * 128, ws_lo, ws_hi, s_loc, F, -INFINITY, p);
      PROFW(0)
      float m4[4] = {p[0], p[1], p[2], p[3]};
#pragma unroll
      for (int j = 4; j < FP; ++j) m4[j & 3] = fmaxf(m4[j & 3], p[j]);
      float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      if (!active) m = 0.f;
      const float mb = m * scale_log2;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < FP; j += 2) {                                // the forward's P, bit for bit
        p[j] = EXP(2) ? fmaf(p[j], scale_log2, -mb) : ex2_approx(fmaf(p[j], scale_log2, -mb));
        p[j + 1] = EXP(2) ? fmaf(p[j + 1], scale_log2, -mb) : ex2_approx(fmaf(p[j + 1], scale_log2, -mb));
        bf16_round2(p[j], p[j + 1]);
        l4[j & 3] += p[j];
        l4[(j + 1) & 3] += p[j + 1];
      }
      const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      const float linv = active ? 1.f / l : 0.f;
#pragma unroll
      for (int j = 0; j < FP; ++j) p[j] = active ? p[j] * linv : 0.f;
      PROFW(1)
      if (s_loc < SPT && !EXP(3)) {
#pragma unroll
        for (int c = 0; c < NCHF; ++c)
          *reinterpret_cast<uint4*>(smem + OFF_P + wg * 32768 + nosw_off<16>(row, s_loc * NCHF + c)) = pack8_bf16(p + c * 8);
      }
    }
    // ================= 3. dP_h = dO_h V_h^T ; dV_h = P_h^T dO_h
    PROFW(2)
    if (!EXP(1)) fence_async_smem();
    PROFW(3)
    tc_fence_before();
    __syncthreads();
    PROFW(4)
    PROF(3)
    if (issuer >= 0) {
      tc_fence_after();
      if (!EXP(0)) {
      const int h = issuer >> 1;
      if ((issuer & 1) == 0) {
        tc_mma_tf32(tmem + TM_S + h * 128, mk_desc(b16, OFF_DO32 + h * 4096, 128, 256),
                    mk_desc(b16, OFF_V32 + h * 4096, 128, 256), ID_S, 0u);
      } else {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // K = tile rows, 16 per step = 2 row groups of the P tile
          tc_mma_bf16(tmem + TM_DV + h * 16, mk_desc(b16, OFF_P + h * 32768 + ks * 4096, 2048, 128),
                      mk_desc(b16, OFF_DO16 + ks * 512, 256, 128), ID_AT, ks ? 1u : 0u);
      }
      }
      tc_commit(bar);
    }
    PROF(4)
    mbar_wait(bar, phase); phase ^= 1u;
    PROF(5)
    tc_fence_after();
    {
      float dp[FP];
      ld_window<FP>(tl + TM_S + wg * 128, ws_lo, ws_hi, s_loc, F, 0.f, dp);
      // delta = sum_j P_ij dP_ij from the very P and dP used (not dO.o): the softmax Jacobian then
      // annihilates any common-mode error of dP exactly (sum_j dS_ij = 0)
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < FP; ++j) d4[j & 3] = fmaf(p[j], dp[j], d4[j & 3]);
      const float delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
#pragma unroll
      for (int j = 0; j < FP; ++j) dp[j] = p[j] * scale * (dp[j] - delta);     // dS (1/sqrt(dh) folded in)
      if (s_loc < SPT && !EXP(3)) {
#pragma unroll
        for (int cc = 0; cc < NCHF; ++cc)
          *reinterpret_cast<uint4*>(smem + OFF_P + wg * 32768 + nosw_off<16>(row, s_loc * NCHF + cc)) = pack8_bf16(dp + cc * 8);
      }
      uint32_t dv[16];
      tc_ld_32x16(tl + TM_DV + wg * 16, dv);
      tc_wait_ld();
      float t8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = ((vmask >> e) & 1u) ? HSELF(dv, e) : 0.f;
      *reinterpret_cast<uint4*>(smem + OFF_DZ + nosw_off<16>(row, 4 + wg)) = pack8_bf16(t8);
    }
    // ================= 4. dQ_h = dS_h K_h ; dK_h = dS_h^T Q_h
    if (!EXP(1)) fence_async_smem();
    tc_fence_before();
    __syncthreads();
    PROF(6)
    if (issuer >= 0) {
      tc_fence_after();
      if (!EXP(0)) {
      const int h = issuer >> 1;
      if ((issuer & 1) == 0) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          tc_mma_bf16(tmem + TM_DQ + h * 16, mk_desc(b16, OFF_P + h * 32768 + ks * 256, 128, 2048),
                      mk_desc(b16, OFF_K16 + ks * 512, 256, 128), ID_AK, ks ? 1u : 0u);
      } else {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          tc_mma_bf16(tmem + TM_DK + h * 16, mk_desc(b16, OFF_P + h * 32768 + ks * 4096, 2048, 128),
                      mk_desc(b16, OFF_Q16 + ks * 512, 256, 128), ID_AT, ks ? 1u : 0u);
      }
      }
      tc_commit(bar);
    }
    PROF(7)
    // under the dQ / dK MMAs: the NEXT step's X tile (its Z MMA rides in the same phase as this
    // step's dX / dW)
    float xh_n[8];
    if (!last) stage_x(xs_n, nit, nactive, xh_n);
    mbar_wait(bar, phase); phase ^= 1u;
    PROF(8)
    tc_fence_after();
    {
      uint32_t dq[16], dk[16];
      tc_ld_32x16(tl + TM_DQ + wg * 16, dq);
      tc_ld_32x16(tl + TM_DK + wg * 16, dk);
      tc_wait_ld();
      float t8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = ((qmask >> e) & 1u) ? HSELF(dq, e) : 0.f;
      *reinterpret_cast<uint4*>(smem + OFF_DZ + nosw_off<16>(row, wg)) = pack8_bf16(t8);
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = ((kmask >> e) & 1u) ? HSELF(dk, e) : 0.f;
      *reinterpret_cast<uint4*>(smem + OFF_DZ + nosw_off<16>(row, 2 + wg)) = pack8_bf16(t8);
    }
    // ================= 5. dX = dZ W^T ; [dW^T | db ; dgamma ; dbeta] += [dZ | g*xhat | g]^T [X | 1] ; Z(next)
    if (!EXP(1)) fence_async_smem();
    tc_fence_before();
    __syncthreads();
    PROF(9)
    if (issuer >= 0) {
      tc_fence_after();
      if (!EXP(0)) {
      if (issuer == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_bf16(tmem + TM_DX, mk_desc(b16, OFF_DZ + ks * 256, 128, 2048),
                      mk_desc(b16, OFF_WT + ks * 256, 128, 1024), ID_DX, ks ? 1u : 0u);
      } else if (issuer == 1) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          tc_mma_bf16(tmem + TM_DW, mk_desc(b16, OFF_DZ + ks * 4096, 2048, 128),
                      mk_desc(b16, OFF_XB + ks * 1024, 512, 128), ID_DW, ks ? 1u : dw_acc);
      } else if (issuer == 2 && !last) {
        issue_proj_3xtf32(tmem + TM_Z, sbase + OFF_X, sbase + OFF_W32, ID_Z);
      }
      }
      tc_commit(bar);
    }
    dw_acc = 1u;
    PROF(10)
    mbar_wait(bar, phase); phase ^= 1u;
    PROF(11)
    tc_fence_after();
    {
      uint32_t dxr[16];
      tc_ld_32x16(tl + TM_DX, dxr);
      tc_wait_ld();
      if (it > 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = active ? __uint_as_float(dxr[u]) : 0.f;   // stays fp32 between iterations
      } else {
        if (active) {
          T* dp = dx + smp * dx_bs + (int64_t)f_loc * dx_ld + hb;
          store4<T>(dp, make_float4(HSELF(dxr, 0), HSELF(dxr, 1), HSELF(dxr, 2), HSELF(dxr, 3)));
          store4<T>(dp + 4, make_float4(HSELF(dxr, 4), HSELF(dxr, 5), HSELF(dxr, 6), HSELF(dxr, 7)));
        }
        if (!last) {
#pragma unroll
          for (int u = 0; u < U; ++u) g[u] = g_n[u];
        }
      }
      if (!last) {
        // dW of this step has consumed XB: the next step's bf16 x row may land now
        *reinterpret_cast<uint4*>(smem + OFF_XB + nosw_off<4>(row, wg)) = pack8_bf16(xh_n);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          a[u] = new_tile ? a_n[u] : xs[u];      // same tile: a of iteration it-1 = this step's x-source
          xs[u] = xs_n[u];
        }
      }
    }
    tile = ntile; it = nit; active = nactive;
  }
  // ---- per-CTA partials in the layout of the FFMA kernel: dW[D][4U] | db[4U] | dgamma[U] | dbeta[U]
  if (wg == 0) {
    uint32_t acc[32];
    tc_ld_32x32(tl + TM_DW, acc);
    float* mine = part + (int64_t)blockIdx.x * (D * N4 + 6 * U);
    if (row < N4) {
#pragma unroll
      for (int d = 0; d < D; ++d) mine[d * N4 + row] = dw_acc ? __uint_as_float(acc[d]) : 0.f;
    }
    if (row < 6 * U) mine[D * N4 + row] = dw_acc ? __uint_as_float(acc[16]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// out[i] = sum over CTAs of part[cta][i]; one warp per output, fixed order (deterministic)
static __global__ void itb_reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                  int nparts, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const float s = warp_ordered_sum(part + i, nparts, n);
  if ((threadIdx.x & 31) == 0) out[i] = s;
}

template <int NCHF, typename T>
static int launch_itc_bwd(const IBwdArgs& a) {
  auto kern = interacting_tc_bwd_kernel<NCHF, T>;
  constexpr int smem = 2 * 32768 + 32768 + 16384 + 16384 + 4 * 2 * 4096 + 3 * 4096 + 8192 + 2048 + (64 + 32) * 4 + 64 + 1024;
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  constexpr int SPT = 128 / (NCHF * 8);
  const int ntiles = (a.B + SPT - 1) / SPT;
  int grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  const int np = 16 * 64 + 6 * 16;
  if (a.saved == nullptr) {
    set_error("interacting_tc_bwd: the saved activations of the tensor-core forward are required");
    return RS_ERR_INVALID;
  }
  if (a.ws_bytes < (size_t)grid * np * sizeof(float)) {
    set_error("interacting_tc_bwd: workspace %zu < %zu", a.ws_bytes, (size_t)grid * np * sizeof(float));
    return RS_ERR_WORKSPACE;
  }
  kern<<<grid, 256, smem, a.st>>>((const T*)a.x, a.x_ld, a.x_bs, (const float*)a.saved, a.W, a.b, a.gm, a.bt, a.eps,
                                  (const T*)a.dy, a.dy_ld, a.dy_bs, (T*)a.dx, a.dx_ld, a.dx_bs, (float*)a.ws, a.B, a.F,
                                  a.L, a.use_res);
  if (int e = check_launch("interacting_tc_bwd")) return e;
  itb_reduce_partials_kernel<<<(np * 32 + 255) / 256, 256, 0, a.st>>>((const float*)a.ws, a.dparams, grid, np);
  return check_launch("interacting_tc_bwd_reduce");
}

int interacting_tc_bwd(const IBwdArgs& a) {
  switch ((a.F + 7) / 8) {
    case 1: return launch_itc_bwd<1, __nv_bfloat16>(a);
    case 2: return launch_itc_bwd<2, __nv_bfloat16>(a);
    case 3: return launch_itc_bwd<3, __nv_bfloat16>(a);
    case 4: return launch_itc_bwd<4, __nv_bfloat16>(a);
    case 5: return launch_itc_bwd<5, __nv_bfloat16>(a);
    default: return launch_itc_bwd<6, __nv_bfloat16>(a);
  }
}

#ifdef RS_ITB_PROFILE
extern "C" int rs_debug_itb_exp(int mode) {
  cudaDeviceSynchronize();
  cudaMemcpyToSymbol(itb_exp_mode, &mode, sizeof(int));
  return 0;
}
extern "C" int rs_debug_itb_profile(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  if (out32) cudaMemcpyFromSymbol(out32, itb_prof, sizeof(unsigned long long) * 32);
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(itb_prof, z, sizeof(z));
  }
  return 0;
}
#endif

}  // namespace rs
