// interacting_tc_bwd.cu — K4 backward on the 5th-gen tensor cores (bf16 mode): the gradient of
// InteractingLayer.call (InteractingLayer.py:37-61) with every contraction issued as tcgen05.mma.
//
// CTA = 128 threads = one 128-row tile of SPT whole samples (sample s = tile rows [s*FP, s*FP+F), FP = F rounded
// up to 8; F = 39 -> 3 samples); thread t IS tile row t and TMEM lane t; two CTAs share an SM (256 TMEM columns,
// ~97 KB of shared memory each).  Flash-style: the forward saved the pre-LayerNorm activations a of every
// iteration (64 B per row) and the per-head softmax statistics lse (8 B per row); iteration inputs are re-derived
// as LayerNorm(a) and P = exp2(S c - lse) is recomputed already normalised (no max / sum pass).  Per (tile,
// iteration), last iteration first, weights shared:
//
//   Z = [X | 1] [W ; b]         3xTF32, A = [x_hi | x_lo | 1 1 0..] in TMEM        -> q k v (ReLU), r mask;
//                                                                                     LayerNorm/ReLU backward -> dT
//   per head h (the two heads are pipelined: the MMAs of one run under the thread work of the other):
//   S_h  = Qx_h Kx_h^T          tf32, A in TMEM.  Qx is the row's q placed in the K slot of its own sample
//   dP_h = dOx_h Vx_h^T         (zeros elsewhere), Kx row j = [k_(0,j) | k_(1,j) | ..]: every row finds ITS sample's
//                               FP keys in the same FP accumulator columns (no 128-wide block-diagonal product)
//        -> P, dS = P (dP - sum_j P dP) as compact [128][FP] bf16 tiles in shared memory
//   dQ_h = dS_h K_h             bf16, A = the dS tile K-major, B = [key][(sample, e)]: a row keeps its sample's 8 columns
//   dV_h = P_h^T dO_h           bf16, M = 64 per sample: A = the compact tile read MN-major (transposed) over the
//   dK_h = dS_h^T Q_h           sample's rows, B = dO / Q rows of that sample; dV lands in TMEM lanes 0-15 of each
//                               subpartition, dK in lanes 16-31 (interleaved half-subpartitions): one tcgen05.ld
//                               hands 32 lanes their (key, head) gradient rows, which they mask and store to dZ
//   dX = dZ W^T ; [dW^T | db ; dgamma ; dbeta] += [dZ | g*xhat | g]^T [X | 1] (accumulated in TMEM for the CTA's
//   lifetime, fixed order => deterministic) ; Z of the NEXT step rides in the same phase.
//
// Every shared-memory operand tile is the no-swizzle canonical layout [row/8][chunk][row%8][16 B] written by the
// thread that owns the row; the same bytes serve as K-major and MN-major operand.  Rows outside a sample carry
// g = 0 and lse = +inf (P = 0), which makes every gradient they could contribute exactly zero.
#include "tc_common.cuh"
#include "interacting_args.cuh"

namespace rs {

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}
// 0xFFFF in each half whose bf16 value is > 0
__device__ __forceinline__ uint32_t bf16x2_pos_mask(uint32_t v) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v), z);
}

struct uint4x2_t { uint4 a, b; };

// fused gradient push (K3's exchange inside K4's backward): the finished dx row of lookup i goes straight into slot
// (rank * cap + k) of the receive buffer of owner = inverse[i] / cap over NVLink (peer stores), the layout the
// all-to-all of the routed gradients would have produced.  inverse == NULL: plain store to dx.
struct ItbScatter {
  void* recv[RS_MAX_PEERS];
  const int32_t* inverse;
  int cap, rank;
};

template <int NCHF> struct ItbSmem {
  using G = ItcGeom<NCHF>;
  static constexpr int PC_BYTES = 16 * NCHF * 128 + 1024;    // [128][FP] bf16 + zeroed pad (operand over-reads)
  static constexpr int XR = G::SPT * G::KP;                  // rows of the sample-expanded [row][16] bf16 operands
  static constexpr int X16_BYTES = (XR / 8) * 256;
  static constexpr int OFF_PC = 0;                           // P   (A of dV, transposed)
  static constexpr int OFF_DS = PC_BYTES;                    // dS  (A of dQ K-major, of dK transposed)
  static constexpr int OFF_KX = 2 * PC_BYTES;                // H x [FP keys][8 SPT] tf32           (B of S)
  static constexpr int OFF_VX = OFF_KX + 2 * G::KX_BYTES;    // H x same, v / sqrt(dh)             (B of dP)
  static constexpr int OFF_K16 = OFF_VX + 2 * G::KX_BYTES;   // H x [KP keys][32] bf16 MN-major     (B of dQ)
  static constexpr int OFF_Q16 = OFF_K16 + 2 * G::VX_BYTES;  // [SPT x KP rows][16] bf16            (B of dK)
  static constexpr int OFF_DO16 = OFF_Q16 + X16_BYTES;       // same                                (B of dV)
  static constexpr int OFF_DZ = OFF_DO16 + X16_BYTES;        // [128][96] bf16: dq dk dv dr | g*xhat | g, 12 chunks/row
  static constexpr int OFF_XB = OFF_DZ + 16 * 12 * 128 + 512;   // [128][32] bf16: x | 1 0.. | 0      (B of dW)
  static constexpr int OFF_W32 = OFF_XB + 8192;              // W_hi | W_lo, each [64 n][16 k] tf32  (B of Z)
  static constexpr int OFF_BT = OFF_W32 + 8192;              // bias tile                           (B of Z)
  static constexpr int OFF_WT = OFF_BT + 2048;               // [16 n = d][64 k] bf16               (B of dX)
  static constexpr int OFF_F = OFF_WT + 2048;                // gamma[16] beta[16]
  static constexpr int OFF_BAR = OFF_F + 128;
  static constexpr int TOTAL = OFF_BAR + 64;
};

template <int NCHF, typename T>
__global__ void __launch_bounds__(128, 2)
interacting_tc_bwd_kernel(const T* __restrict__ x, int64_t x_ld, int64_t x_bs, const float* __restrict__ saved,
                          const float* __restrict__ W, const float* __restrict__ bias,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                          const T* __restrict__ dy, int64_t dy_ld, int64_t dy_bs, T* __restrict__ dx, int64_t dx_ld,
                          int64_t dx_bs, float* __restrict__ part, int B, int F, int L, int use_res,
                          const T* __restrict__ dx_add, const ItbScatter sc) {
  constexpr int D = 16, U = 16, H = 2, DH = 8, N4 = 64;
  using G = ItcGeom<NCHF>;
  using SM = ItbSmem<NCHF>;
  constexpr int FP = G::FP, SPT = G::SPT, NCHK = G::NCHK, KP = G::KP;
  constexpr int FMIN = NCHF == 2 ? 0 : FP - 8;
  extern __shared__ __align__(128) uint8_t itb_smem[];
  uint8_t* smem = itb_smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const float4* gb4 = reinterpret_cast<const float4*>(smem + SM::OFF_F);

  const int row = threadIdx.x;                               // tile row = TMEM lane
  const uint32_t wq = uniform_u32(threadIdx.x >> 5);         // warp = TMEM subpartition (uniform register)
  const int lane = threadIdx.x & 31;
  if (row == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wq == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // zero every operand tile (static zeros: pads, rows of other samples, the spare rows); constants
  for (int i = row; i < SM::OFF_W32 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (row < 2 * U) reinterpret_cast<float*>(smem + SM::OFF_F)[row] = row < U ? gamma[row] : beta[row - U];
  stage_w_3xtf32(smem + SM::OFF_W32, W, row, 128);
  stage_bias_tile(smem + SM::OFF_BT, bias, row, 128);
  {                                                          // B of dX: row n = d, chunk c = 8 consecutive n4 (bf16)
    const int d = row >> 3, c = row & 7;
    float w8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w8[e] = W[d * N4 + c * 8 + e];
    *reinterpret_cast<uint4*>(smem + SM::OFF_WT + nosw_off<8>(d, c)) = pack8_bf16(w8);
  }
  __syncthreads();
  *reinterpret_cast<uint4*>(smem + SM::OFF_XB + nosw_off<4>(row, 2)) = make_uint4(0x00003F80u, 0, 0, 0);   // ones column
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(*tmem_slot);
  const uint32_t tl = tmem + ((wq * 32u) << 16);
  // TMEM columns: A = Z (64), later the dQ / dV|dK accumulators ; expanded q / dO of the current head ;
  // S_h | dP_h, later [x_hi | x_lo | 1 1 0..] (40) and dX (16) ; dW (persistent)
  constexpr uint32_t C_A = 0, C_QX = 64, C_DOX = 96, C_S = 128, C_DP = 176, C_X = 128, C_DX = 168, C_DW = 224;
  {
    uint32_t z16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z16[i] = 0u;
#pragma unroll
    for (int c = 0; c < 64; c += 16) tc_st_32x16(tl + C_QX + c, z16);   // a lane only ever rewrites its own sample's slot
    tc_wait_st();
  }
  const uint32_t sbase = smem_u32(smem);
  const uint32_t b16 = sbase >> 4;
  constexpr uint32_t ID_Z = make_idesc(2, 128, N4, 0, 0);       // tf32
  constexpr uint32_t ID_S = make_idesc(2, 128, FP, 0, 0);       // tf32 (S and dP), K = 8 per sample
  constexpr uint32_t ID_DQ = make_idesc(1, 128, 32, 0, 1);      // bf16, A K-major, B MN-major
  constexpr uint32_t ID_T = make_idesc(1, 64, 8, 1, 1);         // bf16, M = 64, A MN-major (transposed), B MN-major
  constexpr uint32_t ID_DX = make_idesc(1, 128, 16, 0, 0);      // bf16, both K-major
  constexpr uint32_t ID_DW = make_idesc(1, 128, 32, 1, 1);      // bf16, A = dZ^T, B = [X|1]

  const int s_loc = row / FP, f_loc = row - s_loc * FP;
  const bool row_ok = s_loc < SPT && f_loc < F;
  const int ntiles = (B + SPT - 1) / SPT;
  const float scale = 1.f / sqrtf((float)DH);
  const float scale_log2 = ITC_LOG2E * scale;
  const uint32_t ws_lo = min((wq * 32u) / FP, (uint32_t)(SPT - 1)), ws_hi = min((wq * 32u + 31u) / FP, (uint32_t)(SPT - 1));
  const int64_t total_rows = (int64_t)B * F;
  const float* lse_base = saved + (int64_t)L * total_rows * U;
  // dV / dK writer role: lanes 0-15 of warp w own dV of key 16 w + lane, lanes 16-31 dK of the same key
  const int jw = (int)wq * 16 + (lane & 15);
  const bool is_dk = lane >= 16;
  uint32_t phase = 0;
  uint32_t dw_acc = 0;                 // 0 until the first dW MMA of this CTA

  // Rows are prefetched RAW (bf16 pairs stay packed, fp32 stays bits) and converted where they are used: a
  // conversion at the load would make the warp wait for the load right there instead of a step later.
  static_assert(sizeof(T) == 2, "bf16 activations");
  auto load_xsrc = [&](int tile_, int it_, uint32_t (&dst)[U]) {
    const int64_t smp_ = (int64_t)tile_ * SPT + s_loc;
#pragma unroll
    for (int c = 0; c < U; ++c) dst[c] = 0u;
    if (row_ok && smp_ < B) {
      if (it_ == 0) {
#pragma unroll
        for (int c = 0; c < U; c += 4) {
          const uint2 t2 = ldg_nc_u2(reinterpret_cast<const uint2*>(x + smp_ * x_bs + (int64_t)f_loc * x_ld + c));
          dst[c / 2] = t2.x; dst[c / 2 + 1] = t2.y;
        }
      } else {
        const float* sp = saved + ((int64_t)(it_ - 1) * total_rows + smp_ * F + f_loc) * U;
#pragma unroll
        for (int c = 0; c < U; c += 4) {
          const uint4 t4 = ldg_nc_u4(reinterpret_cast<const uint4*>(sp + c));
          dst[c] = t4.x; dst[c + 1] = t4.y; dst[c + 2] = t4.z; dst[c + 3] = t4.w;
        }
      }
    }
  };
  auto load_lse = [&](int tile_, int it_) -> float2 {
    const int64_t smp_ = (int64_t)tile_ * SPT + s_loc;
    if (row_ok && smp_ < B)
      return *reinterpret_cast<const float2*>(lse_base + ((int64_t)it_ * total_rows + smp_ * F + f_loc) * H);
    return make_float2(INFINITY, INFINITY);                  // P = exp2(S c - inf) = 0 for rows outside a sample
  };
  auto load_tile_head = [&](int tile_, float (&a_)[U], uint32_t (&g_)[U / 2]) {
    const int64_t smp_ = (int64_t)tile_ * SPT + s_loc;
    if (row_ok && smp_ < B) {
      const float* sp = saved + ((int64_t)(L - 1) * total_rows + smp_ * F + f_loc) * U;
#pragma unroll
      for (int c = 0; c < U; c += 4) {
        const float4 t4 = ldg_nc_f4(reinterpret_cast<const float4*>(sp + c));
        a_[c] = t4.x; a_[c + 1] = t4.y; a_[c + 2] = t4.z; a_[c + 3] = t4.w;
        const uint2 d2 = ldg_nc_u2(reinterpret_cast<const uint2*>(dy + smp_ * dy_bs + (int64_t)f_loc * dy_ld + c));
        g_[c / 2] = d2.x; g_[c / 2 + 1] = d2.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < U; ++c) a_[c] = 0.f;
#pragma unroll
      for (int c = 0; c < U / 2; ++c) g_[c] = 0u;
    }
  };
  // step input row -> [x_hi | x_lo | 1 1 0..] in TMEM (A of Z); returns the row as packed bf16 (B of dW)
  auto stage_x = [&](const uint32_t (&raw)[U], int it_, bool act_) -> uint4x2_t {
    float xh[U];
    if (act_ && it_ > 0) {
      float src[U];
#pragma unroll
      for (int u = 0; u < U; ++u) src[u] = __uint_as_float(raw[u]);
      float mean, rstd;
      ln_row_stats<U>(src, eps, mean, rstd);      // bit-identical to the forward's LayerNorm
#pragma unroll
      for (int u4 = 0; u4 < U / 4; ++u4) {
        const float4 g4 = gb4[u4], b4 = gb4[U / 4 + u4];
        xh[u4 * 4 + 0] = ln_apply(src[u4 * 4 + 0], mean, rstd, g4.x, b4.x);
        xh[u4 * 4 + 1] = ln_apply(src[u4 * 4 + 1], mean, rstd, g4.y, b4.y);
        xh[u4 * 4 + 2] = ln_apply(src[u4 * 4 + 2], mean, rstd, g4.z, b4.z);
        xh[u4 * 4 + 3] = ln_apply(src[u4 * 4 + 3], mean, rstd, g4.w, b4.w);
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; u += 2) {                      // the bf16 layer input (zeros for rows outside a sample)
        const float2 f2 = unpack_bf16x2(raw[u / 2]);
        xh[u] = f2.x; xh[u + 1] = f2.y;
      }
    }
    uint32_t hi[16], lo[16], one8[8];
#pragma unroll
    for (int c = 0; c < U; ++c) {
      const float h_ = tf32_hi(xh[c]);
      hi[c] = __float_as_uint(h_);
      lo[c] = __float_as_uint(xh[c] - h_);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) one8[i] = i < 2 ? 0x3F800000u : 0u;
    tc_st_32x16(tl + C_X, hi);
    tc_st_32x16(tl + C_X + 16, lo);
    tc_st_32x8(tl + C_X + 32, one8);
    uint4x2_t r;
    r.a = pack8_bf16(xh);
    r.b = pack8_bf16(xh + 8);
    return r;
  };

  // ---- prologue: first step's rows, X operand, Z MMA
  int tile = blockIdx.x, it = L - 1;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nsteps = my_tiles * L;
  bool active = row_ok && (int64_t)tile * SPT + s_loc < B;
  float a[U], g[U];
  uint32_t xs[U];
  float2 lse = load_lse(tile, it);
  {
    uint32_t g_raw[U / 2];
    load_tile_head(tile, a, g_raw);
#pragma unroll
    for (int u = 0; u < U; u += 2) {
      const float2 f2 = unpack_bf16x2(g_raw[u / 2]);
      g[u] = f2.x; g[u + 1] = f2.y;
    }
  }
  load_xsrc(tile, it, xs);
  {
    const uint4x2_t xb = stage_x(xs, it, active);
    *reinterpret_cast<uint4*>(smem + SM::OFF_XB + nosw_off<4>(row, 0)) = xb.a;
    *reinterpret_cast<uint4*>(smem + SM::OFF_XB + nosw_off<4>(row, 1)) = xb.b;
  }
  fence_async_smem();
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  if (wq == 0 && elect_one()) {
    tc_fence_after();
    issue_proj_3xtf32_ts(tmem + C_A, tmem + C_X, sbase + SM::OFF_W32, ID_Z);
    tc_mma_tf32_ts(tmem + C_A, tmem + C_X + 32, make_nosw_desc(sbase + SM::OFF_BT, 128, 256), ID_Z, 1u);
    tc_commit(bar);
  }
  mbar_wait(bar, phase); phase ^= 1u;              // Z of the first step
  tc_fence_after();

  for (int step = 0; step < nsteps; ++step) {
    const bool last = step + 1 == nsteps;
    const bool new_tile = it == 0;                 // the next step starts another tile
    const int ntile = new_tile ? tile + (int)gridDim.x : tile;
    const int nit = new_tile ? L - 1 : it - 1;
    const bool nactive = !last && row_ok && (int64_t)ntile * SPT + s_loc < B;
    const int64_t smp = (int64_t)tile * SPT + s_loc;
    // ================= T1. Z -> q k v (ReLU) for both heads, r mask; LayerNorm / ReLU backward at the stored a
    uint32_t q16[8];                               // this row's q, packed bf16 (ReLU' mask of dq later)
    float q1[DH], dO1[DH];                         // head 1's TMEM operands, staged once head 0's products are done
    {
      uint32_t z[32], zvr[32];
      tc_ld_32x32(tl + C_A, z);                    // q | k
      tc_ld_32x32(tl + C_A + 32, zvr);             // v | r
      float qk[32], vv[U];
#pragma unroll
      for (int u = 0; u < 32; ++u) qk[u] = fmaxf(__uint_as_float(z[u]), 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) vv[u] = fmaxf(__uint_as_float(zvr[u]), 0.f) * scale;   // 1/sqrt(dh) rides on V: dP arrives scaled
      // ---- LayerNorm + ReLU backward (InteractingLayer.py:59-60)
      float mean, rstd;
      ln_row_stats<U>(a, eps, mean, rstd);
      float xhat[U], gg[U];
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
      const float2 nm = make_float2(-mean, -mean), rs2 = make_float2(rstd, rstd);
#pragma unroll
      for (int u4 = 0; u4 < U / 4; ++u4) {
        const float4 g4 = gb4[u4];
        const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int u = u4 * 4 + e;
          const float2 xh2 = fmul2(fadd2(make_float2(a[u], a[u + 1]), nm), rs2);
          const float2 gg2 = fmul2(make_float2(g[u], g[u + 1]), make_float2(gm[e], gm[e + 1]));
          xhat[u] = xh2.x; xhat[u + 1] = xh2.y;
          gg[u] = gg2.x; gg[u + 1] = gg2.y;
          s1 = fadd2(s1, gg2);
          s2 = ffma2(gg2, xh2, s2);
        }
      }
      const float m1 = (s1.x + s1.y) * (1.f / U), m2 = (s2.x + s2.y) * (1.f / U);
      float dT[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float dA = (gg[u] - m1 - xhat[u] * m2) * rstd;
        dT[u] = a[u] > 0.f ? dA : 0.f;             // rows outside a sample have g = 0 -> dA = 0
      }
      // ---- operands of head 0's S / dP first (B tiles in shared memory, expanded q / dO in TMEM), then the MMAs are
      //      issued and everything else of this step's staging runs UNDER them
      auto stage_kv = [&](int h) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          *reinterpret_cast<float4*>(smem + SM::OFF_KX + h * G::KX_BYTES + nosw_off<NCHK>(f_loc, s_loc * 2 + c)) =
              make_float4(qk[U + h * DH + c * 4], qk[U + h * DH + c * 4 + 1], qk[U + h * DH + c * 4 + 2], qk[U + h * DH + c * 4 + 3]);
          *reinterpret_cast<float4*>(smem + SM::OFF_VX + h * G::KX_BYTES + nosw_off<NCHK>(f_loc, s_loc * 2 + c)) =
              make_float4(vv[h * DH + c * 4], vv[h * DH + c * 4 + 1], vv[h * DH + c * 4 + 2], vv[h * DH + c * 4 + 3]);
        }
      };
      if (s_loc < SPT) stage_kv(0);
      // TMEM: head 0's expanded q / dO (own sample's slot; the other slots stay zero)
      for (uint32_t s = ws_lo; s <= ws_hi; ++s) {
        uint32_t v8[8], w8[8];
#pragma unroll
        for (int e = 0; e < DH; ++e) {
          v8[e] = (int)s == s_loc ? __float_as_uint(qk[e]) : 0u;
          w8[e] = (int)s == s_loc ? __float_as_uint(dT[e]) : 0u;
        }
        tc_st_32x8(tl + C_QX + s * 8, v8);
        tc_st_32x8(tl + C_DOX + s * 8, w8);
      }
      fence_async_smem();
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (wq == 0 && elect_one()) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < SPT; ++ks) {
          tc_mma_tf32_ts(tmem + C_S, tmem + C_QX + ks * 8, mk_desc(b16, SM::OFF_KX + ks * 256, 128, NCHK * 128), ID_S, ks ? 1u : 0u);
          tc_mma_tf32_ts(tmem + C_DP, tmem + C_DOX + ks * 8, mk_desc(b16, SM::OFF_VX + ks * 256, 128, NCHK * 128), ID_S, ks ? 1u : 0u);
        }
        tc_commit(bar);
      }
      // ---- under the MMAs: head 1's B tiles, the bf16 operands of dQ / dK / dV, the dr, g*xhat, g columns of dZ
      if (s_loc < SPT) {
        stage_kv(1);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          *reinterpret_cast<uint4*>(smem + SM::OFF_K16 + h * G::VX_BYTES + nosw_off<4>(f_loc, s_loc)) = pack8_bf16(qk + U + h * DH);
          const uint4 qp = pack8_bf16(qk + h * DH);
          q16[h * 4 + 0] = qp.x; q16[h * 4 + 1] = qp.y; q16[h * 4 + 2] = qp.z; q16[h * 4 + 3] = qp.w;
          *reinterpret_cast<uint4*>(smem + SM::OFF_Q16 + nosw_off<2>(s_loc * KP + f_loc, h)) = qp;
          *reinterpret_cast<uint4*>(smem + SM::OFF_DO16 + nosw_off<2>(s_loc * KP + f_loc, h)) = pack8_bf16(dT + h * DH);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) q16[i] = 0u;
      }
      {
        float t8[8];
#pragma unroll
        for (int h = 0; h < H; ++h) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            t8[e] = (use_res && __uint_as_float(zvr[U + h * DH + e]) > 0.f) ? dT[h * DH + e] : 0.f;
          *reinterpret_cast<uint4*>(smem + SM::OFF_DZ + nosw_off<12>(row, 6 + h)) = pack8_bf16(t8);
#pragma unroll
          for (int e = 0; e < 8; ++e) t8[e] = g[h * DH + e] * xhat[h * DH + e];
          *reinterpret_cast<uint4*>(smem + SM::OFF_DZ + nosw_off<12>(row, 8 + h)) = pack8_bf16(t8);
          *reinterpret_cast<uint4*>(smem + SM::OFF_DZ + nosw_off<12>(row, 10 + h)) = pack8_bf16(g + h * DH);
        }
      }
#pragma unroll
      for (int e = 0; e < DH; ++e) { q1[e] = qk[DH + e]; dO1[e] = dT[DH + e]; }
    }
    // ---- prefetch the next step's rows: a whole step of latency cover
    uint32_t xs_n[U], g_n[U / 2];
    float a_n[U];
    float2 lse_n = make_float2(INFINITY, INFINITY);
    if (!last) { load_xsrc(ntile, nit, xs_n); lse_n = load_lse(ntile, nit); }
    if (!last && new_tile) load_tile_head(ntile, a_n, g_n);
    // the layer's first iteration ends the tile: rows to add to dx (the MLP tower's input gradient) and the
    // lookup's slot at its owner, requested a whole step before they are used
    uint32_t add_raw[U / 2];
    int32_t slot = -1;
    if (it == 0 && active) {
      if (dx_add) {
        const T* ap = dx_add + smp * dx_bs + (int64_t)f_loc * dx_ld;
#pragma unroll
        for (int c = 0; c < U; c += 4) {
          const uint2 t2 = ldg_nc_u2(reinterpret_cast<const uint2*>(ap + c));
          add_raw[c / 2] = t2.x; add_raw[c / 2 + 1] = t2.y;
        }
      }
      if (sc.inverse) slot = sc.inverse[smp * F + f_loc];
    }

    // ================= per head: T2 (P, dS) ; the transposed / plain gradient products ; T3 (dq dk dv -> dZ)
    auto softmax_bwd = [&](int h) {
      float s[FP], dp[FP];
      {
        uint32_t t[FP], u[FP];
#pragma unroll
        for (int c0 = 0; c0 < FP; c0 += 8) {
          uint32_t t8[8], u8[8];
          tc_ld_32x8(tl + C_S + c0, t8);
          tc_ld_32x8(tl + C_DP + c0, u8);
#pragma unroll
          for (int j = 0; j < 8; ++j) { t[c0 + j] = t8[j]; u[c0 + j] = u8[j]; }
        }
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < FP; ++j) { s[j] = __uint_as_float(t[j]); dp[j] = __uint_as_float(u[j]); }
      }
      const float nl = -(h == 0 ? lse.x : lse.y);
      const float2 c2 = make_float2(scale_log2, scale_log2), nl2 = make_float2(nl, nl);
      float2 d2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) d2[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < FP; j += 2) {
        float2 e2 = ffma2(make_float2(s[j], s[j + 1]), c2, nl2);
        e2.x = ex2_approx(e2.x); e2.y = ex2_approx(e2.y);
        if (j >= FMIN) { e2.x = j < F ? e2.x : 0.f; e2.y = j + 1 < F ? e2.y : 0.f; }   // padded keys
        s[j] = e2.x; s[j + 1] = e2.y;                                   // P, normalised
        d2[(j >> 1) & 3] = ffma2(e2, make_float2(dp[j], dp[j + 1]), d2[(j >> 1) & 3]);
      }
      const float2 dt = fadd2(fadd2(d2[0], d2[1]), fadd2(d2[2], d2[3]));
      const float nd = -(dt.x + dt.y);                                   // -sum_j P dP  (dP already carries 1/sqrt(dh))
      const float2 nd2 = make_float2(nd, nd);
#pragma unroll
      for (int j = 0; j < FP; j += 2) {
        const float2 ds2 = fmul2(make_float2(s[j], s[j + 1]), fadd2(make_float2(dp[j], dp[j + 1]), nd2));
        dp[j] = ds2.x; dp[j + 1] = ds2.y;                               // dS
      }
#pragma unroll
      for (int c = 0; c < NCHF; ++c) {
        *reinterpret_cast<uint4*>(smem + SM::OFF_PC + nosw_off<NCHF>(row, c)) = pack8_bf16(s + c * 8);
        *reinterpret_cast<uint4*>(smem + SM::OFF_DS + nosw_off<NCHF>(row, c)) = pack8_bf16(dp + c * 8);
      }
    };
    auto issue_grads = [&](int h) {                 // dQ_h ; dV_h, dK_h per sample
#pragma unroll
      for (int ks = 0; ks < KP / 16; ++ks)
        tc_mma_bf16(tmem + C_A, mk_desc(b16, SM::OFF_DS + ks * 256, 128, NCHF * 128),
                    mk_desc(b16, SM::OFF_K16 + h * G::VX_BYTES + ks * 1024, 512, 128), ID_DQ, ks ? 1u : 0u);
#pragma unroll
      for (int s = 0; s < SPT; ++s)
#pragma unroll
        for (int ks = 0; ks < KP / 16; ++ks) {
          const uint32_t arow = (uint32_t)((s * NCHF + 2 * ks) * (NCHF * 128));
          const uint32_t brow = (uint32_t)((s * (KP / 8) + 2 * ks) * 256 + h * 128);
          tc_mma_bf16(tmem + C_A + 32 + s * 8, mk_desc(b16, SM::OFF_PC + arow, NCHF * 128, 128),
                      mk_desc(b16, SM::OFF_DO16 + brow, 256, 128), ID_T, ks ? 1u : 0u);
          tc_mma_bf16(tmem + C_A + 32 + s * 8 + (16u << 16), mk_desc(b16, SM::OFF_DS + arow, NCHF * 128, 128),
                      mk_desc(b16, SM::OFF_Q16 + brow, 256, 128), ID_T, ks ? 1u : 0u);
        }
    };
    auto grads_to_dz = [&](int h) {
      // own row: dq (the 8 columns of the row's sample), ReLU' from the packed q
      {
        uint32_t dq[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dq[e] = 0u;
        for (uint32_t s = ws_lo; s <= ws_hi; ++s) {
          uint32_t t8[8];
          tc_ld_32x8(tl + C_A + s * 8, t8);
          tc_wait_ld();
          if ((int)s == s_loc) {
#pragma unroll
            for (int e = 0; e < 8; ++e) dq[e] = t8[e];
          }
        }
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(dq[0]), __uint_as_float(dq[1])) & bf16x2_pos_mask(q16[h * 4 + 0]);
        v.y = pack_bf16x2(__uint_as_float(dq[2]), __uint_as_float(dq[3])) & bf16x2_pos_mask(q16[h * 4 + 1]);
        v.z = pack_bf16x2(__uint_as_float(dq[4]), __uint_as_float(dq[5])) & bf16x2_pos_mask(q16[h * 4 + 2]);
        v.w = pack_bf16x2(__uint_as_float(dq[6]), __uint_as_float(dq[7])) & bf16x2_pos_mask(q16[h * 4 + 3]);
        *reinterpret_cast<uint4*>(smem + SM::OFF_DZ + nosw_off<12>(row, h)) = v;
      }
      // dV (lanes 0-15) / dK (lanes 16-31) of key jw for every sample; ReLU' from the staged v / k rows
      if (wq * 16 < FP) {                             // uniform
        const uint8_t* msrc = smem + (is_dk ? SM::OFF_KX : SM::OFF_VX) + h * G::KX_BYTES;
        const int jr = jw < FP ? jw : 0;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
          uint32_t t8[8];
          tc_ld_32x8(tl + C_A + 32 + s * 8, t8);
          const uint4 m0 = *reinterpret_cast<const uint4*>(msrc + nosw_off<NCHK>(jr, 2 * s));
          const uint4 m1 = *reinterpret_cast<const uint4*>(msrc + nosw_off<NCHK>(jr, 2 * s + 1));
          tc_wait_ld();
          const uint32_t mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};   // relu outputs: > 0 <=> bits != 0
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = mm[e] ? __uint_as_float(t8[e]) : 0.f;
          if (jw < FP)
            *reinterpret_cast<uint4*>(smem + SM::OFF_DZ + nosw_off<12>(s * FP + jw, (is_dk ? 2 : 4) + h)) = pack8_bf16(t);
        }
      }
    };

    mbar_wait(bar, phase); phase ^= 1u;             // S_0, dP_0
    tc_fence_after();
    softmax_bwd(0);
    for (uint32_t s = ws_lo; s <= ws_hi; ++s) {     // head 1's expanded q / dO (head 0's products are complete)
      uint32_t v8[8], w8[8];
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        v8[e] = (int)s == s_loc ? __float_as_uint(q1[e]) : 0u;
        w8[e] = (int)s == s_loc ? __float_as_uint(dO1[e]) : 0u;
      }
      tc_st_32x8(tl + C_QX + s * 8, v8);
      tc_st_32x8(tl + C_DOX + s * 8, w8);
    }
    fence_async_smem();
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (wq == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < SPT; ++ks) {            // S_1, dP_1 first: the threads need them next
        tc_mma_tf32_ts(tmem + C_S, tmem + C_QX + ks * 8, mk_desc(b16, SM::OFF_KX + G::KX_BYTES + ks * 256, 128, NCHK * 128), ID_S, ks ? 1u : 0u);
        tc_mma_tf32_ts(tmem + C_DP, tmem + C_DOX + ks * 8, mk_desc(b16, SM::OFF_VX + G::KX_BYTES + ks * 256, 128, NCHK * 128), ID_S, ks ? 1u : 0u);
      }
      issue_grads(0);
      tc_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;             // S_1, dP_1 and head 0's gradient products (P / dS tiles free again)
    tc_fence_after();
    grads_to_dz(0);
    softmax_bwd(1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (wq == 0 && elect_one()) {
      tc_fence_after();
      issue_grads(1);
      tc_commit(bar);
    }
    // under head 1's gradient products: the NEXT step's X operand (S / dP columns are dead now)
    uint4x2_t xb_n;
    if (!last) xb_n = stage_x(xs_n, nit, nactive);
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    grads_to_dz(1);
    // ================= dX = dZ W^T ; [dW^T | db ; dgamma ; dbeta] += [dZ | g*xhat | g]^T [X | 1] ; Z(next)
    fence_async_smem();
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (wq == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc_mma_bf16(tmem + C_DX, mk_desc(b16, SM::OFF_DZ + ks * 256, 128, 1536),
                    mk_desc(b16, SM::OFF_WT + ks * 256, 128, 1024), ID_DX, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        tc_mma_bf16(tmem + C_DW, mk_desc(b16, SM::OFF_DZ + ks * 3072, 1536, 128),
                    mk_desc(b16, SM::OFF_XB + ks * 1024, 512, 128), ID_DW, ks ? 1u : dw_acc);
      if (!last) {
        issue_proj_3xtf32_ts(tmem + C_A, tmem + C_X, sbase + SM::OFF_W32, ID_Z);
        tc_mma_tf32_ts(tmem + C_A, tmem + C_X + 32, make_nosw_desc(sbase + SM::OFF_BT, 128, 256), ID_Z, 1u);
      }
      tc_commit(bar);
    }
    dw_acc = 1u;
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    {
      uint32_t dxr[16];
      tc_ld_32x16(tl + C_DX, dxr);
      tc_wait_ld();
      if (it > 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = __uint_as_float(dxr[u]);      // stays fp32 between iterations
      } else {
        if (active) {
          uint32_t pk[U / 2];
#pragma unroll
          for (int u = 0; u < U; u += 2) {
            float v0 = __uint_as_float(dxr[u]), v1 = __uint_as_float(dxr[u + 1]);
            if (dx_add) {
              const float2 a2 = unpack_bf16x2(add_raw[u / 2]);
              v0 += a2.x; v1 += a2.y;
            }
            pk[u / 2] = pack_bf16x2(v0, v1);
          }
          if (sc.inverse) {
            if (slot >= 0) {
              const int owner = slot / sc.cap, k = slot - owner * sc.cap;
              uint4* dst = reinterpret_cast<uint4*>(sc.recv[owner]) + ((int64_t)sc.rank * sc.cap + k) * 2;
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          } else {
            T* dp_ = dx + smp * dx_bs + (int64_t)f_loc * dx_ld;
#pragma unroll
            for (int u = 0; u < U / 2; u += 2) *reinterpret_cast<uint2*>(dp_ + 2 * u) = make_uint2(pk[u], pk[u + 1]);
          }
        }
        if (!last) {
#pragma unroll
          for (int u = 0; u < U; u += 2) {
            const float2 f2 = unpack_bf16x2(g_n[u / 2]);
            g[u] = f2.x; g[u + 1] = f2.y;
          }
        }
      }
      if (!last) {
        // dW of this step has consumed XB: the next step's bf16 x row may land now
        *reinterpret_cast<uint4*>(smem + SM::OFF_XB + nosw_off<4>(row, 0)) = xb_n.a;
        *reinterpret_cast<uint4*>(smem + SM::OFF_XB + nosw_off<4>(row, 1)) = xb_n.b;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          a[u] = new_tile ? a_n[u] : __uint_as_float(xs[u]);      // same tile: a of iteration it-1 = this step's x-source
          xs[u] = xs_n[u];
        }
        lse = lse_n;
      }
    }
    tile = ntile; it = nit; active = nactive;
  }
  // ---- per-CTA partials in the layout of the FFMA kernel: dW[D][4U] | db[4U] | dgamma[U] | dbeta[U]
  {
    uint32_t acc[32];
    tc_ld_32x32(tl + C_DW, acc);
    if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<int*>(part)[-4] = (int)gridDim.x;   // header: partial blocks written
    float* mine = part + (int64_t)blockIdx.x * (D * N4 + 6 * U);
    if (row < N4) {
#pragma unroll
      for (int d = 0; d < D; ++d) mine[d * N4 + row] = dw_acc ? __uint_as_float(acc[d]) : 0.f;
    }
    if (row < 6 * U) mine[D * N4 + row] = dw_acc ? __uint_as_float(acc[16]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (wq == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}

// out[i] = sum over CTAs of part[cta][i]; one warp per output, fixed order (deterministic)
static __global__ void itb_reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                  int max_parts, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  int nparts = reinterpret_cast<const int*>(part)[-4];        // written by the backward that filled `part`
  nparts = nparts < 0 ? 0 : (nparts > max_parts ? max_parts : nparts);
  const float s = warp_ordered_sum(part + i, nparts, n);
  if ((threadIdx.x & 31) == 0) out[i] = s;
}

constexpr int ITB_NP = 16 * 64 + 6 * 16;     // floats of one CTA's [dW | db | dgamma | dbeta] partial

// CTAs of the backward (= partial blocks in the workspace): two per SM, at most one per tile of `spt` samples.
// (A grid trimmed to the number of rounds — 274 CTAs do 2731 tiles in the same 10 rounds as 296 — was measured on one
// and two GPUs: the side branches that moved into the freed slots slowed the backward by 3 - 10 %.)
static int itb_grid(int B, int spt) {
  const int ntiles = (B + spt - 1) / spt;
  const int grid = sm_count() * 2;
  return grid > ntiles ? ntiles : grid;
}

template <int NCHF, typename T>
static int launch_itc_bwd(const IBwdArgs& a) {
  auto kern = interacting_tc_bwd_kernel<NCHF, T>;
  constexpr int smem = ItbSmem<NCHF>::TOTAL;
  RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  using G = ItcGeom<NCHF>;
  const int grid = itb_grid(a.B, G::SPT);
  const int np = ITB_NP;
  if (a.saved == nullptr) {
    set_error("interacting_tc_bwd: the saved activations of the tensor-core forward are required");
    return RS_ERR_INVALID;
  }
  if (a.ws_bytes < 16 + (size_t)grid * np * sizeof(float)) {
    set_error("interacting_tc_bwd: workspace %zu < %zu", a.ws_bytes, 16 + (size_t)grid * np * sizeof(float));
    return RS_ERR_WORKSPACE;
  }
  float* part = (float*)a.ws + 4;              // 16-byte header: the number of partial blocks this launch writes
  ItbScatter sc{};
  if (a.scatter) {
    for (int r = 0; r < RS_MAX_PEERS; ++r) sc.recv[r] = r < a.scatter->world ? a.scatter->peer_recv[r] : nullptr;
    sc.inverse = a.scatter->inverse; sc.cap = a.scatter->cap; sc.rank = a.scatter->rank;
  }
  kern<<<grid, 128, smem, a.st>>>((const T*)a.x, a.x_ld, a.x_bs, (const float*)a.saved, a.W, a.b, a.gm, a.bt, a.eps,
                                  (const T*)a.dy, a.dy_ld, a.dy_bs, (T*)a.dx, a.dx_ld, a.dx_bs, part, a.B, a.F,
                                  a.L, a.use_res, (const T*)a.dx_add, sc);
  if (int e = check_launch("interacting_tc_bwd")) return e;
  if (a.dparams == nullptr) return 0;       // deferred: interacting_tc_bwd_reduce (rs_interacting_bwd_reduce)
  itb_reduce_partials_kernel<<<(np * 32 + 255) / 256, 256, 0, a.st>>>(part, a.dparams, grid, np);
  return check_launch("interacting_tc_bwd_reduce");
}

// Deferred reduction of the per-CTA partials (a.dparams == nullptr above), same geometry as the launch.
static int itb_spt(int F) {
  switch ((F + 7) / 8) {       // the dispatch of interacting_tc_bwd below
    case 1:
    case 2: return ItcGeom<2>::SPT;
    case 3: return ItcGeom<3>::SPT;
    case 4: return ItcGeom<4>::SPT;
    case 5: return ItcGeom<5>::SPT;
    default: return ItcGeom<6>::SPT;
  }
}

int interacting_tc_bwd_reduce(const void* ws, size_t ws_bytes, float* dparams, int B, int F, cudaStream_t st) {
  const int grid = itb_grid(B, itb_spt(F));       // upper bound; the launch recorded its own count in the header
  if (ws_bytes < 16 + (size_t)grid * ITB_NP * sizeof(float)) {
    set_error("interacting_tc_bwd_reduce: workspace %zu < %zu", ws_bytes, 16 + (size_t)grid * ITB_NP * sizeof(float));
    return RS_ERR_WORKSPACE;
  }
  itb_reduce_partials_kernel<<<(ITB_NP * 32 + 255) / 256, 256, 0, st>>>((const float*)ws + 4, dparams, grid, ITB_NP);
  return check_launch("interacting_tc_bwd_reduce");
}

int interacting_tc_bwd(const IBwdArgs& a) {
  switch ((a.F + 7) / 8) {
    case 1:
    case 2: return launch_itc_bwd<2, __nv_bfloat16>(a);
    case 3: return launch_itc_bwd<3, __nv_bfloat16>(a);
    case 4: return launch_itc_bwd<4, __nv_bfloat16>(a);
    case 5: return launch_itc_bwd<5, __nv_bfloat16>(a);
    default: return launch_itc_bwd<6, __nv_bfloat16>(a);
  }
}

}  // namespace rs
