// probe_tc.cu — hardware probes for the tcgen05 mechanisms the round-2 InteractingLayer kernels rely on
// (developer tool, not product code).  Each probe runs one CTA of 128 threads against small-integer data
// (every product exact) and compares with a host loop.
//   1  kind::tf32, A from TMEM (tcgen05.st by the row-owning thread), B K-major no-swizzle, M=128 N=48 K=24
//   2  kind::f16 (bf16), A from TMEM (2 per column), B MN-major no-swizzle, M=128 N=32 K=48
//   3  kind::f16, M=64, A = compact [rows][48] tile read MN-major (transposed), B MN-major, N=8, K=48 rows,
//      D at lane offsets 0 and 16 (interleaved half-subpartitions); prints the lane mapping it finds
//   4  as 1 with N=40 (is N%8 legal at M=128?)  -- run alone: an illegal shape traps
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/probe_tc tools/probe_tc.cu
#include "../recommendsystem_b200/csrc/tc_common.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>
using namespace rs;

__device__ __forceinline__ void tc_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(int mode, const float* Aq, const float* Kx, const float* Ap, const float* Vx, const float* Pc,
             const float* dO, const float* Qm, float* D1, float* D2, float* D3) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  constexpr int OFF_KX = 0, OFF_VX = 8192, OFF_PC = 16384, OFF_DO = 32768, OFF_QM = 40960, OFF_BAR = 49152;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 49152 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // Kx [48 n][24 k] tf32, K-major no-swizzle, 6 chunks per row
  for (int i = tid; i < 48 * 6; i += 128) {
    const int n = i / 6, c = i % 6;
    *reinterpret_cast<float4*>(smem + OFF_KX + nosw_off<6>(n, c)) =
        make_float4(Kx[n * 24 + c * 4], Kx[n * 24 + c * 4 + 1], Kx[n * 24 + c * 4 + 2], Kx[n * 24 + c * 4 + 3]);
  }
  // Vx [48 k][32 n] bf16, MN-major: row = k, 4 chunks of 8 n
  for (int i = tid; i < 48 * 4; i += 128) {
    const int k = i / 4, c = i % 4;
    uint4 v;
    const float* s = Vx + k * 32 + c * 8;
    v.x = pack_bf16x2(s[0], s[1]); v.y = pack_bf16x2(s[2], s[3]); v.z = pack_bf16x2(s[4], s[5]); v.w = pack_bf16x2(s[6], s[7]);
    *reinterpret_cast<uint4*>(smem + OFF_VX + nosw_off<4>(k, c)) = v;
  }
  // Pc [128 i][48 j] bf16, 6 chunks per row (rows 128..135 stay zero)
  for (int i = tid; i < 128 * 6; i += 128) {
    const int r = i / 6, c = i % 6;
    uint4 v;
    const float* s = Pc + r * 48 + c * 8;
    v.x = pack_bf16x2(s[0], s[1]); v.y = pack_bf16x2(s[2], s[3]); v.z = pack_bf16x2(s[4], s[5]); v.w = pack_bf16x2(s[6], s[7]);
    *reinterpret_cast<uint4*>(smem + OFF_PC + nosw_off<6>(r, c)) = v;
  }
  // dO16x / Q16x: [3 samples x 48 rows][16] bf16, rows 40..47 of each sample zero; tile row r = 40 s + f
  for (int i = tid; i < 120 * 2; i += 128) {
    const int r = i / 2, c = i % 2, s_ = r / 40, f = r % 40;
    uint4 v, w;
    const float* a = dO + r * 16 + c * 8;
    const float* b = Qm + r * 16 + c * 8;
    v.x = pack_bf16x2(a[0], a[1]); v.y = pack_bf16x2(a[2], a[3]); v.z = pack_bf16x2(a[4], a[5]); v.w = pack_bf16x2(a[6], a[7]);
    w.x = pack_bf16x2(b[0], b[1]); w.y = pack_bf16x2(b[2], b[3]); w.z = pack_bf16x2(b[4], b[5]); w.w = pack_bf16x2(b[6], b[7]);
    *reinterpret_cast<uint4*>(smem + OFF_DO + nosw_off<2>(s_ * 48 + f, c)) = v;
    *reinterpret_cast<uint4*>(smem + OFF_QM + nosw_off<2>(s_ * 48 + f, c)) = w;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t b16 = smem_u32(smem) >> 4;
  // ---- stage the TMEM A operands: thread = lane = row
  {
    uint32_t r8[8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {                      // Aq: 24 tf32 columns at 0
#pragma unroll
      for (int e = 0; e < 8; ++e) r8[e] = __float_as_uint(Aq[tid * 24 + c * 8 + e]);
      tc_st_32x8(tl + 0 + c * 8, r8);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {                      // Ap: 48 bf16 = 24 columns at 128
#pragma unroll
      for (int e = 0; e < 8; ++e) r8[e] = pack_bf16x2(Ap[tid * 48 + c * 16 + 2 * e], Ap[tid * 48 + c * 16 + 2 * e + 1]);
      tc_st_32x8(tl + 128 + c * 8, r8);
    }
    // clear the D3 region so that untouched lanes read back a marker
#pragma unroll
    for (int e = 0; e < 8; ++e) r8[e] = __float_as_uint(-777.f);
#pragma unroll
    for (int c = 0; c < 6; ++c) tc_st_32x8(tl + 200 + c * 8, r8);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    if (mode == 1 || mode == 4) {
      const int N = mode == 4 ? 40 : 48;
      const uint32_t id = make_idesc(2, 128, N, 0, 0);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
        tc_mma_tf32_ts(tmem + 64, tmem + 0 + ks * 8, mk_desc(b16, OFF_KX + ks * 256, 128, 768), id, ks ? 1u : 0u);
    }
    if (mode == 2) {
      const uint32_t id = make_idesc(1, 128, 32, 0, 1);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
        tc_mma_bf16_ts(tmem + 160, tmem + 128 + ks * 8, mk_desc(b16, OFF_VX + ks * 1024, 512, 128), id, ks ? 1u : 0u);
    }
    if (mode == 3) {
      const uint32_t id = make_idesc(1, 64, 8, 1, 1);
      for (int s = 0; s < 3; ++s)
        for (int h = 0; h < 2; ++h) {
          const uint32_t dcol = 200 + (s * 2 + h) * 8;
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            // A: Pc rows 40 s + 16 ks.. (K), all 64 (48 real) j's (M); B: expanded rows 48 s + 16 ks, head chunk h
            tc_mma_bf16(tmem + dcol, mk_desc(b16, OFF_PC + (5 * s + 2 * ks) * 768, 768, 128),
                        mk_desc(b16, OFF_DO + (6 * s + 2 * ks) * 256 + h * 128, 256, 128), id, ks ? 1u : 0u);
            tc_mma_bf16(tmem + dcol + (16u << 16), mk_desc(b16, OFF_PC + (5 * s + 2 * ks) * 768, 768, 128),
                        mk_desc(b16, OFF_QM + (6 * s + 2 * ks) * 256 + h * 128, 256, 128), id, ks ? 1u : 0u);
          }
        }
    }
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  if (mode == 1 || mode == 4) {
    for (int c = 0; c < 48; c += 8) {
      uint32_t t8[8];
      tc_ld_32x8(tl + 64 + c, t8);
      tc_wait_ld();
      for (int e = 0; e < 8; ++e) D1[tid * 48 + c + e] = __uint_as_float(t8[e]);
    }
  }
  if (mode == 2) {
    for (int c = 0; c < 32; c += 8) {
      uint32_t t8[8];
      tc_ld_32x8(tl + 160 + c, t8);
      tc_wait_ld();
      for (int e = 0; e < 8; ++e) D2[tid * 32 + c + e] = __uint_as_float(t8[e]);
    }
  }
  if (mode == 3) {
    for (int c = 0; c < 48; c += 8) {
      uint32_t t8[8];
      tc_ld_32x8(tl + 200 + c, t8);
      tc_wait_ld();
      for (int e = 0; e < 8; ++e) D3[tid * 48 + c + e] = __uint_as_float(t8[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

static float rnd_int() { return (float)((rand() % 5) - 2); }

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  srand(1234);
  std::vector<float> Aq(128 * 24), Kx(48 * 24), Ap(128 * 48), Vx(48 * 32), Pc(128 * 48), dO(128 * 16), Qm(128 * 16);
  for (auto* v : {&Aq, &Kx, &Ap, &Vx, &Pc, &dO, &Qm})
    for (auto& x : *v) x = rnd_int();
  float *dAq, *dKx, *dAp, *dVx, *dPc, *ddO, *dQm, *dD1, *dD2, *dD3;
  auto up = [](float** d, const std::vector<float>& h) {
    cudaMalloc(d, h.size() * 4);
    cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  };
  up(&dAq, Aq); up(&dKx, Kx); up(&dAp, Ap); up(&dVx, Vx); up(&dPc, Pc); up(&ddO, dO); up(&dQm, Qm);
  cudaMalloc(&dD1, 128 * 48 * 4); cudaMalloc(&dD2, 128 * 32 * 4); cudaMalloc(&dD3, 128 * 48 * 4);
  cudaMemset(dD1, 0, 128 * 48 * 4); cudaMemset(dD2, 0, 128 * 32 * 4); cudaMemset(dD3, 0, 128 * 48 * 4);
  const int smem = 49152 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(mode, dAq, dKx, dAp, dVx, dPc, ddO, dQm, dD1, dD2, dD3);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d: %s\n", mode, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  if (mode == 1 || mode == 4) {
    const int N = mode == 4 ? 40 : 48;
    std::vector<float> D(128 * 48);
    cudaMemcpy(D.data(), dD1, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 24; ++k) ref += Aq[r * 24 + k] * Kx[n * 24 + k];
        if (D[r * 48 + n] != ref && bad++ < 5) printf("  D1[%d][%d] = %g, ref %g\n", r, n, D[r * 48 + n], ref);
      }
    printf("probe %d (tf32 A-from-TMEM, N=%d): %d mismatches\n", mode, N, bad);
  }
  if (mode == 2) {
    std::vector<float> D(128 * 32);
    cudaMemcpy(D.data(), dD2, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 32; ++n) {
        float ref = 0;
        for (int k = 0; k < 48; ++k) ref += Ap[r * 48 + k] * Vx[k * 32 + n];
        if (D[r * 32 + n] != ref && bad++ < 5) printf("  D2[%d][%d] = %g, ref %g\n", r, n, D[r * 32 + n], ref);
      }
    printf("probe 2 (bf16 A-from-TMEM, B MN-major): %d mismatches\n", bad);
  }
  if (mode == 3) {
    std::vector<float> D(128 * 48);
    cudaMemcpy(D.data(), dD3, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad16 = 0;
    for (int s = 0; s < 3; ++s)
      for (int h = 0; h < 2; ++h)
        for (int j = 0; j < 48; ++j)
          for (int e = 0; e < 8; ++e) {
            float ref = 0, refq = 0;
            for (int f = 0; f < 40; ++f) {
              ref += Pc[(40 * s + f) * 48 + j] * dO[(40 * s + f) * 16 + 8 * h + e];
              refq += Pc[(40 * s + f) * 48 + j] * Qm[(40 * s + f) * 16 + 8 * h + e];
            }
            const int lane = (j % 16) + 32 * (j / 16), col = (s * 2 + h) * 8 + e;
            if (D[lane * 48 + col] != ref && bad++ < 5) printf("  D3[lane %d][%d] = %g, ref %g\n", lane, col, D[lane * 48 + col], ref);
            if (D[(lane + 16) * 48 + col] != refq && bad16++ < 5)
              printf("  D3[lane %d][%d] = %g, ref %g\n", lane + 16, col, D[(lane + 16) * 48 + col], refq);
          }
    printf("probe 3 (M=64 transposed compact tile): lanes+0 %d mismatches, lanes+16 %d mismatches\n", bad, bad16);
    int marker = 0;
    for (int l = 96; l < 128; ++l)
      for (int c = 0; c < 48; ++c) marker += D[l * 48 + c] == -777.f;
    printf("  warp 3 lanes still hold the marker in %d of %d cells\n", marker, 32 * 48);
  }
  return 0;
}
