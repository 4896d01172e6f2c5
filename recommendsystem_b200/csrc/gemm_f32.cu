// gemm_f32.cu — fp32 FFMA GEMM with fused epilogues (parity mode of K5).
// C[M,N] = epi(op(A)[M,K] · op(B)[K,N]).  64x64x16 tiles, 256 threads, 4x4
// register blocking.  The bf16 production path is the tcgen05 kernel in
// gemm_tc.cu; rs_gemm dispatches on dtype_ab.
#include "common.cuh"

namespace rs {

int gemm_bf16_tc(const void* A, int64_t lda, int transA, const void* B, int64_t ldb, int transB,
                 void* C, int64_t ldc, const float* bias, const void* aux, int64_t ldaux,
                 int epilogue, int M, int N, int K, int dtype_c, void* ws, size_t ws_bytes,
                 cudaStream_t st);
size_t gemm_tc_workspace_bytes();
bool gemm_tf32x3_usable(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int dtype_c,
                        int epilogue, size_t ws_bytes);
int gemm_tf32x3(const void* A, int64_t lda, int transA, const void* B, int64_t ldb, int transB, void* C,
                int64_t ldc, const float* bias, const void* aux, int64_t ldaux, int epilogue, int M, int N,
                int K, void* ws, size_t ws_bytes, cudaStream_t st);
std::atomic<int> g_fp32_gemm_mode{0};      // 0: 3xTF32 tensor cores when the problem qualifies, 1: FFMA only
int splitk_reduce(const float* partial, int splits, void* C, int64_t ldc, int M, int N, int accumulate,
                  int dtype_c, cudaStream_t st);

constexpr int TM = 64, TN = 64, TK = 16;

template <typename CT>
__device__ __forceinline__ float epi_apply(float acc, int epi, const float* bias, const CT* aux,
                                           int64_t ldaux, const CT* Cold, int64_t m, int n) {
  switch (epi) {
    case RS_EPI_BIAS: return acc + bias[n];
    case RS_EPI_BIAS_RELU: return fmaxf(acc + bias[n], 0.f);
    case RS_EPI_BIAS_SIGMOID: return 1.f / (1.f + expf(-(acc + bias[n])));
    case RS_EPI_MUL_RELU_MASK: return to_f<CT>(aux[m * ldaux + n]) > 0.f ? acc : 0.f;
    case RS_EPI_MUL_DSIGMOID: {
      const float r = to_f<CT>(aux[m * ldaux + n]);
      return acc * r * (1.f - r);
    }
    case RS_EPI_ACCUM: return acc + to_f<CT>(*Cold);
    default: return acc;
  }
}

template <typename CT>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int64_t lda, int transA, const float* __restrict__ B,
             int64_t ldb, int transB, CT* __restrict__ C, int64_t ldc,
             const float* __restrict__ bias, const CT* __restrict__ aux, int64_t ldaux, int epi,
             int M, int N, int K, int k_per_split, float* __restrict__ partial) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    // A tile: 64 x 16 = 1024 elements, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (transA) { mm = idx % TM; kk = idx / TM; }   // stored [K][M]: m fastest
      else { kk = idx % TK; mm = idx / TK; }           // stored [M][K]: k fastest
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < k_end) v = transA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int nn, kk;
      if (transB) { kk = idx % TK; nn = idx / TK; }    // stored [N][K]: k fastest
      else { nn = idx % TN; kk = idx / TN; }            // stored [K][N]: n fastest
      const int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < k_end) v = transB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (partial) {   // split-K: raw partial sums, reduced in split order afterwards
        partial[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
        continue;
      }
      CT* c = C + m * ldc + n;
      *c = from_f<CT>(epi_apply<CT>(acc[i][j], epi, bias, aux, ldaux, c, m, n));
    }
  }
}

}  // namespace rs

using namespace rs;

extern "C" size_t rs_gemm_workspace_bytes(void) { return gemm_tc_workspace_bytes(); }
extern "C" int rs_set_fp32_gemm_mode(int mode) {
  RS_REQUIRE(mode == 0 || mode == 1, "rs_set_fp32_gemm_mode: mode %d", mode);
  return g_fp32_gemm_mode.exchange(mode);
}

extern "C" int rs_gemm(const void* A, int64_t lda, int transA, const void* B, int64_t ldb,
                       int transB, void* C, int64_t ldc, const float* bias, const void* aux,
                       int64_t ldaux, int epilogue, int M, int N, int K, int dtype_ab, int dtype_c,
                       void* ws, size_t ws_bytes, void* stream) {
  RS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: M=%d N=%d K=%d", M, N, K);
  RS_REQUIRE(epilogue >= RS_EPI_NONE && epilogue <= RS_EPI_ACCUM, "gemm: epilogue %d", epilogue);
  if (epilogue >= RS_EPI_BIAS && epilogue <= RS_EPI_BIAS_SIGMOID)
    RS_REQUIRE(bias != nullptr, "gemm: bias epilogue without bias");
  if (epilogue == RS_EPI_MUL_RELU_MASK || epilogue == RS_EPI_MUL_DSIGMOID)
    RS_REQUIRE(aux != nullptr, "gemm: mask epilogue without aux");
  cudaStream_t st = as_stream(stream);
  if (dtype_ab == RS_BF16)
    return gemm_bf16_tc(A, lda, transA, B, ldb, transB, C, ldc, bias, aux, ldaux, epilogue, M, N, K,
                        dtype_c, ws, ws_bytes, st);
  RS_REQUIRE(dtype_ab == RS_F32, "gemm: bad operand dtype %d", dtype_ab);
  if (g_fp32_gemm_mode.load(std::memory_order_relaxed) == 0 && gemm_tf32x3_usable(A, lda, B, ldb, M, N, K, dtype_c, epilogue,
                                                                                 ws != nullptr ? ws_bytes : 0))
    return gemm_tf32x3(A, lda, transA, B, ldb, transB, C, ldc, bias, aux, ldaux, epilogue, M, N, K, ws, ws_bytes, st);
  // deep-K, small-output problems (weight gradients over the batch): split K across CTAs
  const int tiles = (int)(cdiv(N, TN) * cdiv(M, TM));
  int splits = 1;
  if ((epilogue == RS_EPI_NONE || epilogue == RS_EPI_ACCUM) && ws != nullptr && tiles * 2 <= sm_count() * 2 &&
      K >= 1024) {
    splits = (sm_count() * 2) / tiles;
    if (splits > K / 256) splits = K / 256;
    if (splits < 1) splits = 1;
    if ((size_t)splits * M * N * sizeof(float) > ws_bytes) splits = 1;
  }
  int kps = (int)cdiv(cdiv(K, splits), TK) * TK;
  splits = (int)cdiv(K, kps);
  float* partial = splits > 1 ? (float*)ws : nullptr;
  dim3 grid((unsigned)cdiv(N, TN), (unsigned)cdiv(M, TM), (unsigned)splits);
  if (dtype_c == RS_F32)
    sgemm_kernel<float><<<grid, 256, 0, st>>>((const float*)A, lda, transA, (const float*)B, ldb,
                                              transB, (float*)C, ldc, bias, (const float*)aux, ldaux,
                                              epilogue, M, N, K, kps, partial);
  else if (dtype_c == RS_BF16)
    sgemm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        (const float*)A, lda, transA, (const float*)B, ldb, transB, (__nv_bfloat16*)C, ldc, bias,
        (const __nv_bfloat16*)aux, ldaux, epilogue, M, N, K, kps, partial);
  else { set_error("gemm: bad C dtype"); return RS_ERR_INVALID; }
  if (int e = check_launch("sgemm")) return e;
  if (splits > 1) return splitk_reduce(partial, splits, C, ldc, M, N, epilogue == RS_EPI_ACCUM, dtype_c, st);
  return 0;
}
