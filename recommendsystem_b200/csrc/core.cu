// core.cu — error reporting, launch accounting and the small glue kernels
// (2-D copy/cast, add, activation backward, column sums, BCE loss).
#include "common.cuh"
#include <cuda.h>
#include <string.h>
#include <mutex>
#include <string.h>

namespace rs {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

template <typename S, typename D>
__global__ void copy2d_kernel(const S* __restrict__ src, int64_t lds, D* __restrict__ dst,
                              int64_t ldd, int M, int N) {
  const int64_t total = (int64_t)M * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t m = t / N;
    const int n = (int)(t % N);
    dst[m * ldd + n] = from_f<D>(to_f<S>(src[m * lds + n]));
  }
}

template <typename T>
__global__ void add2d_kernel(const T* __restrict__ a, int64_t lda, const T* __restrict__ b,
                             int64_t ldb, T* __restrict__ y, int64_t ldy, int M, int N) {
  const int64_t total = (int64_t)M * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t m = t / N;
    const int n = (int)(t % N);
    y[m * ldy + n] = from_f<T>(to_f<T>(a[m * lda + n]) + to_f<T>(b[m * ldb + n]));
  }
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ ref,
                               int64_t ldref, T* __restrict__ y, int64_t ldy, int M, int N,
                               int kind) {
  const int64_t total = (int64_t)M * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t m = t / N;
    const int n = (int)(t % N);
    const float g = to_f<T>(x[m * ldx + n]);
    const float r = to_f<T>(ref[m * ldref + n]);
    const float o = kind == 0 ? (r > 0.f ? g : 0.f) : g * r * (1.f - r);
    y[m * ldy + n] = from_f<T>(o);
  }
}

// Deterministic column sum in two launches (no atomics): block (bx, by) sums rows
// [by*CS_ROWS, (by+1)*CS_ROWS) of a 128-column strip; a thread owns 4 adjacent columns (one
// 16-/8-byte load per row), the 8 row-lanes of the block are combined in a fixed order, and
// a second launch adds the per-strip partials in row-block order.
constexpr int CS_ROWS = 128;
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ part,
                                      int M, int N, int vec_ok) {
  __shared__ float4 sm[8][32];
  const int col = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * CS_ROWS;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    const int r1 = min(M, r0 + CS_ROWS);
    if (vec_ok && col + 4 <= N && r1 - r0 == CS_ROWS) {
      // full strip: all 16 row loads of the thread are in flight before the first add (same summation order as the
      // loop below; the dependent-load version spent ~10 us on a 4 MB matrix)
      float4 v[CS_ROWS / 8];
#pragma unroll
      for (int k = 0; k < CS_ROWS / 8; ++k) v[k] = load4<T>(x + (int64_t)(r0 + threadIdx.y + 8 * k) * ldx + col);
#pragma unroll
      for (int k = 0; k < CS_ROWS / 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    } else if (vec_ok && col + 4 <= N) {
      for (int r = r0 + threadIdx.y; r < r1; r += 8) {
        const float4 v = load4<T>(x + (int64_t)r * ldx + col);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    } else {
      for (int r = r0 + threadIdx.y; r < r1; r += 8) {
        const T* p = x + (int64_t)r * ldx + col;
        acc.x += to_f<T>(p[0]);
        if (col + 1 < N) acc.y += to_f<T>(p[1]);
        if (col + 2 < N) acc.z += to_f<T>(p[2]);
        if (col + 3 < N) acc.w += to_f<T>(p[3]);
      }
    }
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float4 s = sm[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = sm[k][threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* dst = part + (int64_t)blockIdx.y * N + col;
    dst[0] = s.x;
    if (col + 1 < N) dst[1] = s.y;
    if (col + 2 < N) dst[2] = s.z;
    if (col + 3 < N) dst[3] = s.w;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts,
                                    int N) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per column
  if (col >= N) return;
  const float s = warp_ordered_sum(part + col, nparts, N);
  if ((threadIdx.x & 31) == 0) out[col] = s;
}

// BCE on clip(p_raw,1e-6,1) with the sigmoid' of the producing Dense folded in.
// One block, ordered tree reduce => deterministic.  B*k is small (<= a few 10^5).
template <typename T>
__global__ void bce_kernel(const T* __restrict__ p_raw, const float* __restrict__ y, float a,
                           float* __restrict__ loss_out, T* __restrict__ dz, int B, int k) {
  __shared__ float red[1024];
  const int total = B * k;
  const float invB = 1.f / (float)B;
  float acc = 0.f;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const float pr = to_f<T>(p_raw[t]);
    const float p = fminf(fmaxf(pr, 1e-6f), 1.0f);
    const float yy = y[t];
    acc += -yy * logf(p + 1e-6f) - (a - yy) * logf(1.0f - p + 1e-6f);
    if (dz) {
      const float dp = (-yy / (p + 1e-6f) + (a - yy) / (1.0f - p + 1e-6f)) * invB;
      const float pass = (pr >= 1e-6f && pr <= 1.0f) ? 1.f : 0.f;
      dz[t] = from_f<T>(dp * pass * pr * (1.f - pr));
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = red[0] * invB;
}

// dst[n][m] = src[m][n]; 32x32 tiles through padded shared memory so that both the
// global read (along n) and the global write (along m) are coalesced.
template <typename E>
__global__ void transpose2d_kernel(const E* __restrict__ src, int64_t lds, E* __restrict__ dst,
                                   int64_t ldd, int M, int N) {
  __shared__ E tile[32][33];
  const int n0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int m = m0 + r, n = n0 + threadIdx.x;
    if (m < M && n < N) tile[r][threadIdx.x] = src[(int64_t)m * lds + n];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, m = m0 + threadIdx.x;
    if (m < M && n < N) dst[(int64_t)n * ldd + m] = tile[threadIdx.x][r];
  }
}

}  // namespace rs

using namespace rs;

extern "C" {

int rs_abi_version(void) { return RS_ABI_VERSION; }
const char* rs_last_error(void) { return g_err; }
uint64_t rs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int rs_built_for_sm100a(void) { return 1; }

static __global__ void debug_timestamp_kernel(unsigned long long* dst) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *dst = t;
}

int rs_debug_timestamp(unsigned long long* dst, void* stream) {
  RS_REQUIRE(dst != nullptr, "debug_timestamp: dst is NULL");
  debug_timestamp_kernel<<<1, 1, 0, as_stream(stream)>>>(dst);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("debug_timestamp: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

static inline int grid_for(int64_t total, int threads) {
  int64_t b = cdiv(total, threads);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int rs_copy2d(const void* src, int64_t lds, int sdt, void* dst, int64_t ldd, int ddt, int M, int N,
              void* stream) {
  if (M <= 0 || N <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for((int64_t)M * N, 256);
  if (sdt == RS_F32 && ddt == RS_F32)
    copy2d_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, M, N);
  else if (sdt == RS_F32 && ddt == RS_BF16)
    copy2d_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, lds, (__nv_bfloat16*)dst, ldd, M, N);
  else if (sdt == RS_BF16 && ddt == RS_F32)
    copy2d_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, lds, (float*)dst, ldd, M, N);
  else if (sdt == RS_BF16 && ddt == RS_BF16)
    copy2d_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, M, N);
  else { set_error("copy2d: bad dtypes"); return RS_ERR_INVALID; }
  return check_launch("copy2d");
}

int rs_add2d(const void* a, int64_t lda, const void* b, int64_t ldb, void* y, int64_t ldy, int M,
             int N, int dtype, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for((int64_t)M * N, 256);
  if (dtype == RS_F32)
    add2d_kernel<float><<<g, 256, 0, st>>>((const float*)a, lda, (const float*)b, ldb, (float*)y, ldy, M, N);
  else if (dtype == RS_BF16)
    add2d_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb, (__nv_bfloat16*)y, ldy, M, N);
  else { set_error("add2d: bad dtype"); return RS_ERR_INVALID; }
  return check_launch("add2d");
}

int rs_act_bwd(const void* x, int64_t ldx, const void* ref, int64_t ldref, void* y, int64_t ldy,
               int M, int N, int dtype, int kind, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  RS_REQUIRE(kind == 0 || kind == 1, "act_bwd: kind=%d", kind);
  cudaStream_t st = as_stream(stream);
  const int g = grid_for((int64_t)M * N, 256);
  if (dtype == RS_F32)
    act_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)x, ldx, (const float*)ref, ldref, (float*)y, ldy, M, N, kind);
  else if (dtype == RS_BF16)
    act_bwd_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)ref, ldref, (__nv_bfloat16*)y, ldy, M, N, kind);
  else { set_error("act_bwd: bad dtype"); return RS_ERR_INVALID; }
  return check_launch("act_bwd");
}

size_t rs_colsum_workspace_bytes(int M, int N) {
  return (size_t)cdiv(M > 0 ? M : 1, CS_ROWS) * (size_t)(N > 0 ? N : 1) * sizeof(float);
}

int rs_colsum(const void* x, int64_t ldx, int dtype, float* out, int M, int N, void* ws,
              size_t ws_bytes, void* stream) {
  if (N <= 0) return 0;
  RS_REQUIRE(M > 0, "colsum: M=%d", M);
  if (ws_bytes < rs_colsum_workspace_bytes(M, N)) { set_error("colsum: workspace too small"); return RS_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int nparts = (int)cdiv(M, CS_ROWS);
  dim3 grid((unsigned)cdiv(N, 128), (unsigned)nparts), block(32, 8);
  const int esz = dtype == RS_F32 ? 4 : 2;
  const int vec_ok = ((uintptr_t)x % (4 * esz) == 0) && ((ldx * esz) % (4 * esz) == 0);
  if (dtype == RS_F32)
    colsum_partial_kernel<float><<<grid, block, 0, st>>>((const float*)x, ldx, (float*)ws, M, N, vec_ok);
  else if (dtype == RS_BF16)
    colsum_partial_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, ldx, (float*)ws, M, N, vec_ok);
  else { set_error("colsum: bad dtype"); return RS_ERR_INVALID; }
  if (int e = check_launch("colsum_partial")) return e;
  colsum_final_kernel<<<(unsigned)cdiv((int64_t)N * 32, 256), 256, 0, st>>>((const float*)ws, out, nparts, N);
  return check_launch("colsum_final");
}

int rs_transpose2d(const void* src, int64_t lds, void* dst, int64_t ldd, int M, int N, int dtype,
                   void* stream) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(M, 32)), block(32, 8);
  if (dtype == RS_F32)
    transpose2d_kernel<float><<<grid, block, 0, as_stream(stream)>>>((const float*)src, lds, (float*)dst, ldd, M, N);
  else if (dtype == RS_BF16)
    transpose2d_kernel<uint16_t><<<grid, block, 0, as_stream(stream)>>>((const uint16_t*)src, lds, (uint16_t*)dst, ldd, M, N);
  else { set_error("transpose2d: bad dtype"); return RS_ERR_INVALID; }
  return check_launch("transpose2d");
}

int rs_bce_sigmoid_fwd_bwd(const void* p_raw, int dtype, const float* y, float a, float* loss_out,
                           void* dz, int B, int k, void* stream) {
  RS_REQUIRE(B > 0 && k > 0, "bce: B=%d k=%d", B, k);
  cudaStream_t st = as_stream(stream);
  if (dtype == RS_F32)
    bce_kernel<float><<<1, 1024, 0, st>>>((const float*)p_raw, y, a, loss_out, (float*)dz, B, k);
  else if (dtype == RS_BF16)
    bce_kernel<__nv_bfloat16><<<1, 1024, 0, st>>>((const __nv_bfloat16*)p_raw, y, a, loss_out, (__nv_bfloat16*)dz, B, k);
  else { set_error("bce: bad dtype"); return RS_ERR_INVALID; }
  return check_launch("bce");
}

// ---- CUDA IPC: peer mappings of row-sharded embedding tables (rs_embed_gather_peer_fwd) ----
int rs_ipc_export(const void* ptr, unsigned char* handle64, unsigned long long* offset) {
  RS_REQUIRE(ptr && handle64 && offset, "ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = nullptr;
  if (!range_fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      set_error("ipc_export: cuMemGetAddressRange unavailable");
      return RS_ERR_UNSUPPORTED;
    }
    range_fn = (RangeFn)p;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range_fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) {
    set_error("ipc_export: %p is not inside a device allocation", ptr);
    return RS_ERR_INVALID;
  }
  cudaIpcMemHandle_t h;
  RS_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  memcpy(handle64, &h, 64);
  *offset = (unsigned long long)((CUdeviceptr)ptr - base);
  return 0;
}

int rs_ipc_import(const unsigned char* handle64, unsigned long long offset, void** mapped) {
  RS_REQUIRE(handle64 && mapped, "ipc_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  RS_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *mapped = (char*)base + offset;
  return 0;
}

}  // extern "C"
