// labels.cu — the steps either side of the train step:
//   * staytime label transform of one batch (staytime/parse.py:30-68): Gaussian-smoothed 400-bin
//     distribution + capped watch time, short / long play labels, sample weights;
//   * streaming binary-classification metrics (rough_rank/model.py:215-219, staytime/model.py:78-82):
//     Keras AUC (200 thresholds, ROC, interpolation), BinaryAccuracy, CTR, COPC.
// Both are HBM-bound element-wise / histogram work: coalesced 16-byte accesses, shared-memory
// histograms, integer counts (exact, order independent) and ordered partial sums (deterministic).
#include "common.cuh"

namespace rs {

// ------------------------------------------------------------------ staytime labels
struct StayArgs {
  const int64_t* watch_ms; const uint8_t* landing; const float* bins; int nbins; int B;
  float* label; int64_t* short_label; int64_t* long_label; float* weight;
  int64_t short_ms, long_ms; float cap_s, neg_two_sigma2, div_num, width, landing_weight;
  float rcp_n2s2, rcp_div;     // fp32 reciprocals of the two constant divisors
};

// x / c for a constant c with r = fl(1/c): one Newton correction on the product gives the correctly
// rounded quotient wherever the intermediate results are normal (Markstein), at full FMA rate also for
// the denormal tail of the Gaussian - most of the 400 bins - where the IEEE division falls into its slow path.
__device__ __forceinline__ float div_const(float x, float c, float r) {
  const float q = x * r;
  return fmaf(fmaf(-q, c, x), r, q);
}

// One warp per sample: the 401-float row is written as 13 coalesced warp stores (rows are not 16-byte
// aligned, scalar stores keep every row's bytes in full 32-byte sectors except at its two ends); the
// watch time is read once per row and the bin centres come from shared memory.
__global__ void __launch_bounds__(256)
staytime_label_kernel(StayArgs a) {
  extern __shared__ float s_bins[];
  for (int i = threadIdx.x; i < a.nbins; i += blockDim.x) s_bins[i] = a.bins[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = a.nbins + 1;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < a.B; b += warps) {
    const int64_t w = a.watch_ms[b];
    // parse.py:41-43  wt = cast(wt, f32) / 1000.0 ; where(wt > 160, 160, wt)
    float wt = __fdiv_rn((float)w, 1000.0f);
    wt = wt > a.cap_s ? a.cap_s : wt;
    float* out = a.label + (int64_t)b * row;
#pragma unroll 4
    for (int j = lane; j < row; j += 32) {
      float v = wt;                                      // parse.py:64  concat([label, wt], -1)
      if (j < a.nbins) {
        // parse.py:53-63  label = exp(|bin - wt|^2 / (-2 sigma^2)) / (sqrt(2 pi) sigma) * width
        const float dist = s_bins[j] - wt;
        const float sq = dist * dist;
        // exp(x) < 2^-150 rounds to +0 in fp32 whatever follows: skip the tail of the Gaussian (about half of
        // the 400 bins of a row; a warp's 32 consecutive bins take the branch together)
        const float xarg = div_const(sq, a.neg_two_sigma2, a.rcp_n2s2);
        v = xarg < -106.f ? 0.f : div_const(expf(xarg), a.div_num, a.rcp_div) * a.width;
      }
      out[j] = v;
    }
    if (lane == 0) {                                     // per-sample scalars (parse.py:31-39, 66)
      if (a.short_label) a.short_label[b] = w > a.short_ms ? 1 : 0;
      if (a.long_label) a.long_label[b] = w > a.long_ms ? 1 : 0;
      if (a.weight) a.weight[b] = (a.landing && a.landing[b]) ? a.landing_weight : 1.0f;
    }
  }
}

// ------------------------------------------------------------------ binary metrics
// State (int64 / double words, RS_METRIC_STATE_WORDS(T) of them):
//   [0, T+1)        pos_hist[k]: positives whose prediction exceeds exactly k thresholds
//   [T+1, 2T+2)     neg_hist[k]
//   2T+2            n            2T+3  correct (pred > 0.5 == label > 0.5)
//   2T+4            sum_label (double bits)      2T+5  sum_pred (double bits)
constexpr int METRIC_THREADS = 256;

template <typename T>
__device__ __forceinline__ float metric_val(const T* p, int64_t i);
template <>
__device__ __forceinline__ float metric_val<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float metric_val<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}

template <typename PT>
__global__ void __launch_bounds__(METRIC_THREADS)
binary_metrics_kernel(const PT* __restrict__ pred, const float* __restrict__ label,
                      const float* __restrict__ thresholds, int T, int64_t n, float acc_threshold,
                      unsigned long long* __restrict__ state, double* __restrict__ part) {
  extern __shared__ unsigned char metric_smem[];
  float* thr = reinterpret_cast<float*>(metric_smem);                       // [T]
  unsigned int* hist = reinterpret_cast<unsigned int*>(thr + ((T + 3) & ~3));   // [2][T+1]
  __shared__ double red[2][METRIC_THREADS / 32];
  __shared__ unsigned int cnt_correct;
  for (int i = threadIdx.x; i < T; i += blockDim.x) thr[i] = thresholds[i];
  for (int i = threadIdx.x; i < 2 * (T + 1); i += blockDim.x) hist[i] = 0u;
  if (threadIdx.x == 0) cnt_correct = 0u;
  __syncthreads();
  double sl = 0.0, sp = 0.0;
  unsigned int correct = 0;
  const float t_first = thr[0];
  const float t_scale = (T > 1 && thr[T - 1] > thr[0]) ? (float)(T - 1) / (thr[T - 1] - thr[0]) : 0.f;
  // four independent elements per thread and iteration (strided by the grid so that every load
  // instruction of a warp stays coalesced)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    float p[4], y[4];
    int lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      p[u] = in ? metric_val<PT>(pred, i) : 0.f;
      y[u] = in ? label[i] : 0.f;
    }
    // k = #{t : p > thr[t]} for ascending thresholds (Keras compares pred > threshold per threshold):
    // start from the bucket an evenly spaced grid would give (Keras' thresholds are one) and walk to the
    // exact answer - usually 0 or 1 steps; correct for any ascending thresholds
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float g = fminf(fmaxf((p[u] - t_first) * t_scale, 0.f), (float)T);     // NaN -> 0
      int k = (int)g;
      while (k < T && p[u] > thr[k]) ++k;
      while (k > 0 && !(p[u] > thr[k - 1])) --k;
      lo[u] = k;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * stride < n) {
        const bool pos = y[u] > 0.5f;               // labels are cast to bool in Keras' confusion matrix
        atomicAdd(&hist[(pos ? 0 : T + 1) + lo[u]], 1u);
        correct += ((p[u] > acc_threshold) == pos) ? 1u : 0u;
        sl += (double)y[u];
        sp += (double)p[u];
      }
    }
  }
  // ordered block reduction of the two sums: lanes (shuffle tree) -> warps in index order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sl += __shfl_down_sync(0xffffffffu, sl, o);
    sp += __shfl_down_sync(0xffffffffu, sp, o);
    correct += __shfl_down_sync(0xffffffffu, correct, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sl;
    red[1][threadIdx.x >> 5] = sp;
    atomicAdd(&cnt_correct, correct);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (T + 1); i += blockDim.x)
    if (hist[i]) atomicAdd(&state[i], (unsigned long long)hist[i]);
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < METRIC_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
    atomicAdd(&state[2 * T + 3], (unsigned long long)cnt_correct);
  }
}

__global__ void __launch_bounds__(METRIC_THREADS)
binary_metrics_fold_kernel(const double* __restrict__ part, int nparts, int T, int64_t n,
                           unsigned long long* __restrict__ state) {
  // fixed summation tree: thread t takes partials t, t + 256, ...; then a shared-memory tree
  __shared__ double sa[METRIC_THREADS], sb[METRIC_THREADS];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nparts; i += METRIC_THREADS) { a += part[2 * i]; b += part[2 * i + 1]; }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = METRIC_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* sums = reinterpret_cast<double*>(state + 2 * T + 4);
    sums[0] += sa[0];
    sums[1] += sb[0];
    state[2 * T + 2] += (unsigned long long)n;
  }
}

// out[0] AUC  out[1] accuracy  out[2] CTR = sum(label)/n  out[3] COPC = sum(label)/sum(pred)
// out[4] n    out[5] mean prediction
__global__ void binary_metrics_result_kernel(const unsigned long long* __restrict__ state, int T,
                                             double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long* ph = state;
  const unsigned long long* nh = state + T + 1;
  double P = 0.0, N = 0.0;
  for (int k = 0; k <= T; ++k) { P += (double)ph[k]; N += (double)nh[k]; }
  // threshold i: tp_i = #pos with k > i ; walk i = 0 .. T-1 keeping the suffix sums
  double tp = P - (double)ph[0], fp = N - (double)nh[0];
  double auc = 0.0;
  double x_prev = N > 0.0 ? fp / N : 0.0, y_prev = P > 0.0 ? tp / P : 0.0;       // div_no_nan
  for (int i = 1; i < T; ++i) {
    tp -= (double)ph[i];
    fp -= (double)nh[i];
    const double x = N > 0.0 ? fp / N : 0.0, y = P > 0.0 ? tp / P : 0.0;
    auc += (x_prev - x) * (y_prev + y) * 0.5;
    x_prev = x; y_prev = y;
  }
  const double n = (double)state[2 * T + 2];
  const double* sums = reinterpret_cast<const double*>(state + 2 * T + 4);
  out[0] = auc;
  out[1] = n > 0.0 ? (double)state[2 * T + 3] / n : 0.0;
  out[2] = n > 0.0 ? sums[0] / n : 0.0;
  out[3] = sums[1] > 0.0 ? sums[0] / sums[1] : 0.0;
  out[4] = n;
  out[5] = n > 0.0 ? sums[1] / n : 0.0;
}

static int metric_grid(int64_t n) {
  int64_t g = cdiv(n, (int64_t)METRIC_THREADS * 4);
  const int64_t cap = (int64_t)sm_count() * 8;      // 2048 threads per SM: one full wave
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace rs

using namespace rs;

extern "C" {

int rs_staytime_labels(const int64_t* watch_ms, const uint8_t* landing, const float* bins, int nbins,
                       int B, float* staytime_label, int64_t* short_label, int64_t* long_label,
                       float* sample_weight, int64_t short_ms, int64_t long_ms, float cap_s, float sigma,
                       float left, float right, float landing_weight, void* stream) {
  RS_REQUIRE(B >= 0 && nbins >= 2 && nbins <= 8192, "staytime_labels: B=%d nbins=%d", B, nbins);
  RS_REQUIRE(sigma > 0.f, "staytime_labels: sigma=%g", (double)sigma);
  if (B == 0) return 0;
  RS_REQUIRE(watch_ms && bins && staytime_label, "staytime_labels: null pointer");
  StayArgs a;
  a.watch_ms = watch_ms; a.landing = landing; a.bins = bins; a.nbins = nbins; a.B = B;
  a.label = staytime_label; a.short_label = short_label; a.long_label = long_label; a.weight = sample_weight;
  a.short_ms = short_ms; a.long_ms = long_ms; a.cap_s = cap_s;
  // the constants as Python computes them (double) before TF casts them to float32 (parse.py:56-62)
  a.neg_two_sigma2 = (float)(-2.0 * (double)sigma * (double)sigma);
  a.div_num = (float)(sqrt(2.0 * 3.141592653589793) * (double)sigma);
  a.width = (float)(((double)right - (double)left) / (double)(nbins - 1));
  a.landing_weight = landing_weight;
  a.rcp_n2s2 = 1.0f / a.neg_two_sigma2;
  a.rcp_div = 1.0f / a.div_num;
  int64_t grid = cdiv(B, 256 / 32);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  staytime_label_kernel<<<(unsigned)grid, 256, nbins * sizeof(float), as_stream(stream)>>>(a);
  return check_launch("staytime_labels");
}

size_t rs_binary_metrics_state_bytes(int num_thresholds) { return (size_t)(2 * num_thresholds + 6) * 8; }
size_t rs_binary_metrics_workspace_bytes(int64_t n) { return (size_t)metric_grid(n > 0 ? n : 1) * 2 * sizeof(double); }

int rs_binary_metrics_update(const void* pred, int pred_dtype, const float* label, int64_t n,
                             const float* thresholds, int num_thresholds, float acc_threshold,
                             void* state, void* ws, size_t ws_bytes, void* stream) {
  RS_REQUIRE(n >= 0 && num_thresholds >= 2 && num_thresholds <= 4096, "binary_metrics: n=%lld T=%d", (long long)n,
             num_thresholds);
  RS_REQUIRE(pred_dtype == RS_F32 || pred_dtype == RS_BF16, "binary_metrics: bad pred dtype %d", pred_dtype);
  if (n == 0) return 0;
  RS_REQUIRE(pred && label && thresholds && state, "binary_metrics: null pointer");
  const int grid = metric_grid(n);
  if (ws_bytes < (size_t)grid * 2 * sizeof(double) || ws == nullptr) {
    set_error("binary_metrics: workspace %zu < %zu", ws_bytes, (size_t)grid * 2 * sizeof(double));
    return RS_ERR_WORKSPACE;
  }
  const int T = num_thresholds;
  const size_t smem = (size_t)((T + 3) & ~3) * 4 + (size_t)2 * (T + 1) * 4;
  cudaStream_t st = as_stream(stream);
  if (pred_dtype == RS_F32)
    binary_metrics_kernel<float><<<grid, METRIC_THREADS, smem, st>>>(
        (const float*)pred, label, thresholds, T, n, acc_threshold, (unsigned long long*)state, (double*)ws);
  else
    binary_metrics_kernel<__nv_bfloat16><<<grid, METRIC_THREADS, smem, st>>>(
        (const __nv_bfloat16*)pred, label, thresholds, T, n, acc_threshold, (unsigned long long*)state, (double*)ws);
  if (int e = check_launch("binary_metrics_update")) return e;
  binary_metrics_fold_kernel<<<1, METRIC_THREADS, 0, st>>>((const double*)ws, grid, T, n, (unsigned long long*)state);
  return check_launch("binary_metrics_fold");
}

int rs_binary_metrics_result(const void* state, int num_thresholds, double* out6, void* stream) {
  RS_REQUIRE(state && out6 && num_thresholds >= 2, "binary_metrics_result: bad arguments");
  binary_metrics_result_kernel<<<1, 32, 0, as_stream(stream)>>>((const unsigned long long*)state, num_thresholds, out6);
  return check_launch("binary_metrics_result");
}

}  // extern "C"
