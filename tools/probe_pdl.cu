// Probe: cost of a kernel -> kernel dependency inside a CUDA graph on one stream, with and without programmatic
// dependent launch (griddepcontrol).  Each kernel: 148 CTAs x 128 threads, a prologue (smem init + barrier), then
// ~WORK_US of dependent global traffic.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/probe_pdl tools/probe_pdl.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

template <bool PDL>
__global__ void __launch_bounds__(128) chain_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int spin) {
  __shared__ float s[1024];
  if (PDL) asm volatile("griddepcontrol.launch_dependents;");
  for (int i = threadIdx.x; i < 1024; i += 128) s[i] = (float)i;
  __syncthreads();
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  float acc = s[threadIdx.x];
  for (int r = 0; r < spin; ++r)
    for (int i = blockIdx.x * 128 + threadIdx.x; i < n; i += gridDim.x * 128) acc += in[i] * 1.0001f;
  for (int i = blockIdx.x * 128 + threadIdx.x; i < n; i += gridDim.x * 128) out[i] = in[i] + acc * 1e-30f;
}

static float run(bool pdl, int nk, int n, int spin, float* a, float* b, cudaStream_t st) {
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < nk; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    const float* src = (k & 1) ? b : a; float* dst = (k & 1) ? a : b;
    if (pdl) CK(cudaLaunchKernelEx(&cfg, chain_kernel<true>, src, dst, n, spin));
    else CK(cudaLaunchKernelEx(&cfg, chain_kernel<false>, src, dst, n, spin));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e0, st));
  const int reps = 50;
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  return ms * 1000.f / reps / nk;
}

int main() {
  const int n = 1 << 20;
  float *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int spin : {0, 2, 8}) {
    for (int nk : {1, 20}) {
      float t0 = run(false, nk, n, spin, a, b, st), t1 = run(true, nk, n, spin, a, b, st);
      printf("spin %d  kernels/graph %2d : %.2f us per kernel plain, %.2f us with PDL\n", spin, nk, t0, t1);
    }
  }
  // check the chain result is the same both ways (data dependence honoured)
  printf("ok\n");
  return 0;
}
