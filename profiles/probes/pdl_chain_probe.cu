// Probe for DESIGN.md §10 item 8 (NOT part of the product; written at the end of round 1, compile-checked only —
// run it first thing in round 2):  what does one launch of a DEPENDENT chain cost inside a CUDA graph on a B200, and
// how much of it does programmatic dependent launch (PDL) remove?
//
// A chain of `len` kernels, each reading what its predecessor wrote (x[i] = x[i] + 1 over n floats), captured into a
// graph three ways:
//   plain    ordinary stream order (full dependency edges)
//   pdl      every kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization, waits with
//            griddepcontrol.wait before touching memory and triggers its dependents at its first instruction
//   pdl-late the same, trigger after the main loop (just before the stores drain)
// Prints microseconds per launch for a few problem sizes (1 CTA ... one full wave of 4 CTAs per SM).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pdl_chain_probe pdl_chain_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int PDL>   // 0 plain, 1 trigger first, 2 trigger late
__global__ void __launch_bounds__(256) step(float* __restrict__ x, int64_t n) {
  if (PDL == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += 1.f;
  if (PDL == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <int PDL>
static cudaError_t launch(float* x, int64_t n, int blocks, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = PDL ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, step<PDL>, x, n);
}

template <int PDL>
static float time_chain(float* x, int64_t n, int blocks, int len, cudaStream_t st) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < len; ++i)
    if (launch<PDL>(x, n, blocks, st) != cudaSuccess) break;
  if (cudaStreamEndCapture(st, &g) != cudaSuccess || cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
    printf("capture/instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return -1.f;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, st);
    cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best * 1e3f / len;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int len = 400;
  const int64_t nmax = (int64_t)sms * 4 * 256 * 64;
  float* x = nullptr;
  cudaMalloc(&x, nmax * sizeof(float));
  cudaMemset(x, 0, nmax * sizeof(float));
  cudaStream_t st;
  cudaStreamCreate(&st);
  struct { const char* what; int blocks; int64_t n; } cases[] = {
      {"1 CTA, 256 floats", 1, 256},
      {"1 CTA per SM, 4 floats per thread", sms, (int64_t)sms * 256 * 4},
      {"4 CTAs per SM, 4 floats per thread (one wave)", sms * 4, (int64_t)sms * 4 * 256 * 4},
      {"4 CTAs per SM, 64 floats per thread", sms * 4, nmax}};
  for (auto& c : cases) {
    const float a = time_chain<0>(x, c.n, c.blocks, len, st);
    const float b = time_chain<1>(x, c.n, c.blocks, len, st);
    const float d = time_chain<2>(x, c.n, c.blocks, len, st);
    printf("%-48s plain %6.2f us/launch   pdl %6.2f   pdl-late %6.2f\n", c.what, a, b, d);
  }
  // the chain must still be a chain: 3 variants x (1 + 6 graph launches) x len increments of x[0] per case that covers it
  float h = 0.f;
  cudaMemcpy(&h, x, sizeof(float), cudaMemcpyDeviceToHost);
  printf("x[0] = %.0f (expected %d)\n", h, 4 * 3 * 7 * len);
  cudaFree(x);
  return 0;
}
