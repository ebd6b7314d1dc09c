// Shared helpers for the dpi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/dpi_b200.h"

namespace dpi {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DPI_ERR_CUDA;
  }
  return DPI_OK;
}

#define DPI_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      dpi::set_error(__VA_ARGS__);             \
      return DPI_ERR_INVALID_ARG;              \
    }                                          \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kStatsMaxBlocks = 592;  // 4 CTAs per SM on a 148-SM B200
constexpr int kStatsThreads = 256;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// activation and its derivative expressed through the OUTPUT value (all supported activations
// are invertible in that sense; LeakyReLU is applied in place by the reference, base.py:102)
// The low byte of an `act` argument is the activation code; bit 8 (DPI_ACT_ROUND_TF32) asks the kernel to round
// what it stores to TF32 (round-to-nearest, ties away) because a tcgen05 kind::tf32 MMA will read it and would
// otherwise TRUNCATE the fp32 mantissa.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 maybe_round4(float4 r, int act) {
  if (act & DPI_ACT_ROUND_TF32) {
    r.x = round_tf32(r.x); r.y = round_tf32(r.y); r.z = round_tf32(r.z); r.w = round_tf32(r.w);
  }
  return r;
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act & 0xff) {
    case DPI_ACT_LEAKY_RELU: return x > 0.f ? x : 0.2f * x;
    case DPI_ACT_RELU: return x > 0.f ? x : 0.f;
    case DPI_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case DPI_ACT_TANH: return tanhf(x);
    case DPI_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
__device__ __forceinline__ float act_grad_from_out(float o, int act) {
  switch (act & 0xff) {
    case DPI_ACT_LEAKY_RELU: return o > 0.f ? 1.f : 0.2f;
    case DPI_ACT_RELU: return o > 0.f ? 1.f : 0.f;
    case DPI_ACT_ELU: return o > 0.f ? 1.f : o + 1.f;
    case DPI_ACT_TANH: return 1.f - o * o;
    case DPI_ACT_SIGMOID: return o * (1.f - o);
    default: return 1.f;
  }
}

// Thread-slot decomposition used by every per-channel streaming kernel: the tensor is viewed as
// nvox x G float4 groups (G = C/4).  `slots` (a multiple of G) threads are active; slot q owns
// channel group q % G for its whole life and visits voxels q/G, q/G + slots/G, ...
struct SlotPlan {
  int blocks;
  int64_t slots;       // active thread slots, multiple of G
  int64_t vox_step;    // slots / G
};
inline SlotPlan make_slot_plan(int64_t nvox, int C, int max_blocks = kStatsMaxBlocks, int threads = kStatsThreads) {
  const int G = C / 4;
  int64_t total = nvox * G;
  int64_t want_blocks = ceil_div64(total, (int64_t)threads * 4);  // >=4 float4 per thread
  if (want_blocks < 1) want_blocks = 1;
  if (want_blocks > max_blocks) want_blocks = max_blocks;
  // need at least G slots
  while (want_blocks * threads < G) ++want_blocks;
  SlotPlan p;
  p.blocks = (int)want_blocks;
  int64_t q = (int64_t)p.blocks * threads;
  p.slots = (q / G) * G;
  p.vox_step = p.slots / G;
  return p;
}

}  // namespace dpi
