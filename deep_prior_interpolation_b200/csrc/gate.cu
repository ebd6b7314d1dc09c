// Grid-attention gate of the attention MultiRes U-Net (architectures/attention.py:86-113, `return x * psi`):
// a one-channel attention map psi, already up-sampled to the resolution of the skip tensor x, scales every channel
// of x.  Channels-last fp32, 16-byte channel groups; psi is a 4-channel-padded tensor whose channel 0 is the map.
//
// HBM-bound: forward reads x (+4 B of psi per voxel) and writes y; backward reads dy and x once and produces both
// dx = dy * psi and dpsi = sum_c dy * x.  The channel sum is a fixed-order register sum per lane followed by an xor
// butterfly over the lanes that share a voxel: no atomics, bit-reproducible.
#include "dpi_common.cuh"

namespace dpi {

__global__ void __launch_bounds__(256)
gate_mul_fwd_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ psi, int64_t psi_ld,
                    float* __restrict__ y, int64_t y_ld, int64_t nvox, int G, int flags) {
  const int64_t total = nvox * G;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = q / G;
    const int g = (int)(q - v * G);
    const float s = __ldg(psi + v * psi_ld);
    float4 a = *reinterpret_cast<const float4*>(x + v * x_ld + 4 * g);
    a.x *= s; a.y *= s; a.z *= s; a.w *= s;
    a = maybe_round4(a, flags);
    *reinterpret_cast<float4*>(y + v * y_ld + 4 * g) = a;
  }
}

// WD lanes (a power of two <= 32) share one voxel; lane l of the group owns channel groups l, l + WD, ...
template <int WD>
__global__ void __launch_bounds__(256)
gate_mul_bwd_kernel(const float* __restrict__ dy, int64_t dy_ld, const float* __restrict__ x, int64_t x_ld,
                    const float* __restrict__ psi, int64_t psi_ld, float* __restrict__ dx, int64_t dx_ld,
                    float* __restrict__ dpsi, int64_t dpsi_ld, int64_t nvox, int G, int accumulate_dx) {
  constexpr int VPW = 32 / WD;                        // voxels per warp and step
  const int lane = threadIdx.x & 31;
  const int sub = lane / WD, l = lane % WD;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // the trip count is uniform over the warp: every lane takes part in the shuffles
  for (int64_t base = warp * VPW; base < nvox; base += nwarps * VPW) {
    const int64_t v = base + sub;
    float acc = 0.f;
    if (v < nvox) {
      const float s = __ldg(psi + v * psi_ld);
      for (int g = l; g < G; g += WD) {
        const float4 d = *reinterpret_cast<const float4*>(dy + v * dy_ld + 4 * g);
        const float4 a = *reinterpret_cast<const float4*>(x + v * x_ld + 4 * g);
        acc += (d.x * a.x + d.y * a.y) + (d.z * a.z + d.w * a.w);
        // (__fmul_rn: no contraction with the accumulate below - autograd adds the rounded product)
        float4 r = make_float4(__fmul_rn(d.x, s), __fmul_rn(d.y, s), __fmul_rn(d.z, s), __fmul_rn(d.w, s));
        float* o = dx + v * dx_ld + 4 * g;
        if (accumulate_dx) {
          const float4 old = *reinterpret_cast<const float4*>(o);
          r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
        }
        *reinterpret_cast<float4*>(o) = r;
      }
    }
#pragma unroll
    for (int m = WD / 2; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (l == 0 && v < nvox) *reinterpret_cast<float4*>(dpsi + v * dpsi_ld) = make_float4(acc, 0.f, 0.f, 0.f);
  }
}

static int check_gate(const float* p, int64_t ld, int C, const char* what) {
  if (!p || !aligned16(p) || (ld & 3) || ld < C) {
    set_error("%s: pointer must be 16-byte aligned and the pitch a multiple of 4 >= C (ld=%lld, C=%d)", what,
              (long long)ld, C);
    return DPI_ERR_INVALID_ARG;
  }
  return DPI_OK;
}

}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_gate_mul_fwd(const float* x, int64_t x_ld, const float* psi, int64_t psi_ld, float* y, int64_t y_ld,
                     int64_t nvox, int C, int flags, void* stream) {
  DPI_REQUIRE(C > 0 && (C & 3) == 0 && nvox > 0, "dpi_gate_mul_fwd: bad channel count %d / voxel count %lld", C,
              (long long)nvox);
  int rc = check_gate(x, x_ld, C, "dpi_gate_mul_fwd(x)");
  if (rc) return rc;
  rc = check_gate(y, y_ld, C, "dpi_gate_mul_fwd(y)");
  if (rc) return rc;
  rc = check_gate(psi, psi_ld, 4, "dpi_gate_mul_fwd(psi)");
  if (rc) return rc;
  const int G = C / 4;
  int64_t blocks = ceil_div64(nvox * G, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  gate_mul_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, x_ld, psi, psi_ld, y, y_ld, nvox, G, flags);
  return check_launch("dpi_gate_mul_fwd");
}

int dpi_gate_mul_bwd(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld, const float* psi, int64_t psi_ld,
                     float* dx, int64_t dx_ld, float* dpsi, int64_t dpsi_ld, int64_t nvox, int C, int accumulate_dx,
                     void* stream) {
  DPI_REQUIRE(C > 0 && (C & 3) == 0 && nvox > 0, "dpi_gate_mul_bwd: bad channel count %d / voxel count %lld", C,
              (long long)nvox);
  int rc = check_gate(dy, dy_ld, C, "dpi_gate_mul_bwd(dy)");
  if (rc) return rc;
  rc = check_gate(x, x_ld, C, "dpi_gate_mul_bwd(x)");
  if (rc) return rc;
  rc = check_gate(dx, dx_ld, C, "dpi_gate_mul_bwd(dx)");
  if (rc) return rc;
  rc = check_gate(psi, psi_ld, 4, "dpi_gate_mul_bwd(psi)");
  if (rc) return rc;
  rc = check_gate(dpsi, dpsi_ld, 4, "dpi_gate_mul_bwd(dpsi)");
  if (rc) return rc;
  const int G = C / 4;
  int wd = 1;
  while (wd < G && wd < 32) wd <<= 1;
  const int vpw = 32 / wd;
  int64_t blocks = ceil_div64(nvox, (int64_t)vpw * 8);      // 8 warps per CTA, one step each as a start
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = (cudaStream_t)stream;
#define DPI_GATE_BWD(W)                                                                                           \
  gate_mul_bwd_kernel<W><<<(unsigned)blocks, 256, 0, st>>>(dy, dy_ld, x, x_ld, psi, psi_ld, dx, dx_ld, dpsi, dpsi_ld, \
                                                          nvox, G, accumulate_dx)
  switch (wd) {
    case 1: DPI_GATE_BWD(1); break;
    case 2: DPI_GATE_BWD(2); break;
    case 4: DPI_GATE_BWD(4); break;
    case 8: DPI_GATE_BWD(8); break;
    case 16: DPI_GATE_BWD(16); break;
    default: DPI_GATE_BWD(32); break;
  }
#undef DPI_GATE_BWD
  return check_launch("dpi_gate_mul_bwd");
}

}  // extern "C"
