// Loop-body kernels that are not convolutions: input-noise perturbation (Philox), the masked
// sampling-operator loss with its gradient and the SNR / Pearson sums, fused flat Adam.
#include "dpi_common.cuh"

namespace dpi {

// ---- Philox4x32-10 -------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// four N(0,1) samples from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint32_t r[4];
  philox4x32_10(seed, idx, offset, r);
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = (r[0] + 0.5f) * k * 1.0f, u1 = (r[1] + 0.5f) * k;
  const float u2 = (r[2] + 0.5f) * k * 1.0f, u3 = (r[3] + 0.5f) * k;
  // clamp away from 0 (r+0.5 rounds to >= 0.5 in fp32, so u >= 1.16e-10)
  const float ra = sqrtf(-2.f * __logf(fminf(fmaxf(u0, 1e-10f), 1.f)));
  const float rb = sqrtf(-2.f * __logf(fminf(fmaxf(u2, 1e-10f), 1.f)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

__global__ void noise_axpy_kernel(const float* __restrict__ z, const float* __restrict__ eps, float* __restrict__ out,
                                  int64_t n4, float sigma, uint64_t seed, uint64_t offset, int rnd) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z) + i);
    float4 e;
    if (eps) e = __ldg(reinterpret_cast<const float4*>(eps) + i);
    else e = normal4(seed, offset, (uint64_t)i);
    float4 r;
    r.x = fmaf(sigma, e.x, zv.x); r.y = fmaf(sigma, e.y, zv.y);
    r.z = fmaf(sigma, e.z, zv.z); r.w = fmaf(sigma, e.w, zv.w);
    r = maybe_round4(r, rnd);
    reinterpret_cast<float4*>(out)[i] = r;
  }
}

// same, with the Philox offset read from device memory (CUDA-graph replay)
__global__ void noise_axpy_dev_kernel(const float* __restrict__ z, float* __restrict__ out, int64_t n4, float sigma,
                                      uint64_t seed, const uint64_t* __restrict__ counter, int rnd) {
  const uint64_t offset = counter[0];
  seed += counter[1];   // per-patch seed lives on the device so one captured graph serves every patch
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z) + i);
    const float4 e = normal4(seed, offset, (uint64_t)i);
    float4 r;
    r.x = fmaf(sigma, e.x, zv.x); r.y = fmaf(sigma, e.y, zv.y);
    r.z = fmaf(sigma, e.z, zv.z); r.w = fmaf(sigma, e.w, zv.w);
    r = maybe_round4(r, rnd);
    reinterpret_cast<float4*>(out)[i] = r;
  }
}

// `--data_forgetting_factor F` (main.py:86-97,153-155): during the first F iterations the decimated data, repeated
// along the channel axis and normalised to the noise std once per patch, is added to the network input with the
// weight logspace(0,-4,F)[iteration].  The iteration index is read from device memory so that the launch can live in
// a replayed CUDA graph; from iteration F on the kernel returns without touching memory.
__global__ void add_data_dev_kernel(float* __restrict__ zin, int64_t ld, int C, int64_t nvox,
                                    const float* __restrict__ data, int64_t data_ld, int Cd,
                                    const float* __restrict__ weights, int F, const uint64_t* __restrict__ counter,
                                    int rnd) {
  const uint64_t it = counter[0];
  if (it >= (uint64_t)F) return;
  const float w = weights[it];
  const int64_t total = nvox * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / C;
    const int c = (int)(i - v * C);
    float r = fmaf(w, __ldg(data + v * data_ld + (c % Cd)), zin[v * ld + c]);
    if (rnd) r = round_tf32(r);
    zin[v * ld + c] = r;
  }
}

// Depth-wise FIR along one axis of a contiguous [outer][T][inner] tensor, "same" size, zero beyond the ends:
//   y[o][t][i] = sum_m taps[m] * x[o][t + pad - m][i],  pad = ntaps / 2
// = ConvolveKernel_1d.forward (utils/processing.py:34-67: conv_transpose{1,2,3}d with the taps along the time axis,
// padding = pad, groups = channels), the input-noise pre-filter of main.py:66-84.  Runs once per patch.
__global__ void fir_axis_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t outer, int64_t T,
                                int64_t inner, const float* __restrict__ taps, int ntaps) {
  const int pad = ntaps / 2;
  const int64_t total = outer * T * inner;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t in = i % inner;
    const int64_t t = (i / inner) % T;
    const int64_t o = i / (inner * T);
    const float* col = x + o * T * inner + in;
    float acc = 0.f;
    for (int m = 0; m < ntaps; ++m) {
      const int64_t ts = t + pad - m;
      if (ts >= 0 && ts < T) acc = fmaf(__ldg(taps + m), __ldg(col + ts * inner), acc);
    }
    y[i] = acc;
  }
}

// End-of-iteration bookkeeping kept on the device so the loop needs no host round trip
// (main.py:165-182): history row, best-output flag, iteration counter, Adam step.
//   counter[0] = iteration index (also the Philox offset), hyper = {lr, adam_step}
//   best_state = {loss_min, copy_flag}
__global__ void iteration_end_kernel(const double* __restrict__ scalars, double* __restrict__ hyper,
                                     uint64_t* __restrict__ counter, double* __restrict__ history, int64_t max_iters,
                                     double* __restrict__ best_state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint64_t it = counter[0];
  if ((int64_t)it < max_iters && history) {
    history[it * 4 + 0] = scalars[0];
    history[it * 4 + 1] = scalars[1];
    history[it * 4 + 2] = scalars[2];
    history[it * 4 + 3] = hyper[0];
  }
  const double loss = scalars[0];
  const bool better = (it == 0) || (loss <= best_state[0]);
  if (better) best_state[0] = loss;
  best_state[1] = better ? 1.0 : 0.0;
  counter[0] = it + 1;
  hyper[1] += 1.0;
}

__global__ void copy_if_flag_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n4,
                                    const double* __restrict__ best_state) {
  if (best_state[1] == 0.0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
}

__global__ void fill_normal_kernel(float* __restrict__ out, int64_t n, float mean, float std, uint64_t seed,
                                   uint64_t offset) {
  const int64_t n4 = (n + 3) / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 e = normal4(seed, offset, (uint64_t)i);
    const float v[4] = {e.x, e.y, e.z, e.w};
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = fmaf(std, v[j], mean);
  }
}

// ---- masked loss + metrics ---------------------------------------------------------------------------
constexpr int kLossBlocks = 592, kLossThreads = 256, kLossSums = 8;
// sums: 0 sum|d| or sum d^2   1 sum img^2   2 sum (img-out)^2   3 sum out   4 sum img   5 sum out*img
//       6 sum out^2           7 unused

__global__ void __launch_bounds__(kLossThreads)
masked_loss_kernel(const float* __restrict__ out, const float* __restrict__ img, const float* __restrict__ mask,
                   int64_t n, int64_t n_logical, int kind, float* __restrict__ dout, double* __restrict__ partial) {
  double s[kLossSums];
#pragma unroll
  for (int k = 0; k < kLossSums; ++k) s[k] = 0.0;
  const float inv_n = 1.f / (float)n_logical;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(out) + i);
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(img) + i);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mask) + i);
    const float o[4] = {o4.x, o4.y, o4.z, o4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
    float gr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = o[j] * m[j] - t[j] * m[j];
      if ((kind & 0xff) == DPI_LOSS_MAE) {
        s[0] += fabsf(d);
        gr[j] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * m[j] * inv_n;
      } else {
        s[0] += (double)d * d;
        gr[j] = 2.f * d * m[j] * inv_n;
      }
      if (d != d) gr[j] = d;  // propagate NaN
      const float e = t[j] - o[j];
      s[1] += (double)t[j] * t[j];
      s[2] += (double)e * e;
      s[3] += o[j];
      s[4] += t[j];
      s[5] += (double)o[j] * t[j];
      s[6] += (double)o[j] * o[j];
    }
    if (dout) reinterpret_cast<float4*>(dout)[i] = maybe_round4(make_float4(gr[0], gr[1], gr[2], gr[3]), kind);
  }
  // fixed-order block reduction
  __shared__ double sm[kLossThreads / 32][kLossSums];
#pragma unroll
  for (int k = 0; k < kLossSums; ++k) {
    double v = s[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLossSums) {
    double v = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) v += sm[w][threadIdx.x];
    partial[(size_t)blockIdx.x * kLossSums + threadIdx.x] = v;
  }
}

__global__ void masked_loss_finalize_kernel(const double* __restrict__ partial, int nblk, int64_t n_logical,
                                            double* __restrict__ scalars) {
  __shared__ double tot[kLossSums];
  if (threadIdx.x < kLossSums) {
    double v = 0.0;
    for (int b = 0; b < nblk; ++b) v += partial[(size_t)b * kLossSums + threadIdx.x];
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double N = (double)n_logical;
    const double loss = tot[0] / N;
    const double snr = 10.0 * log10(tot[1] / tot[2]);
    const double mo = tot[3] / N, mt = tot[4] / N;
    const double cov = tot[5] - N * mo * mt;
    const double vo = tot[6] - N * mo * mo, vt = tot[1] - N * mt * mt;
    const double pc = cov / (sqrt(vt) * sqrt(vo));
    scalars[0] = loss;
    scalars[1] = snr;
    scalars[2] = pc;
    scalars[3] = (loss != loss) ? 1.0 : 0.0;
    scalars[4] = tot[1];
    scalars[5] = tot[2];
    scalars[6] = tot[3];
    scalars[7] = tot[4];
  }
}

// ---- Adam ------------------------------------------------------------------------------------------
// torch.optim.Adam single-tensor semantics:  m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ;
// p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const double* __restrict__ hyper, double lr_h,
                            int64_t step_h, double beta1_d, double beta2_d, double eps_d, double wd_d) {
  double lr = lr_h;
  double step = (double)step_h;
  if (hyper) { lr = hyper[0]; step = hyper[1]; }
  const double bc1 = 1.0 - pow(beta1_d, step);
  const double bc2 = 1.0 - pow(beta2_d, step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  // python-double scalars are rounded to fp32 once, exactly like torch's Scalar -> float conversion
  const float omb1 = (float)(1.0 - beta1_d), beta2 = (float)beta2_d, omb2 = (float)(1.0 - beta2_d);
  const float eps = (float)eps_d, wd = (float)wd_d;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(p)[i];
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = gg[j];
      if (wd != 0.f) gj = fmaf(wd, pp[j], gj);
      mm[j] = mm[j] + omb1 * (gj - mm[j]);                      // lerp form used by torch
      vv[j] = beta2 * vv[j] + omb2 * gj * gj;
      const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
      pp[j] = pp[j] - step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  // tail
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float gj = g[i];
    if (wd != 0.f) gj = fmaf(wd, p[i], gj);
    const float mj = m[i] + omb1 * (gj - m[i]);
    const float vj = beta2 * v[i] + omb2 * gj * gj;
    m[i] = mj; v[i] = vj;
    p[i] = p[i] - step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
  }
}

}  // namespace dpi

using namespace dpi;

// y = a * y + b * x  (gradient accumulation over the batch rows of the shared-network mode; a == 0 never reads y)
__global__ void __launch_bounds__(256) axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b,
                                                    int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float4 r;
    if (a == 0.f) {
      r = make_float4(b * xv.x, b * xv.y, b * xv.z, b * xv.w);
    } else {
      const float4 yv = reinterpret_cast<const float4*>(y)[i];
      r = make_float4(fmaf(a, yv.x, b * xv.x), fmaf(a, yv.y, b * xv.y), fmaf(a, yv.z, b * xv.z), fmaf(a, yv.w, b * xv.w));
    }
    reinterpret_cast<float4*>(y)[i] = r;
  }
}

extern "C" {

int dpi_noise_axpy(const float* z, const float* eps, float* out, int64_t n, float sigma, uint64_t seed,
                   uint64_t offset, int round_tf32, void* stream) {
  DPI_REQUIRE(z && out && aligned16(z) && aligned16(out) && (n & 3) == 0 && (!eps || aligned16(eps)),
              "dpi_noise_axpy: need 16B-aligned pointers and n %% 4 == 0");
  const int64_t n4 = n >> 2;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  noise_axpy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(z, eps, out, n4, sigma, seed, offset,
                                                              round_tf32 ? DPI_ACT_ROUND_TF32 : 0);
  return check_launch("dpi_noise_axpy");
}

int dpi_noise_axpy_dev(const float* z, float* out, int64_t n, float sigma, uint64_t seed,
                       const uint64_t* counter_dev, int round_tf32, void* stream) {
  DPI_REQUIRE(z && out && counter_dev && aligned16(z) && aligned16(out) && (n & 3) == 0,
              "dpi_noise_axpy_dev: need 16B-aligned pointers and n %% 4 == 0");
  const int64_t n4 = n >> 2;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  noise_axpy_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(z, out, n4, sigma, seed, counter_dev,
                                                                  round_tf32 ? DPI_ACT_ROUND_TF32 : 0);
  return check_launch("dpi_noise_axpy_dev");
}

int dpi_add_data_dev(float* zin, int64_t ld, int C, int64_t nvox, const float* data, int64_t data_ld, int Cd,
                     const float* weights_dev, int F, const uint64_t* counter_dev, int round_tf32, void* stream) {
  DPI_REQUIRE(zin && data && weights_dev && counter_dev && C > 0 && Cd > 0 && ld >= C && data_ld >= Cd && nvox > 0 && F > 0,
              "dpi_add_data_dev: bad arguments");
  const int64_t total = nvox * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  add_data_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(zin, ld, C, nvox, data, data_ld, Cd, weights_dev, F,
                                                                counter_dev, round_tf32);
  return check_launch("dpi_add_data_dev");
}

int dpi_fir_axis(const float* x, float* y, int64_t outer, int64_t T, int64_t inner, const float* taps_dev, int ntaps,
                 void* stream) {
  DPI_REQUIRE(x && y && x != y && taps_dev && outer > 0 && T > 0 && inner > 0 && ntaps > 0 && (ntaps & 1),
              "dpi_fir_axis: need distinct x / y, positive sizes and an odd number of taps (got %d)", ntaps);
  const int64_t total = outer * T * inner;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  fir_axis_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, outer, T, inner, taps_dev, ntaps);
  return check_launch("dpi_fir_axis");
}

int dpi_iteration_end(const double* scalars, double* hyper_dev, uint64_t* counter_dev, double* history,
                      int64_t max_iters, double* best_state, const float* out, float* best, int64_t n,
                      void* stream) {
  DPI_REQUIRE(scalars && hyper_dev && counter_dev && best_state, "dpi_iteration_end: null pointer");
  iteration_end_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scalars, hyper_dev, counter_dev, history, max_iters,
                                                          best_state);
  int rc = check_launch("dpi_iteration_end");
  if (rc) return rc;
  if (out && best) {
    DPI_REQUIRE(aligned16(out) && aligned16(best) && (n & 3) == 0, "dpi_iteration_end: unaligned output buffers");
    const int64_t n4 = n >> 2;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    copy_if_flag_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, best, n4, best_state);
    rc = check_launch("dpi_iteration_end(copy)");
  }
  return rc;
}

int dpi_fill_normal(float* out, int64_t n, float mean, float std, uint64_t seed, uint64_t offset,
                    void* stream) {
  DPI_REQUIRE(out && n >= 0, "dpi_fill_normal: bad arguments");
  int blocks = (int)(((n + 3) / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  fill_normal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, n, mean, std, seed, offset);
  return check_launch("dpi_fill_normal");
}

int64_t dpi_loss_workspace_bytes(void) { return (int64_t)kLossBlocks * kLossSums * 8; }

int dpi_masked_loss(const float* out, const float* img, const float* mask, int64_t n, int64_t n_logical,
                    int kind, float* dout, void* workspace, int64_t workspace_bytes, double* scalars_out,
                    void* stream) {
  DPI_REQUIRE(out && img && mask && scalars_out && aligned16(out) && aligned16(img) && aligned16(mask) &&
                  (!dout || aligned16(dout)) && (n & 3) == 0 && n > 0 && n_logical > 0,
              "dpi_masked_loss: need 16B-aligned pointers and n %% 4 == 0");
  if (!workspace || workspace_bytes < dpi_loss_workspace_bytes()) {
    set_error("dpi_masked_loss: workspace too small");
    return DPI_ERR_WORKSPACE;
  }
  int blocks = (int)(((n >> 2) + kLossThreads - 1) / kLossThreads);
  if (blocks > kLossBlocks) blocks = kLossBlocks;
  masked_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(out, img, mask, n, n_logical, kind, dout,
                                                                       (double*)workspace);
  int rc = check_launch("dpi_masked_loss");
  if (rc) return rc;
  masked_loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, n_logical,
                                                                 scalars_out);
  return check_launch("dpi_masked_loss(finalize)");
}

static int adam_launch(float* p, const float* g, float* m, float* v, int64_t n, const double* hyper, double lr,
                       int64_t step, double beta1, double beta2, double eps, double wd, void* stream) {
  DPI_REQUIRE(p && g && m && v && aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && n > 0,
              "dpi_adam_step: need 16B-aligned pointers");
  int blocks = (int)(((n >> 2) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper, lr, step, beta1, beta2, eps, wd);
  return check_launch("dpi_adam_step");
}

int dpi_axpby(float* y, const float* x, float a, float b, int64_t n, void* stream) {
  DPI_REQUIRE(y && x && n >= 0 && (n & 3) == 0 && aligned16(y) && aligned16(x), "dpi_axpby: unaligned / null buffers");
  if (n == 0) return DPI_OK;
  const int64_t n4 = n >> 2;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  axpby_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, x, a, b, n4);
  return check_launch("dpi_axpby");
}

int dpi_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                  double eps, double weight_decay, int64_t step, void* stream) {
  DPI_REQUIRE(step >= 1, "dpi_adam_step: step counts from 1");
  return adam_launch(p, g, m, v, n, nullptr, lr, step, beta1, beta2, eps, weight_decay, stream);
}

int dpi_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const double* hyper_dev,
                      double beta1, double beta2, double eps, double weight_decay, void* stream) {
  DPI_REQUIRE(hyper_dev, "dpi_adam_step_dev: null hyper pointer");
  return adam_launch(p, g, m, v, n, hyper_dev, 0.0, 1, beta1, beta2, eps, weight_decay, stream);
}

}  // extern "C"
