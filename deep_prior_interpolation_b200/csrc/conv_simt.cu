// Exact-fp32 implicit-GEMM convolutions on the CUDA cores (DPI_PREC_FP32).
//
// This is the reference-accurate path: forward, data-gradient and weight-gradient of the
// 3x3x3 / 1x1x1 / stride-2 convolutions of the MultiRes U-Net computed with FFMA and fp32
// accumulation, no operand rounding.  It is also what the tcgen05 path (conv_tc.cu) falls back to
// for shapes it does not cover (stride 2) and the checker it is validated against on the GPU.
//
// Forward and dgrad are the same "gather convolution"
//     out[v][n] = sum_{tap,c} in[src(v,tap)][c] * Wp[n][tap][c]
// with  src = v*s + tap - pad           (forward)
//       src = (v + pad - tap)/s         (dgrad; only when divisible and in range)
// GEMM view: M = output voxels, N = output channels, K = taps*C.  CTA tile BM x BN, K step 16,
// register micro-tile TM x TN.
#include "dpi_common.cuh"
#include "conv_geom.cuh"

namespace dpi {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_gather_simt(const float* __restrict__ in, int64_t in_ld, const float* __restrict__ Wp,
                 const float* __restrict__ bias, float* __restrict__ out, int64_t out_ld, GatherGeom g,
                 int accumulate) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int APAD = 4;
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int64_t nvox_out = (int64_t)g.Do * g.Ho * g.Wo;
  const int64_t m_base = (int64_t)blockIdx.x * BM;
  const int n_base = blockIdx.y * BN;

  // A-tile loader mapping: BM rows x 4 float4 per row
  constexpr int A_ITERS = (BM * 4) / NT;
  static_assert((BM * 4) % NT == 0, "A loader mapping");
  int a_d[A_ITERS], a_h[A_ITERS], a_w[A_ITERS];
  bool a_ok[A_ITERS];
#pragma unroll
  for (int i = 0; i < A_ITERS; ++i) {
    const int idx = tid + i * NT;
    const int row = idx >> 2;
    int64_t v = m_base + row;
    a_ok[i] = v < nvox_out;
    if (!a_ok[i]) v = 0;
    a_w[i] = (int)(v % g.Wo);
    v /= g.Wo;
    a_h[i] = (int)(v % g.Ho);
    a_d[i] = (int)(v / g.Ho);
  }
  // B-tile loader mapping: BN rows (n) x 4 float4
  constexpr int B_ITERS = (BN * 4 + NT - 1) / NT;

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int taps = g.kd * g.kh * g.kw;
  const int Ktot = taps * g.C;
  for (int tap = 0; tap < taps; ++tap) {
    const int tkw = tap % g.kw;
    const int tkh = (tap / g.kw) % g.kh;
    const int tkd = tap / (g.kw * g.kh);
    // per-row source voxel for this tap
    int64_t a_src[A_ITERS];
#pragma unroll
    for (int i = 0; i < A_ITERS; ++i) {
      int sd_, sh_, sw_;
      bool ok = a_ok[i];
      if (!g.transposed) {
        sd_ = a_d[i] * g.sd + tkd - g.pd;
        sh_ = a_h[i] * g.sh + tkh - g.ph;
        sw_ = a_w[i] * g.sw + tkw - g.pw;
      } else {
        const int nd = a_d[i] + g.pd - tkd, nh = a_h[i] + g.ph - tkh, nw = a_w[i] + g.pw - tkw;
        ok = ok && nd >= 0 && nh >= 0 && nw >= 0 && (nd % g.sd == 0) && (nh % g.sh == 0) && (nw % g.sw == 0);
        sd_ = nd / g.sd; sh_ = nh / g.sh; sw_ = nw / g.sw;
      }
      ok = ok && sd_ >= 0 && sd_ < g.Di && sh_ >= 0 && sh_ < g.Hi && sw_ >= 0 && sw_ < g.Wi;
      a_src[i] = ok ? (((int64_t)sd_ * g.Hi + sh_) * g.Wi + sw_) : -1;
    }
    for (int c0 = 0; c0 < g.C; c0 += BK) {
      // ---- load tiles
#pragma unroll
      for (int i = 0; i < A_ITERS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx >> 2, kq = (idx & 3) * 4;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_src[i] >= 0 && c0 + kq < g.C)
          val = __ldg(reinterpret_cast<const float4*>(in + a_src[i] * in_ld + c0 + kq));
        As[kq + 0][row] = val.x;
        As[kq + 1][row] = val.y;
        As[kq + 2][row] = val.z;
        As[kq + 3][row] = val.w;
      }
#pragma unroll
      for (int i = 0; i < B_ITERS; ++i) {
        const int idx = tid + i * NT;
        if (idx < BN * 4) {
          const int n = idx >> 2, kq = (idx & 3) * 4;
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n_base + n < g.N && c0 + kq < g.C)
            val = __ldg(reinterpret_cast<const float4*>(Wp + (int64_t)(n_base + n) * Ktot + tap * g.C + c0 + kq));
          Bs[kq + 0][n] = val.x;
          Bs[kq + 1][n] = val.y;
          Bs[kq + 2][n] = val.z;
          Bs[kq + 3][n] = val.w;
        }
      }
      __syncthreads();
      // ---- compute
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t v = m_base + ty * TM + i;
    if (v >= nvox_out) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n_base + tx * TN + j;
      if (n >= g.N) continue;
      float r = acc[i][j];
      if (bias) r += __ldg(bias + n);
      float* o = out + v * out_ld + n;
      if (accumulate) r += *o;
      *o = r;
    }
  }
}

// ---- weight gradient -------------------------------------------------------------------------
// dW[n][tap][c] = sum_v dy[v][n] * x[src(v,tap)][c].  CTA = (voxel chunk, tap, n-tile, c-tile);
// partial tiles go to workspace[chunk][N][taps][C]; wgrad_reduce sums chunks in order.
constexpr int WG_BN = 32, WG_BC = 32, WG_BK = 32, WG_T = 2;  // 16x16 threads, 2x2 micro-tile

__global__ void __launch_bounds__(256)
conv_wgrad_simt(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ dy, int64_t dy_ld,
                float* __restrict__ partial, GatherGeom g, int nchunks, int64_t chunk_vox, int n_tiles,
                int c_tiles) {
  __shared__ __align__(16) float Ds[WG_BK][WG_BN + 4];  // dy tile  [voxel][n]
  __shared__ __align__(16) float Xs[WG_BK][WG_BC + 4];  // x tile   [voxel][c]
  const int tid = threadIdx.x;
  const int taps = g.kd * g.kh * g.kw;
  int b = blockIdx.x;
  const int c_tile = b % c_tiles; b /= c_tiles;
  const int n_tile = b % n_tiles; b /= n_tiles;
  const int tap = b % taps;
  const int chunk = b / taps;
  const int tkw = tap % g.kw, tkh = (tap / g.kw) % g.kh, tkd = tap / (g.kw * g.kh);
  const int n0 = n_tile * WG_BN, c0 = c_tile * WG_BC;
  const int64_t nvox_out = (int64_t)g.Do * g.Ho * g.Wo;
  const int64_t v_begin = (int64_t)chunk * chunk_vox;
  const int64_t v_end = min(nvox_out, v_begin + chunk_vox);

  const int tx = tid & 15, ty = tid >> 4;
  float acc[WG_T][WG_T] = {{0.f, 0.f}, {0.f, 0.f}};
  // loader: 32 voxels x 8 float4 = 256 float4 per tile, one per thread
  const int lrow = tid >> 3, lq = (tid & 7) * 4;
  for (int64_t v0 = v_begin; v0 < v_end; v0 += WG_BK) {
    const int64_t v = v0 + lrow;
    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < v_end) {
      if (n0 + lq < g.N) dv = __ldg(reinterpret_cast<const float4*>(dy + v * dy_ld + n0 + lq));
      int64_t t = v;
      const int w = (int)(t % g.Wo); t /= g.Wo;
      const int h = (int)(t % g.Ho);
      const int d = (int)(t / g.Ho);
      const int sd_ = d * g.sd + tkd - g.pd, sh_ = h * g.sh + tkh - g.ph, sw_ = w * g.sw + tkw - g.pw;
      if (sd_ >= 0 && sd_ < g.Di && sh_ >= 0 && sh_ < g.Hi && sw_ >= 0 && sw_ < g.Wi && c0 + lq < g.C)
        xv = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)sd_ * g.Hi + sh_) * g.Wi + sw_) * x_ld + c0 + lq));
    }
    *reinterpret_cast<float4*>(&Ds[lrow][lq]) = dv;
    *reinterpret_cast<float4*>(&Xs[lrow][lq]) = xv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_BK; ++kk) {
      const float a0 = Ds[kk][ty * 2], a1 = Ds[kk][ty * 2 + 1];
      const float b0 = Xs[kk][tx * 2], b1 = Xs[kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  float* dst = partial + (int64_t)chunk * g.N * taps * g.C;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + ty * 2 + i, c = c0 + tx * 2 + j;
      if (n < g.N && c < g.C) dst[((int64_t)n * taps + tap) * g.C + c] = acc[i][j];
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int64_t n, float* __restrict__ dw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nchunks; ++k) s += partial[(int64_t)k * n + i];
    dw[i] = s;
  }
}

static int geom_from_api(const dpi_conv_geom* a, GatherGeom& g, bool transposed) {
  if (!a) { set_error("conv: null geometry"); return DPI_ERR_INVALID_ARG; }
  auto okk = [](int k) { return k == 1 || k == 3; };
  if (!okk(a->kd) || !okk(a->kh) || !okk(a->kw) || (a->stride != 1 && a->stride != 2) || a->D < 1 ||
      a->H < 1 || a->W < 1 || a->Cin < 4 || a->Cout < 4 || (a->Cin & 3) || (a->Cout & 3)) {
    set_error("conv: unsupported geometry D=%d H=%d W=%d Cin=%d Cout=%d k=(%d,%d,%d) stride=%d", a->D, a->H,
              a->W, a->Cin, a->Cout, a->kd, a->kh, a->kw, a->stride);
    return DPI_ERR_INVALID_ARG;
  }
  const int sd = a->kd > 1 ? a->stride : 1, sh = a->kh > 1 ? a->stride : 1, sw = a->kw > 1 ? a->stride : 1;
  const int pd = (a->kd - 1) / 2, ph = (a->kh - 1) / 2, pw = (a->kw - 1) / 2;
  const int Do = (a->D + 2 * pd - a->kd) / sd + 1, Ho = (a->H + 2 * ph - a->kh) / sh + 1,
            Wo = (a->W + 2 * pw - a->kw) / sw + 1;
  g.kd = a->kd; g.kh = a->kh; g.kw = a->kw;
  g.sd = sd; g.sh = sh; g.sw = sw; g.pd = pd; g.ph = ph; g.pw = pw;
  g.transposed = transposed ? 1 : 0;
  g.thin_c = 0;
  g.wC = 0;
  if (!transposed) {
    g.Di = a->D; g.Hi = a->H; g.Wi = a->W; g.Do = Do; g.Ho = Ho; g.Wo = Wo; g.C = a->Cin; g.N = a->Cout;
  } else {
    g.Di = Do; g.Hi = Ho; g.Wi = Wo; g.Do = a->D; g.Ho = a->H; g.Wo = a->W; g.C = a->Cout; g.N = a->Cin;
  }
  return DPI_OK;
}

template <int BM, int BN, int TM, int TN>
static int launch_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                         int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st) {
  const int64_t nvox = (int64_t)g.Do * g.Ho * g.Wo;
  dim3 grid((unsigned)ceil_div64(nvox, BM), (unsigned)((g.N + BN - 1) / BN));
  conv_gather_simt<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(in, in_ld, Wp, bias, out, out_ld, g,
                                                                         accumulate);
  return check_launch("conv_gather_simt");
}

int conv_gather_simt_dispatch(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                              int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st) {
  if (g.N <= 8) return launch_gather<128, 8, 4, 2>(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
  if (g.N <= 16) return launch_gather<128, 16, 4, 4>(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
  if (g.N <= 32) return launch_gather<128, 32, 8, 4>(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
  return launch_gather<128, 64, 8, 8>(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
}

struct WgradPlan { int nchunks; int64_t chunk_vox; int n_tiles, c_tiles; };
static WgradPlan wgrad_plan(const GatherGeom& g) {
  WgradPlan p;
  const int taps = g.kd * g.kh * g.kw;
  p.n_tiles = (g.N + WG_BN - 1) / WG_BN;
  p.c_tiles = (g.C + WG_BC - 1) / WG_BC;
  const int64_t nvox = (int64_t)g.Do * g.Ho * g.Wo;
  const int64_t base = (int64_t)taps * p.n_tiles * p.c_tiles;
  int64_t want = (148 * 8 + base - 1) / base;  // ~8 CTAs per SM in total
  const int64_t max_chunks = (nvox + 255) / 256;
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  p.chunk_vox = ((nvox + want - 1) / want + WG_BK - 1) / WG_BK * WG_BK;
  p.nchunks = (int)((nvox + p.chunk_vox - 1) / p.chunk_vox);
  return p;
}

int geom_public(const dpi_conv_geom* a, GatherGeom& g, bool transposed) { return geom_from_api(a, g, transposed); }

}  // namespace dpi

using namespace dpi;

namespace dpi {
StatsRequest*& stats_request() {
  static thread_local StatsRequest* rq = nullptr;
  return rq;
}
}  // namespace dpi

extern "C" {

int dpi_conv_fwd_stats(const float* x, int64_t x_ld, const float* w, const float* bias, float* y, int64_t y_ld,
                       const dpi_conv_geom* geom, int precision, void* stats_ws, void* stream) {
  if (!stats_ws) return dpi_conv_fwd(x, x_ld, w, bias, y, y_ld, geom, precision, stream);
  StatsRequest rq{stats_ws, false};
  stats_request() = &rq;
  int rc = dpi_conv_fwd(x, x_ld, w, bias, y, y_ld, geom, precision, stream);
  stats_request() = nullptr;
  if (rc || rq.done) return rc;
  // the kernel that ran cannot emit statistics itself: one streaming pass over y
  GatherGeom g;
  rc = geom_from_api(geom, g, false);
  if (rc) return rc;
  return dpi_channel_stats(y, y_ld, (int64_t)g.Do * g.Ho * g.Wo, g.N, stats_ws, stream);
}

int dpi_conv_fwd(const float* x, int64_t x_ld, const float* w, const float* bias, float* y, int64_t y_ld,
                 const dpi_conv_geom* geom, int precision, void* stream) {
  GatherGeom g;
  int rc = geom_from_api(geom, g, false);
  if (rc) return rc;
  DPI_REQUIRE(x && w && y && aligned16(x) && aligned16(w) && aligned16(y) && !(x_ld & 3) && !(y_ld & 3) &&
                  x_ld >= g.C && y_ld >= g.N,
              "dpi_conv_fwd: pointers must be 16B aligned and pitches multiples of 4 covering the channels");
  if (precision == DPI_PREC_TF32) {
    rc = conv_tc_gather(x, x_ld, w, bias, y, y_ld, g, 0, (cudaStream_t)stream);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  return conv_gather_simt_dispatch(x, x_ld, w, bias, y, y_ld, g, 0, (cudaStream_t)stream);
}

int dpi_conv_dgrad(const float* dy, int64_t dy_ld, const float* wt, float* dx, int64_t dx_ld,
                   const dpi_conv_geom* geom, int accumulate, int precision, void* stream) {
  GatherGeom g;
  int rc = geom_from_api(geom, g, true);
  if (rc) return rc;
  DPI_REQUIRE(dy && wt && dx && aligned16(dy) && aligned16(wt) && aligned16(dx) && !(dy_ld & 3) &&
                  !(dx_ld & 3) && dy_ld >= g.C && dx_ld >= g.N,
              "dpi_conv_dgrad: pointers must be 16B aligned and pitches multiples of 4 covering the channels");
  if (precision == DPI_PREC_TF32) {
    rc = conv_tc_gather(dy, dy_ld, wt, nullptr, dx, dx_ld, g, accumulate, (cudaStream_t)stream);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  return conv_gather_simt_dispatch(dy, dy_ld, wt, nullptr, dx, dx_ld, g, accumulate, (cudaStream_t)stream);
}

int dpi_conv_dgrad_fused(const float* dy, int64_t dy_ld, const float* wt, float* dx, int64_t dx_ld,
                         const dpi_conv_geom* geom, int full_tap_channels, int accumulate, int precision, void* stream) {
  GatherGeom g;
  int rc = geom_from_api(geom, g, true);
  if (rc) return rc;
  DPI_REQUIRE(dy && wt && dx && aligned16(dy) && aligned16(wt) && aligned16(dx) && !(dy_ld & 3) &&
                  !(dx_ld & 3) && dy_ld >= g.C && dx_ld >= g.N,
              "dpi_conv_dgrad_fused: pointers must be 16B aligned and pitches multiples of 4 covering the channels");
  DPI_REQUIRE(full_tap_channels >= 4 && !(full_tap_channels & 3) && full_tap_channels < g.C && g.kh == 3 && g.kw == 3 &&
                  g.sd == 1 && g.sh == 1 && g.sw == 1,
              "dpi_conv_dgrad_fused: needs a stride-1 3x3(x3) geometry and 4 <= full_tap_channels < Cout");
  g.thin_c = full_tap_channels;
  if (precision == DPI_PREC_TF32) {
    rc = conv_tc_gather(dy, dy_ld, wt, nullptr, dx, dx_ld, g, accumulate, (cudaStream_t)stream);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  return conv_gather_simt_dispatch(dy, dy_ld, wt, nullptr, dx, dx_ld, g, accumulate, (cudaStream_t)stream);
}

int dpi_conv_dgrad_fused_supported(const dpi_conv_geom* geom, int full_tap_channels) {
  GatherGeom g;
  if (geom_from_api(geom, g, true)) return 0;
  if (full_tap_channels < 4 || (full_tap_channels & 3) || full_tap_channels >= g.C || g.kh != 3 || g.kw != 3 || g.sd != 1 ||
      g.sh != 1 || g.sw != 1)
    return 0;
  g.thin_c = full_tap_channels;
  return conv_tc_march_supported(g);
}

int64_t dpi_conv_wgrad_workspace_bytes(const dpi_conv_geom* geom) {
  GatherGeom g;
  if (geom_from_api(geom, g, false)) return -1;
  WgradPlan p = wgrad_plan(g);
  const int64_t simt = (int64_t)p.nchunks * g.N * g.kd * g.kh * g.kw * g.C * (int64_t)sizeof(float);
  const int64_t tc = conv_tc_wgrad_workspace_bytes(g);
  return simt > tc ? simt : tc;
}

int dpi_conv_wgrad(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* dw,
                   const dpi_conv_geom* geom, void* workspace, int64_t workspace_bytes, int precision,
                   void* stream) {
  GatherGeom g;
  int rc = geom_from_api(geom, g, false);
  if (rc) return rc;
  DPI_REQUIRE(x && dy && dw && aligned16(x) && aligned16(dy) && !(x_ld & 3) && !(dy_ld & 3) && x_ld >= g.C &&
                  dy_ld >= g.N,
              "dpi_conv_wgrad: pointers must be 16B aligned and pitches multiples of 4 covering the channels");
  const int taps = g.kd * g.kh * g.kw;
  const int64_t wn = (int64_t)g.N * taps * g.C;
  if (precision == DPI_PREC_TF32 && workspace) {
    int nchunks = 0;
    rc = conv_tc_wgrad(x, x_ld, dy, dy_ld, (float*)workspace, workspace_bytes, g, &nchunks, (cudaStream_t)stream);
    if (rc == DPI_OK) {
      int rb = (int)((wn + 255) / 256);
      if (rb > 148 * 8) rb = 148 * 8;
      wgrad_reduce_kernel<<<rb, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, nchunks, wn, dw);
      return check_launch("wgrad_reduce_kernel");
    }
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  WgradPlan p = wgrad_plan(g);
  if (!workspace || workspace_bytes < (int64_t)p.nchunks * wn * (int64_t)sizeof(float)) {
    set_error("dpi_conv_wgrad: workspace too small (%lld < %lld)", (long long)workspace_bytes,
              (long long)(p.nchunks * wn * sizeof(float)));
    return DPI_ERR_WORKSPACE;
  }
  const int64_t blocks = (int64_t)p.nchunks * taps * p.n_tiles * p.c_tiles;
  conv_wgrad_simt<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, x_ld, dy, dy_ld, (float*)workspace, g,
                                                                     p.nchunks, p.chunk_vox, p.n_tiles, p.c_tiles);
  rc = check_launch("conv_wgrad_simt");
  if (rc) return rc;
  int rb = (int)((wn + 255) / 256);
  if (rb > 148 * 8) rb = 148 * 8;
  wgrad_reduce_kernel<<<rb, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, p.nchunks, wn, dw);
  return check_launch("wgrad_reduce_kernel");
}

}  // extern "C"
