// Geometry of the "gather convolution" shared by the CUDA-core and tcgen05 conv kernels.
#pragma once
#include "dpi_common.cuh"

#include <stdlib.h>

namespace dpi {

// Dynamic shared memory a persistent tcgen05 conv CTA may take (DPI_TC_SMEM_KB for the forward / data-gradient kernels,
// DPI_TC_WGRAD_SMEM_KB for the weight-gradient kernels; at most 227 KB).  What a conv CTA leaves free decides whether the
// CTAs of a streaming kernel with static shared memory (BatchNorm statistics / backward reduce: 16 KB + 1 KB) can share
// the SM with it, i.e. whether HBM-bound work of another lane really overlaps the tensor-bound conv.
inline int tc_smem_budget(bool wgrad) {
  static const int v[2] = {
      [] { const char* e = getenv("DPI_TC_SMEM_KB"); const int kb = e ? atoi(e) : 227; return (kb < 96 || kb > 227 ? 227 : kb) * 1024; }(),
      [] { const char* e = getenv("DPI_TC_WGRAD_SMEM_KB"); const int kb = e ? atoi(e) : 227; return (kb < 96 || kb > 227 ? 227 : kb) * 1024; }()};
  return v[wgrad ? 1 : 0];
}

struct GatherGeom {
  int Di, Hi, Wi;      // spatial size of `in`
  int Do, Ho, Wo;      // spatial size of `out`
  int C, N;            // reduction channels, output channels
  int kd, kh, kw;
  int sd, sh, sw;      // stride per axis
  int pd, ph, pw;      // padding per axis
  int transposed;      // 0 forward, 1 dgrad
  int thin_c;          // > 0: only the first thin_c reduction channels carry all taps, the others are non-zero at the
                       // centre tap only (fused data gradient of a 3x3(x3) conv and the 1x1 conv that shares its
                       // input, dpi_conv_dgrad_fused).  A HINT: the weights are zero there, skipping is optional.
  int wC;              // > 0: C is a sub-range of the reduction channels - the packed weights Wp[n][tap][c] have this channel
                       // pitch (conv_tc_march_gather splits the reduction of wide-input, narrow-output convs in two)
};


// Fused BatchNorm statistics: dpi_conv_fwd_stats parks a request here (thread-local) while it dispatches the conv; a
// kernel that can emit the per-CTA partial rows itself (conv_tc_march.cu) takes it and sets `done`, otherwise the
// caller runs the separate statistics pass over y afterwards.
struct StatsRequest {
  void* ws;
  bool done;
};
StatsRequest*& stats_request();

// CUDA-core path (conv_simt.cu)
int conv_gather_simt_dispatch(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                              int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);
// tcgen05 path (conv_tc.cu); returns DPI_ERR_UNSUPPORTED when the shape is not covered
int conv_tc_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                   int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);

// tcgen05 3x3(x3) conv with shared-memory halo reuse (conv_tc_halo.cu)
int conv_tc_halo_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                        int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);
// persistent column-marching variant with resident weights (conv_tc_march.cu); UNSUPPORTED when the packed weights
// do not fit in shared memory next to three plane stages
int conv_tc_march_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                         int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);
// 1x1(x1) convs through the same persistent pipeline (bare tiles, one tap)
int conv_tc_march_1x1(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                      const GatherGeom& g, int accumulate, cudaStream_t st);
// data gradient of the stride-2 3x3(x3) convs: one march per output parity class (conv_tc_march.cu)
int conv_tc_march_dgrad_s2(const float* dy, int64_t dy_ld, const float* Wt, float* dx, int64_t dx_ld,
                           const GatherGeom& g, int accumulate, cudaStream_t st);
// role-swapped kernel for four output channels (conv_tc_swap.cu): (tap, cout) pairs as the M rows, the voxels of an
// input halo tile as the N columns, per-voxel gather of the 27 contributions in the epilogue
int conv_tc_swap_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                        const GatherGeom& g, int accumulate, cudaStream_t st);
int conv_tc_swap_supported(const GatherGeom& g);
// 1 when conv_tc_march_gather would take this problem (resident weights fit) - the kernels that honour thin_c
int conv_tc_march_supported(const GatherGeom& g);
// tcgen05 weight gradient (conv_tc_wgrad.cu): writes nchunks partial slabs [N][taps][C] into `partial`
int64_t conv_tc_wgrad_workspace_bytes(const GatherGeom& g);
int conv_tc_wgrad(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial, int64_t partial_bytes,
                  const GatherGeom& g, int* nchunks_out, cudaStream_t st);
int64_t conv_tc_wgrad_kw_workspace_bytes(const GatherGeom& g);
int conv_tc_wgrad_kw(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial, int64_t partial_bytes,
                     const GatherGeom& g, int* nchunks_out, cudaStream_t st);
int64_t conv_tc_wgrad_march_workspace_bytes(const GatherGeom& g);
int conv_tc_wgrad_march(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial,
                        int64_t partial_bytes, const GatherGeom& g, int* nchunks_out, cudaStream_t st);
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int64_t n, float* __restrict__ dw);

}  // namespace dpi
