// Geometry of the "gather convolution" shared by the CUDA-core and tcgen05 conv kernels.
#pragma once
#include "dpi_common.cuh"

namespace dpi {

struct GatherGeom {
  int Di, Hi, Wi;      // spatial size of `in`
  int Do, Ho, Wo;      // spatial size of `out`
  int C, N;            // reduction channels, output channels
  int kd, kh, kw;
  int sd, sh, sw;      // stride per axis
  int pd, ph, pw;      // padding per axis
  int transposed;      // 0 forward, 1 dgrad
};


// CUDA-core path (conv_simt.cu)
int conv_gather_simt_dispatch(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                              int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);
// tcgen05 path (conv_tc.cu); returns DPI_ERR_UNSUPPORTED when the shape is not covered
int conv_tc_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                   int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);

}  // namespace dpi
