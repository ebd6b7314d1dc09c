// tcgen05 / TMA implicit-GEMM convolution (DPI_PREC_TF32) -- placeholder until the kernel lands.
#include "conv_geom.cuh"
namespace dpi {
int conv_tc_gather(const float*, int64_t, const float*, const float*, float*, int64_t, const GatherGeom&, int,
                   cudaStream_t) {
  return DPI_ERR_UNSUPPORTED;
}
}  // namespace dpi
