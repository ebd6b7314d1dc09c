// Probe: how does a 4-D TMA box over (channel, d, w, h) - a NON-innermost box dimension (d) whose global stride is the
// largest - land in shared memory?  Used to design the kd-packed weight-gradient kernel (conv_tc_wgrad_march.cu).
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_box_probe tma_box_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int nfloats, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 32768;
  float* s = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = -7.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nfloats * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(base), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  for (int spin = 0; !ok && spin < (1 << 22); ++spin)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < nfloats + 64; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)fp;
  const int N = 8, D = 6, H = 16, W = 8, ld = 8;
  std::vector<float> h((size_t)D * H * W * ld);
  for (int d = 0; d < D; ++d) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < ld; ++c)
    h[(((size_t)d * H + y) * W + x) * ld + c] = c + 10.f * d + 100.f * x + 1000.f * y;      // value encodes (y, x, d, c)
  float *g, *out; cudaMalloc(&g, h.size() * 4); cudaMemcpy(g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 65536);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const int nb = 8, pb = 4, TW = 8, HY = 4;
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {N, D, W, H};
    cuuint64_t strides[3] = {(cuuint64_t)H * W * ld * 4, (cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4};
    cuuint32_t box[4] = {nb, pb, TW, HY};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mode ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d encode rc=%d\n", mode, (int)r);
    const int nfl = nb * pb * TW * HY;
    probe<<<1, 128, 40 * 1024>>>(tm, out, nfl, 0, 1, 0, 2);          // channels 0.., planes 1..4, w 0.., h 2..
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    std::vector<float> o(nfl + 64);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    for (int row = 0; row < 12; ++row) {            // 128-byte rows
      printf("  smem row %2d:", row);
      for (int i = 0; i < 32; ++i) printf(" %5.0f", o[row * 32 + i]);
      printf("\n");
    }
  }
  {
    // mode 2: a full 128-byte inner box (32 channels) under SWIZZLE_128B_ATOM_32B: the complete chunk permutation
    const int ld2 = 32;
    std::vector<float> h2((size_t)4 * 16 * ld2);
    for (int r = 0; r < 64; ++r) for (int c = 0; c < ld2; ++c) h2[(size_t)r * ld2 + c] = c + 100.f * r;
    float* g2; cudaMalloc(&g2, h2.size() * 4); cudaMemcpy(g2, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[4] = {32, 16, 4, 1};
    cuuint64_t strides[3] = {(cuuint64_t)ld2 * 4, (cuuint64_t)16 * ld2 * 4, (cuuint64_t)64 * ld2 * 4};
    cuuint32_t box[4] = {32, 16, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g2, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode 2 encode rc=%d\n", (int)r);
    probe<<<1, 128, 40 * 1024>>>(tm, out, 32 * 16, 0, 0, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    std::vector<float> o(32 * 16 + 64);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    for (int row = 0; row < 16; ++row) {
      printf("  smem row %2d:", row);
      for (int i = 0; i < 32; ++i) printf(" %5.0f", o[row * 32 + i]);
      printf("\n");
    }
  }
  return 0;
}
