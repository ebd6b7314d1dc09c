// Probe for DESIGN.md §10 item 1 (NOT part of the product; written at the end of round 1, compile-checked only —
// run it first thing in round 2):  how should the epilogue of the march conv kernels write its 16 x 8-voxel x N-channel
// tile of a channels-last fp32 tensor [D][H][W][ld]?
//
//   mode 0  per-thread stores (today): thread = voxel, float4 stores of its N channels at pitch ld
//   mode 1  the same, accumulating (dgrad): the old values are read first
//   mode 2  tile staged in shared memory [16][8][N], one TMA tensor store (cp.async.bulk.tensor.4d ... bulk_group)
//   mode 3  the same with the TMA reduction (cp.reduce.async.bulk.tensor.4d ... .add.f32): accumulate without reading
//
// Every CTA (128 threads, persistent grid) walks tiles in (d, th, tw) order like the march kernel's epilogue warps; the
// values are synthetic.  Prints GB/s of useful bytes (N of ld channels written; mode 1 also counts the read).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o epilogue_store_probe epilogue_store_probe.cu -lcuda
//   ./epilogue_store_probe            (D H W = 256 128 128; N / ld pairs of the full-resolution layers)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int TH = 16, TW = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) store_probe(const __grid_constant__ CUtensorMap tmap, float* __restrict__ out, int D,
                                                   int H, int W, int N, int64_t ld, int mode) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);                 // [2][TH*TW][N]: double-buffered tile
  const int tiles_w = W / TW, tiles_h = H / TH;
  const int64_t n_tiles = (int64_t)D * tiles_h * tiles_w;
  const int row = threadIdx.x;                                        // voxel of the tile: h = row >> 3, w = row & 7
  int buf = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tw = (int)(t % tiles_w), th = (int)((t / tiles_w) % tiles_h), d = (int)(t / ((int64_t)tiles_w * tiles_h));
    const int h = th * TH + (row >> 3), w = tw * TW + (row & 7);
    const float base = (float)(t & 1023) * 1e-3f + (float)row;
    if (mode <= 1) {
      float* o = out + (((int64_t)d * H + h) * W + w) * ld;
      for (int n = 0; n < N; n += 4) {
        float4 v = make_float4(base + n, base + n + 1, base + n + 2, base + n + 3);
        if (mode == 1) {
          const float4 old = *reinterpret_cast<const float4*>(o + n);
          v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
        }
        *reinterpret_cast<float4*>(o + n) = v;
      }
    } else {
      float* s = stage + (size_t)buf * TH * TW * N + (size_t)row * N;
      // the buffer we are about to overwrite was handed to the TMA two tiles ago: wait until it has been READ
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
      for (int n = 0; n < N; n += 4)
        *reinterpret_cast<float4*>(s + n) = make_float4(base + n, base + n + 1, base + n + 2, base + n + 3);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t src = smem_u32(stage + (size_t)buf * TH * TW * N);
        if (mode == 2)
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(src), "r"(0), "r"(tw * TW), "r"(th * TH), "r"(d)
                       : "memory");
        else
          asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(src), "r"(0), "r"(tw * TW), "r"(th * TH), "r"(d)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      buf ^= 1;
    }
  }
  if (mode >= 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int D = argc > 3 ? atoi(argv[1]) : 256, H = argc > 3 ? atoi(argv[2]) : 128, W = argc > 3 ? atoi(argv[3]) : 128;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
    printf("cuTensorMapEncodeTiled unavailable\n");
    return 1;
  }
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fp);
  const int cases[][2] = {{4, 4}, {8, 8}, {16, 16}, {28, 28}, {72, 72}, {16, 72}, {56, 72}};   // {N, ld}
  const int64_t nvox = (int64_t)D * H * W;
  float* buf = nullptr;
  cudaMalloc(&buf, nvox * 72 * sizeof(float));
  cudaMemset(buf, 0, nvox * 72 * sizeof(float));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TH * TW * 72 * 4);
  for (auto& c : cases) {
    const int N = c[0];
    const int64_t ld = c[1];
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)N, TW, TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("N=%d ld=%lld: cuTensorMapEncodeTiled failed (%d)\n", N, (long long)ld, (int)r);
      continue;
    }
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
      for (int mode = 0; mode < 4; ++mode) {
        const size_t smem = mode >= 2 ? (size_t)2 * TH * TW * N * 4 : 0;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEventRecord(e0);
          store_probe<<<sms * ctas_per_sm, 128, smem>>>(tmap, buf, D, H, W, N, ld, mode);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms = 0.f;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep && ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        const double bytes = (double)nvox * N * 4 * ((mode == 1) ? 2 : 1);
        printf("N=%2d ld=%2lld ctas/sm=%d mode=%d  %8.1f us  %6.2f TB/s of useful bytes%s\n", N, (long long)ld, ctas_per_sm, mode,
               best * 1e3, bytes / (best * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
      }
    }
  }
  cudaFree(buf);
  return 0;
}
