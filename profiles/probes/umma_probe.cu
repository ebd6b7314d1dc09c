// Probe: how do tcgen05 K-major / MN-major SW128 descriptors address a TMA-written tile when the start address is
// shifted by whole rows and the 8-row group stride (SBO) is not a multiple of 1024 B?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) { for (uint32_t s = 0; !mbar_try(bar, ph); ++s) if (s > (1u << 24)) __trap(); }

struct Variant { int shift_rows; int sbo; int base_off; int mn_major; int lbo; int layout; };

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, Variant v, float* out /*128x32*/) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;                 // 256 rows x 128 B = 32 KB
  const uint32_t b_base = base + 32768;         // identity 32x32, SW128 K-major, 4 KB
  const uint32_t bar0 = base + 32768 + 4096, bar1 = bar0 + 8, slot = bar0 + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B = identity (n,k) in SW128 K-major layout
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    const uint32_t addr = b_base + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
    const float val = (n == k) ? 1.f : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(val));
  }
  if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (tid == 0) {
    mbar_expect_tx(bar0, 256 * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(a_base), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar0), "r"(0), "r"(0) : "memory");
    mbar_wait(bar0, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(v.mn_major ? 1 : 0) << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = 0, bd = 0;
      uint32_t a_start;
      if (!v.mn_major) {
        a_start = a_base + v.shift_rows * 128 + 32 * k;          // K-major: advance 32 B inside the swizzle atom
        ad |= (uint64_t)((a_start >> 4) & 0x3FFF);
        ad |= (uint64_t)1 << 16;
        ad |= (uint64_t)((v.sbo >> 4) & 0x3FFF) << 32;
      } else {
        // MN-major A: M = 128 "channels" = 4 blocks of 32 floats at LBO; K = 8 rows per MMA (one 1024 B atom);
        // K step k advances 8 rows.
        a_start = a_base + (v.shift_rows + 8 * k) * 128;
        ad |= (uint64_t)((a_start >> 4) & 0x3FFF);
        ad |= (uint64_t)((v.lbo >> 4) & 0x3FFF) << 16;
        ad |= (uint64_t)((v.sbo >> 4) & 0x3FFF) << 32;
      }
      ad |= (uint64_t)1 << 46;
      ad |= (uint64_t)(v.base_off & 7) << 49;
      ad |= (uint64_t)(v.layout & 7) << 61;
      const uint32_t b_start = b_base + 32 * k;
      bd |= (uint64_t)((b_start >> 4) & 0x3FFF);
      bd |= (uint64_t)1 << 16;
      bd |= (uint64_t)(1024 >> 4) << 32;
      bd |= (uint64_t)1 << 46;
      bd |= (uint64_t)2 << 61;
      const uint32_t acc = k > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
  }
  mbar_wait(bar1, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = tid;
  for (int c = 0; c < 32; c += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[row * 32 + c + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int R = 256;
  std::vector<float> h(R * 32);
  float *dX, *dOut;
  cudaMalloc(&dX, R * 32 * 4);
  cudaMalloc(&dOut, 128 * 32 * 4);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap tm;
  cuuint64_t dims[2] = {32, (cuuint64_t)R};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, 256};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  CUtensorMap tm2;
  r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode2 failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  std::vector<float> o(128 * 32);
  // ---- K-major variants: rows (voxels) = M
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < R * 32; ++i) h[i] = pass == 0 ? (float)(i / 32) : (float)(i % 32);
    cudaMemcpy(dX, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int sbo : {1024, 1280, 1152}) {
      for (int shift = 0; shift < 4; ++shift) {
        for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
          Variant v{shift, sbo, bo_mode ? (shift & 7) : 0, 0, 0, 2};
          cudaMemset(dOut, 0xff, 128 * 32 * 4);
          probe<<<1, 128, 40 * 1024>>>(tm, v, dOut);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("K-major sbo=%d shift=%d bo=%d: CUDA error %s\n", sbo, shift, v.base_off, cudaGetErrorString(e)); return 2; }
          cudaMemcpy(o.data(), dOut, o.size() * 4, cudaMemcpyDeviceToHost);
          int ok = 0;
          for (int m = 0; m < 128; ++m) {
            const int exp_row = shift + (m / 8) * (sbo / 128) + (m % 8);
            bool good = true;
            for (int n = 0; n < 32; ++n) good &= (o[m * 32 + n] == (pass == 0 ? (float)exp_row : (float)n));
            ok += good;
          }
          printf("Kmajor %s sbo=%4d shift=%d base_off=%d : %3d/128 rows as expected", pass == 0 ? "rowid" : "colid", sbo, shift, v.base_off, ok);
          if (ok != 128) {
            printf("   e.g. m=0..11 n=0,4:");
            for (int m = 0; m < 12; ++m) printf(" (%g,%g)", o[m * 32], o[m * 32 + 4]);
          }
          printf("\n");
        }
      }
    }
  }
  // ---- MN-major variants: A[m = channel][k = row]; with B = identity over k (32), D[m][n] = X[row n (+shift...)][channel m]
  // here the TMA tile is [rows][32 ch]; M=128 needs 4 channel blocks at LBO: we reuse the same 32 channels by pointing
  // LBO at a row offset (block j = rows shifted by lbo/128), so D[32*j + c][n] = X[shift + lbo_rows*j + n][c]
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < R * 32; ++i) h[i] = pass == 0 ? (float)(i / 32) : (float)(i % 32);
    cudaMemcpy(dX, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int cfg = 0; cfg < 4; ++cfg)
    for (int lbo : {4096, 5120}) {
      for (int shift = 0; shift < 3; ++shift) {
        for (int bo_mode = 0; bo_mode < 1; ++bo_mode) {
          // cfg 0: ATOM_32B tile + layout 1 + SBO 512 ; cfg 1: ATOM_32B + layout 1 + SBO 1024 ;
          // cfg 2: plain SW128 tile + layout 2 + SBO 1024 (known not to work for tf32?) ; cfg 3: ATOM_32B + layout 1, LBO/SBO swapped
          Variant v{shift, cfg == 1 ? 1024 : 512, 0, 1, lbo, cfg == 2 ? 2 : 1};
          if (cfg == 2) v.sbo = 1024;
          if (cfg == 3) { v.sbo = lbo; v.lbo = 512; }
          printf("cfg%d ", cfg);
          cudaMemset(dOut, 0xff, 128 * 32 * 4);
          probe<<<1, 128, 40 * 1024>>>(cfg == 2 ? tm : tm2, v, dOut);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("MN-major lbo=%d shift=%d: CUDA error %s\n", lbo, shift, cudaGetErrorString(e)); return 2; }
          cudaMemcpy(o.data(), dOut, o.size() * 4, cudaMemcpyDeviceToHost);
          int ok = 0;
          for (int m = 0; m < 128; ++m) {
            const int j = m / 32, c = m % 32;
            bool good = true;
            for (int n = 0; n < 32; ++n) {
              const int exp_row = shift + (lbo / 128) * j + n;
              good &= (o[m * 32 + n] == (pass == 0 ? (float)exp_row : (float)c));
            }
            ok += good;
          }
          printf("MNmajor %s lbo=%4d shift=%d base_off=%d : %3d/128 rows as expected", pass == 0 ? "rowid" : "colid", lbo, shift, v.base_off, ok);
          if (ok != 128) {
            printf("   e.g. m=0,1,33 n=0..9:");
            for (int m : {0, 1, 33}) { printf(" |"); for (int n = 0; n < 10; ++n) printf(" %g", o[m * 32 + n]); }
          }
          printf("\n");
        }
      }
    }
  }
  return 0;
}
