// Probe: is the ~39 clk per tcgen05.mma floor of small-N kind::tf32 MMAs a property of the ISSUING THREAD or of the SM?
// ncu shows the tensor pipe only 11.5 % active in the 4 -> 8 march kernel (27 MMAs of N = 16 per plane), i.e. ~8 clk of pipe
// time per MMA, while a single issuing thread cannot issue faster than one per ~39 clk (umma_rate4.cu).  Here W = 1..4
// warps of ONE CTA issue `iters` MMAs each (whole-warp control flow, elected lane), every warp into its own TMEM columns,
// and W = 1..2 CTAs per SM do the same; reported: clk per MMA per issuer and aggregate MMAs per 1000 clk per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct Cfg { int N, warps, iters, tmem_cols; };

__global__ void __launch_bounds__(128) rate(Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 32 * 1024, bar0 = base + 64 * 1024, slot = bar0 + 64;
  for (uint32_t i = threadIdx.x * 4; i < 64 * 1024; i += 128 * 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + i), "r"(0u));
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * w));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"((uint32_t)c.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const int warp = threadIdx.x >> 5;
  long long dt = 0;
  if (warp < c.warps) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [&](uint32_t addr) {
      uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)1 << 16; d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32; d |= (uint64_t)2 << 61; d |= (uint64_t)1 << 46;
      return d;
    };
    uint64_t ads[4], bds[4];
    for (int j = 0; j < 4; ++j) { ads[j] = desc(a_base + j * 32u); bds[j] = desc(b_base + j * 32u); }
    const uint32_t dcol = tmem + (uint32_t)(warp * (c.tmem_cols / 4));
    const uint32_t bar = bar0 + 8u * warp;
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(dcol), "l"(ads[j]), "l"(bds[j]), "r"(idesc), "r"(1u) : "memory");
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    __syncwarp();
    while (!mbar_try(bar, 0)) {}
    dt = clock64() - t0;
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = dt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)c.tmem_cols) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 296 * 4 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  std::vector<long long> h(296 * 4);
  for (int N : {16, 32, 48, 96}) {
    for (int ctas : {1, 2}) {
      for (int warps : {1, 2, 3, 4}) {
        Cfg c{N, warps, 4000, 256};
        if (warps * N > 256 / 1 && N * 4 > 256) c.tmem_cols = 512;
        if (ctas == 2 && c.tmem_cols > 256) continue;
        const int grid = 148 * ctas;
        cudaMemset(d, 0, 296 * 4 * 8);
        rate<<<grid, 128, 68 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d warps=%d ctas=%d: error %s\n", N, warps, ctas, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, grid * 4 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid * 4; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("N=%3d CTAs/SM=%d issuing warps/CTA=%d : %6.1f clk per MMA per issuer, %6.1f MMAs per 1000 clk per SM\n", N, ctas, warps,
               (double)mx / c.iters, 1000.0 * c.iters * warps * ctas / mx);
      }
    }
  }
  return 0;
}
