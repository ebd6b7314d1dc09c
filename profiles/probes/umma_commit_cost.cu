// Probe: what does a tcgen05.commit cost the ISSUING thread?  The packed march spends ~1690 clk per plane-step inside its
// elected issue region whether it issues 9 or 18 MMAs (DPI_TC_MARCH_PROF=1), and a plane-step without any MMA still costs
// ~870 clk (profiles/r1_march_bottleneck_isolation.txt).  One warp (elected lane) runs `iters` rounds of: n MMAs (N = 48),
// then c commits to c different mbarriers; reported: clk per round.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}

struct Cfg { int n_mma, n_commit, iters, wait_each; };

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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [&](uint32_t addr) {
      uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)1 << 16; d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32; d |= (uint64_t)2 << 61; d |= (uint64_t)1 << 46;
      return d;
    };
    const uint64_t ad = desc(a_base), bd = desc(b_base);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      if (elect_one()) {
        for (int j = 0; j < c.n_mma; ++j)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tmem), "l"(ad + (uint64_t)(2 * (j & 3))), "l"(bd + (uint64_t)(2 * (j & 3))), "r"(idesc), "r"(1u) : "memory");
        for (int k = 0; k < c.n_commit; ++k)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8u * k) : "memory");
      }
      __syncwarp();
      if (c.wait_each) { while (!mbar_try(bar0, phase)) {} phase ^= 1u; }
    }
    const long long t1 = clock64();
    // drain before leaving
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 24u) : "memory");
    __syncwarp();
    while (!mbar_try(bar0 + 24u, (uint32_t)((c.n_commit > 3 ? c.iters : 0) & 1))) {}
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  std::vector<long long> h(148);
  const Cfg cfgs[] = {{0, 1, 2000, 0}, {0, 2, 2000, 0}, {9, 0, 2000, 0}, {9, 1, 2000, 0}, {9, 2, 2000, 0}, {18, 2, 2000, 0},
                      {36, 2, 2000, 0}, {9, 1, 2000, 1}, {18, 1, 2000, 1}, {1, 1, 2000, 1}, {0, 1, 2000, 1}};
  for (const Cfg& c : cfgs) {
    rate<<<148, 128, 68 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("per round: %2d MMAs (N=48) + %d commits%s : %7.1f clk\n", c.n_mma, c.n_commit, c.wait_each ? " + wait for the first commit's barrier" : "",
           (double)mx / c.iters);
  }
  return 0;
}
