// Probe: issue rate of tcgen05.mma kind::tf32 for the operand shapes / majors the conv kernels use.
// One CTA per SM (grid 148), each issuing `iters` back-to-back MMAs on resident smem operands; cycles per MMA
// from clock64 around the issue loop + final commit wait.
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

struct Cfg { int M, N, a_mn, b_mn, iters, a_stride_rows; };

__global__ void __launch_bounds__(128) rate(Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 96 * 1024, bar = base + 160 * 1024, slot = bar + 8;
  // zero operands
  for (uint32_t i = threadIdx.x * 4; i < 160 * 1024; i += 128 * 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + i), "r"(0u));
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    auto desc = [&](uint32_t addr, int mn, uint32_t sbo_k) {
      uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
      if (mn) { d |= (uint64_t)((16384 >> 4) & 0x3FFF) << 16; d |= (uint64_t)(512 >> 4) << 32; d |= (uint64_t)1 << 61; }
      else    { d |= (uint64_t)1 << 16; d |= (uint64_t)((sbo_k >> 4) & 0x3FFF) << 32; d |= (uint64_t)2 << 61; }
      d |= (uint64_t)1 << 46;
      return d;
    };
    uint64_t ads[8], bds[8];
    for (int j = 0; j < 8; ++j) {
      ads[j] = desc(a_base + (c.a_mn ? j * 1024u : (j & 3) * 32u + (j >> 2) * 1280u), c.a_mn, c.a_stride_rows ? 1280 : 1024);
      bds[j] = desc(b_base + (c.b_mn ? j * 1024u : (j & 3) * 32u), c.b_mn, 1024);
    }
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem), "l"(ads[j]), "l"(bds[j]), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    while (!mbar_try(bar, 0)) {}
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  std::vector<long long> h(148);
  const Cfg cfgs[] = {
      {128, 16, 0, 0, 4000, 0},  {128, 16, 0, 0, 4000, 1}, {128, 32, 0, 0, 4000, 0}, {128, 64, 0, 0, 4000, 0},
      {128, 128, 0, 0, 4000, 0}, {128, 256, 0, 0, 4000, 0}, {64, 16, 0, 0, 4000, 0},  {64, 64, 0, 0, 4000, 0},
      {64, 256, 0, 0, 4000, 0},  {128, 16, 1, 1, 4000, 0}, {128, 32, 1, 1, 4000, 0}, {128, 64, 1, 1, 4000, 0},
      {64, 16, 1, 1, 4000, 0},   {64, 64, 1, 1, 4000, 0},  {128, 16, 1, 0, 4000, 0}, {128, 16, 0, 1, 4000, 0},
      {128, 8, 0, 0, 4000, 0},   {64, 8, 0, 0, 4000, 0},   {64, 8, 1, 1, 4000, 0},
  };
  for (const Cfg& c : cfgs) {
    for (int grid : {148}) {
      rate<<<grid, 128, 162 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("M=%d N=%d aMN=%d bMN=%d: error %s\n", c.M, c.N, c.a_mn, c.b_mn, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("M=%3d N=%3d A=%s B=%s shiftA=%d grid=%3d : %.1f clk/MMA  (%.0f MAC/clk/SM)\n", c.M, c.N, c.a_mn ? "MN" : "K ", c.b_mn ? "MN" : "K ",
             c.a_stride_rows, grid, (double)mx / c.iters, (double)c.M * c.N * 8 * c.iters / mx);
    }
  }
  return 0;
}
