// tcgen05 weight gradient of 3x3(x3) stride-1 convolutions: kw packed into M, kh packed into N, marching along d.
//
// conv_tc_wgrad_kw.cu spends, per 128-voxel tile, 9 (kd,kh) x 16 MMAs (N = Cout <= 32) and re-loads the x tile for
// each of the nine (kd,kh) pairs: 200 KB of L2 -> SMEM traffic per tile, which is what bounds the full-resolution
// layers (1254 us each at 256x128x128 whatever the channel count; profiles/r1_op_profile_c.txt).  Here
//   * A (x, MN-major, 32-byte-atom 128-byte swizzle): M block kw = the same 32 channels one voxel further along w
//     (LBO = one 128-byte row), as before;
//   * B (dy, MN-major): the dy tile carries a one-row halo along h and N block j = the same 32 output channels j
//     h-rows further down (LBO = 8 rows = 1024 B): columns [32j, 32j+32) of the accumulator collect tap kh = 2 - j.
//     One MMA (M 128, N 96, K 8 voxels) therefore covers NINE taps (3 kw x 3 kh);
//   * the CTA is persistent and marches along d over its (h,w) tile column: x plane d is loaded ONCE and multiplied
//     with the dy planes d+1, d, d-1 (kd = 0, 1, 2; three accumulators of 96 columns that stay in TMEM for the whole
//     kernel), dy planes live in a 4-deep ring.  L2 -> SMEM traffic per 128 voxels: 20 KB (x) + 18 KB (dy).
// MMAs per 128 voxels and 32-channel block: 3 x 16 at N = 96 (64 clk each) instead of 9 x 16 (39 clk each).
// Split-K partials (one slab per worker CTA) go to workspace[worker][n][tap][c] and are reduced in a fixed order by
// wgrad_reduce_kernel: bit-reproducible.
#include <cuda.h>
#include <stdlib.h>
#include "conv_geom.cuh"

namespace dpi {
namespace wgm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 27)) __trap();
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// MN-major operand WITHOUT swizzle: core matrices of 8 K rows x 16 bytes (4 MN elements), K rows 16 bytes apart, the
// next 4 MN elements `mn_stride` bytes further (written to both stride fields: K = 8 is a single K group)
__device__ __forceinline__ uint64_t make_mn_desc_noswz(uint32_t saddr, uint32_t mn_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((mn_stride >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((mn_stride >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// MN-major, 128-byte swizzle with 32-byte atoms: LBO = stride between 32-element MN blocks, SBO = stride between
// 4-row K groups (512 B when rows are contiguous)
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

constexpr int TH = 16, TW = 8;                 // voxel tile (h, w): 16 K-steps of 8 voxels along w
constexpr int WW = TW + 2;                     // x rows per h-row (w halo)
constexpr int HY = TH + 2;                     // dy h-rows per plane (h halo)
constexpr int kXBytes = TH * WW * 128;         // 20480: one x plane tile, 32 channels
constexpr int kYBytes = HY * TW * 128;         // 18432: one dy plane tile, 32 channels
constexpr int kYRing = 4;
constexpr int kMaxStages = 6;
constexpr int kBN = 96;                        // 3 kh blocks x 32 output channels
constexpr int kThreads = 192;

struct Params {
  int Do, Ho, Wo;
  int tiles_w, tiles_h;
  int C, N, nkd, pd, taps;
  int c_tiles, n_tiles, workers;
  int seg_len, n_segs, n_units;
  int stages;
  uint32_t idesc, tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_march_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_dy,
                           float* __restrict__ partial, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t xbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // 2 KB of slack after the x ring: the kw = 3 pseudo-block of the last K step reads one row past a stage
  const uint32_t ybase = xbase + (uint32_t)p.stages * kXBytes + 2048u;
  const uint32_t bar_base = ybase + kYRing * kYBytes;
  auto x_full = [&](int s) { return bar_base + 8u * s; };
  auto x_empty = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto y_full = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto y_empty = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kYRing + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kMaxStages + 2 * kYRing);
  const uint32_t tmem_slot = done_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int worker = b % p.workers; b /= p.workers;
  const int nt = b % p.n_tiles;
  const int ct = b / p.n_tiles;
  const int c_base = ct * 32, n_base = nt * 32;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 1); }
    for (int s = 0; s < kYRing; ++s) { mbar_init(y_full(s), 1); mbar_init(y_empty(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_d;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

  const int nx_extra = p.nkd - 1;               // x planes per unit = L + nkd - 1 (d halo)

  if (warp == 0) {
    // ================= TMA producer (whole-warp control flow, one elected lane issues) =================
    int s = 0;
    uint32_t ph = 1;
    uint32_t yc = 0;                              // dy planes loaded so far
    for (int u = worker; u < p.n_units; u += p.workers) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int w0 = tw * TW, h0 = th * TH, d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int pz = 0; pz < L + nx_extra; ++pz) {
        if (pz < L) {
          // dy plane d_lo + pz: first needed by x plane pz (kd = 0)
          const int ys = (int)(yc & (kYRing - 1));
          mbar_wait(y_empty(ys), ((yc >> 2) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(y_full(ys), kYBytes);
            tma_load_4d(ybase + (uint32_t)ys * kYBytes, &tma_dy, y_full(ys), n_base, w0, h0 - 1, d_lo + pz);
          }
          __syncwarp();
          ++yc;
        }
        mbar_wait(x_empty(s), ph);
        if (elect_one()) {
          mbar_expect_tx(x_full(s), kXBytes);
          tma_load_4d(xbase + (uint32_t)s * kXBytes, &tma_x, x_full(s), c_base, w0 - 1, h0, d_lo - p.pd + pz);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole-warp control flow, one elected lane issues) =================
    int s = 0;
    uint32_t ph = 0;
    uint32_t yc_base = 0;
    uint32_t started = 0;                          // bit kd: accumulator kd has been written
    const uint64_t xdesc0 = make_mn_desc(xbase, 128, 512);      // M block kw = one row (voxel) further along w
    const uint64_t ydesc0 = make_mn_desc(ybase, TW * 128, 512);  // N block j = one h-row (8 rows) further down
    for (int u = worker; u < p.n_units; u += p.workers) {
      const int seg = u / (p.tiles_w * p.tiles_h);
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int pz = 0; pz < L + nx_extra; ++pz) {
        if (pz < L) {
          const uint32_t yc = yc_base + (uint32_t)pz;
          mbar_wait(y_full((int)(yc & (kYRing - 1))), (yc >> 2) & 1u);
        }
        mbar_wait(x_full(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad0 = xdesc0 + (uint64_t)((uint32_t)s * (uint32_t)(kXBytes >> 4));
          for (int kd = 0; kd < p.nkd; ++kd) {
            const int do_rel = pz - kd;             // x plane d pairs with dy plane d - kd + pd
            if (do_rel < 0 || do_rel >= L) continue;
            const uint32_t ys = (yc_base + (uint32_t)do_rel) & (kYRing - 1);
            const uint64_t bd0 = ydesc0 + (uint64_t)(ys * (uint32_t)(kYBytes >> 4));
            const uint32_t dcol = tmem_d + (uint32_t)(kd * kBN);
            const uint32_t acc0 = (started >> kd) & 1u;
#pragma unroll
            for (int g = 0; g < TH; ++g)            // K step g = h-row g: x rows 10g.., dy rows 8g..
              umma_tf32(dcol, ad0 + (uint64_t)(g * WW * 8), bd0 + (uint64_t)(g * TW * 8), p.idesc, g == 0 ? acc0 : 1u);
          }
          umma_commit(x_empty(s));
          // dy plane pz - (nkd-1) had its last use (kd = nkd-1) here
          const int yd = pz - nx_extra;
          if (yd >= 0) umma_commit(y_empty((int)((yc_base + (uint32_t)yd) & (kYRing - 1))));
        }
        __syncwarp();
        for (int kd = 0; kd < p.nkd; ++kd) {
          const int do_rel = pz - kd;
          if (do_rel >= 0 && do_rel < L) started |= 1u << kd;
        }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      yc_base += (uint32_t)L;
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ================= epilogue: lanes [32kw, 32kw+32) x columns [32j, 32j+32) of accumulator kd ==============
    //                   -> dW[n][(kd, kh = 2 - j, kw)][c]
    const int kw = warp & 3;                  // TMEM lane quarter == kw tap (quarter 3 is the unused pseudo-tap)
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = partial + (int64_t)worker * p.N * p.taps * p.C;
    const int c = c_base + lane;
    for (int kd = 0; kd < p.nkd; ++kd) {
      for (int j = 0; j < 3; ++j) {
        const int tap = (kd * 3 + (2 - j)) * 3 + kw;
        for (int nn = 0; nn < 32; nn += 16) {
          float v[16];
          tmem_ld16(tmem_d + ((uint32_t)(kw * 32) << 16) + (uint32_t)(kd * kBN + j * 32 + nn), v);
          if (kw < 3 && c < p.C) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = n_base + nn + i;
              if (n < p.N) dst[((int64_t)n * p.taps + tap) * p.C + c] = v[i];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
  }
}

// =====================================================================================================================
// Thin outputs (Cout <= 16, 3-D): the kd taps packed into the N block THROUGH THE TMA BOX.
//
// The kernel above spends three MMAs (kd = 0, 1, 2; M 128 x N 96 x K 8, 64 clk each) per 8 voxels whatever the channel
// counts, and with Cout = 4 / 8 / 16 three quarters or more of every 32-channel N block are zeros (4 -> 8, 8 -> 13,
// 25 -> 16, 25 -> 1, 64 -> 4, 67 -> 4: the full-resolution layers, ~450 us each, 1280 us for 67 -> 4).  Narrower
// MN-major blocks would need the 64- / 32-byte swizzled layouts, which kind::tf32 does not read correctly (tried; the
// 128-byte / 32-byte-atom layout is the one that works).  So the BLOCK STAYS 32 wide and is filled with useful data
// instead with pb = 32/nb consecutive dy PLANES: dy is presented to the TMA as (channel, d, w, h) and one UNSWIZZLED box
// of [nb channels][pb planes][8 w][18 h] lands densely as 128 bytes per voxel = pb planes side by side.  (A swizzled box
// cannot do that - the TMA gives every innermost box row its own 128-byte shared-memory row - and the unswizzled
// MN-major descriptor form returned zeros for kind::tf32; both probed, profiles/r2_probe_tma_box_layout.txt.)  The four
// epilogue warps, idle until the end of the kernel anyway, therefore put each tile into the operand layout the tensor
// core reads: within 128-byte row r, 32-byte chunk j moves to chunk j ^ (r & 3) (the 128-byte swizzle with 32-byte atoms,
// read off the same probe) - one row per thread, in place.  The MMA (same descriptors, same layout as above) then
// multiplies x plane d with dy planes d-1 .. d+pb-2 at once:
//   nb = 8  (Cout <= 8):  pb = 4, ONE MMA per 8 voxels covers all 27 taps (3x fewer MMAs);
//   nb = 16 (Cout <= 16): pb = 2, two windows (d-1, d) and (d+1, d+2): two MMAs instead of three.
// Accumulator a (one per window), column 32 j + nb q + n  <->  tap kh = 2 - j, dy plane d - 1 + a pb + q (kd = 2 - (a pb +
// q); planes past kd = 0 are junk columns that are never stored), output channel n (= 4 group + channel in group).
// Work units partition the X planes (every x plane meets all its dy planes, out-of-range planes are TMA zero fill), so
// there are no halo planes and nothing is counted twice.  A stage = the x tile + its dy window boxes.
struct ParamsKd {
  int Do, Ho, Wo;
  int tiles_w, tiles_h;
  int C, N, taps;
  int c_tiles, workers;
  int seg_len, n_segs, n_units;
  int stages, nb, pb, nwin;
  uint32_t idesc, tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 2)
conv_tc_wgrad_kdpack_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_dy,
                            float* __restrict__ partial, const ParamsKd p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)(kXBytes + p.nwin * kYBytes);
  // (the kw = 3 pseudo-block of the last K step reads one row past the x tile: into the dy tile of the same stage)
  const uint32_t bar_base = sbase + (uint32_t)p.stages * stage_bytes + 2048u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // TMA bytes of a stage have landed
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };       // the MMAs of a stage have completed
  auto ready_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };   // dy tiles of a stage are in operand layout
  const uint32_t done_bar = bar_base + 8u * (3 * kMaxStages);
  const uint32_t tmem_slot = done_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = blockIdx.x % p.workers;
  const int ct = blockIdx.x / p.workers;
  const int c_base = ct * 32;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(ready_bar(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_d;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer =================
    int s = 0;
    uint32_t ph = 1;
    const uint32_t tx = stage_bytes;
    for (int u = worker; u < p.n_units; u += p.workers) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int w0 = tw * TW, h0 = th * TH, d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int pz = 0; pz < L; ++pz) {
        const int d = d_lo + pz;
        mbar_wait(empty_bar(s), ph);
        if (elect_one()) {
          const uint32_t dst = sbase + (uint32_t)s * stage_bytes;
          mbar_expect_tx(full_bar(s), tx);
          tma_load_4d(dst, &tma_x, full_bar(s), c_base, w0 - 1, h0, d);
          for (int a = 0; a < p.nwin; ++a)      // dy planes d - 1 + a pb .. : (channel, d, w, h) box, unswizzled
            tma_load_4d(dst + (uint32_t)kXBytes + (uint32_t)a * kYBytes, &tma_dy, full_bar(s), 0, d - 1 + a * p.pb, w0, h0 - 1);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole-warp control flow, one elected lane issues) =================
    int s = 0;
    uint32_t ph = 0;
    uint32_t started = 0;
    const uint64_t xdesc0 = make_mn_desc(sbase, 128, 512);                    // M block kw = one voxel further along w
    const uint64_t ydesc0 = make_mn_desc(sbase + (uint32_t)kXBytes, TW * 128, 512);   // N block j = one h-row further down
    const uint64_t stage_u = (uint64_t)(stage_bytes >> 4);
    uint64_t xd = xdesc0, yd = ydesc0;
    for (int u = worker; u < p.n_units; u += p.workers) {
      const int seg = u / (p.tiles_w * p.tiles_h);
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int pz = 0; pz < L; ++pz) {
        mbar_wait(ready_bar(s), ph);          // (implies full_bar(s): the x tile has landed too)
        tc_fence_after();
        if (elect_one()) {
          if (p.nwin == 1) {
#pragma unroll
            for (int g = 0; g < TH; ++g)
              umma_tf32(tmem_d, xd + (uint64_t)(g * WW * 8), yd + (uint64_t)(g * TW * 8), p.idesc, g == 0 ? started : 1u);
          } else {
#pragma unroll
            for (int g = 0; g < TH; ++g)
              umma_tf32(tmem_d, xd + (uint64_t)(g * WW * 8), yd + (uint64_t)(g * TW * 8), p.idesc, g == 0 ? started : 1u);
#pragma unroll
            for (int g = 0; g < TH; ++g)
              umma_tf32(tmem_d + 96u, xd + (uint64_t)(g * WW * 8), yd + (uint64_t)(kYBytes >> 4) + (uint64_t)(g * TW * 8), p.idesc,
                        g == 0 ? started : 1u);
          }
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        started = 1u;
        xd += stage_u; yd += stage_u;
        if (++s == p.stages) { s = 0; ph ^= 1u; xd = xdesc0; yd = ydesc0; }
      }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ================= dy tiles -> operand layout (all through the main loop), then the epilogue =================
    {
      const int t = (int)threadIdx.x - 64;                  // 0..127
      int s = 0;
      uint32_t ph = 0;
      for (int u = worker; u < p.n_units; u += p.workers) {
        const int seg = u / (p.tiles_w * p.tiles_h);
        const int L = min(p.seg_len, p.Do - seg * p.seg_len);
        for (int pz = 0; pz < L; ++pz) {
          mbar_wait(full_bar(s), ph);
          uint8_t* tile = smem_raw + (sbase - smem_u32(smem_raw)) + (size_t)s * stage_bytes + kXBytes;
          for (int a = 0; a < p.nwin; ++a, tile += kYBytes) {
            // 16-byte item i = 8 row + chunk: consecutive threads take consecutive items (conflict-free: a one-row-per-
            // thread version put all 32 lanes on the same four banks and made this loop, not the MMAs, the bound);
            // all reads of the tile happen before any write
            constexpr int kItems = HY * TW * 8;            // 1152 = 9 x 128
            float4* t4 = reinterpret_cast<float4*>(tile);
            float4 v[kItems / 128];
#pragma unroll
            for (int k = 0; k < kItems / 128; ++k) v[k] = t4[t + k * 128];
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int k = 0; k < kItems / 128; ++k) {
              const int it = t + k * 128, r = it >> 3, i = it & 7;
              t4[(r << 3) | ((((i >> 1) ^ (r & 3)) << 1) | (i & 1))] = v[k];
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (t == 0) mbar_arrive(ready_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    const int kw = warp & 3;                  // TMEM lane quarter == kw tap (quarter 3 is the unused pseudo-tap)
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = partial + (int64_t)worker * p.N * p.taps * p.C;
    const int c = c_base + lane;
    for (int a = 0; a < p.nwin; ++a) {
      for (int j = 0; j < 3; ++j) {
        for (int nn = 0; nn < 32; nn += 16) {
          float v[16];
          tmem_ld16(tmem_d + ((uint32_t)(kw * 32) << 16) + (uint32_t)(a * 96 + j * 32 + nn), v);
          if (kw < 3 && c < p.C) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = nn + i;
              const int q = col / p.nb, n = col - q * p.nb;
              const int kd = 2 - (a * p.pb + q);
              if (kd >= 0 && n < p.N) dst[((int64_t)n * p.taps + ((kd * 3 + (2 - j)) * 3 + kw)) * p.C + c] = v[i];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// persistent CTAs: the SM count, or DPI_TC_WGRAD_MARCH_CTAS (test knob: several units per worker on small problems)
static int cta_budget() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  const char* e = getenv("DPI_TC_WGRAD_MARCH_CTAS");
  if (e && e[0]) {
    const int v = atoi(e);
    if (v > 0) return v;
  }
  return n;
}

static bool plan(const GatherGeom& g, Params& p) {
  if (g.transposed || g.sd != 1 || g.sh != 1 || g.sw != 1 || g.kw != 3 || g.kh != 3 || (g.kd != 3 && g.kd != 1)) return false;
  if ((g.C & 3) || (g.N & 3)) return false;
  const char* e = getenv("DPI_TC_WGRAD_MARCH");
  if (e && e[0] == '0') return false;
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N; p.nkd = g.kd; p.pd = g.pd;
  p.taps = g.kd * 9;
  p.tiles_w = (g.Wo + TW - 1) / TW;
  p.tiles_h = (g.Ho + TH - 1) / TH;
  p.c_tiles = (g.C + 31) / 32;
  p.n_tiles = (g.N + 31) / 32;
  // every (c block, n block) pair re-streams both operands: past a few pairs conv_tc_wgrad_kw.cu moves less data
  const int pairs = p.c_tiles * p.n_tiles;
  if (pairs > 8) return false;
  const int budget = cta_budget();
  if (pairs > budget) return false;
  int workers = budget / pairs;
  const int ncol = p.tiles_w * p.tiles_h;
  // segments of dy planes: minimise (rounds over the workers) x (x planes a unit streams)
  double best = 1e30;
  p.seg_len = g.Do; p.n_segs = 1;
  for (int want = 1; want <= g.Do; ++want) {
    const int len = (g.Do + want - 1) / want;
    const int segs = (g.Do + len - 1) / len;
    const int64_t units = (int64_t)ncol * segs;
    const int64_t rounds = (units + workers - 1) / workers;
    const double cost = (double)rounds * (len + p.nkd - 1 + 0.5);
    if (cost < best - 1e-9) { best = cost; p.seg_len = len; p.n_segs = segs; }
  }
  p.n_units = ncol * p.n_segs;
  if (workers > p.n_units) workers = p.n_units;
  p.workers = workers;
  p.stages = (tc_smem_budget(true) - 1024 - 2048 - 256 - kYRing * kYBytes) / kXBytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 2) return false;
  p.tmem_cols = p.nkd == 1 ? 128u : 512u;        // nkd x 96 columns
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kBN >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  return true;
}

static bool plan_kd(const GatherGeom& g, ParamsKd& p) {
  if (g.transposed || g.sd != 1 || g.sh != 1 || g.sw != 1 || g.kw != 3 || g.kh != 3 || g.kd != 3) return false;
  if ((g.C & 3) || (g.N & 3) || g.N > 16) return false;
  {
    const char* e = getenv("DPI_TC_WGRAD_KDPACK");
    if (e && e[0] == '0') return false;
    const char* e2 = getenv("DPI_TC_WGRAD_MARCH");
    if (e2 && e2[0] == '0') return false;
  }
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N; p.taps = 27;
  p.tiles_w = (g.Wo + TW - 1) / TW;
  p.tiles_h = (g.Ho + TH - 1) / TH;
  p.c_tiles = (g.C + 31) / 32;
  if (p.c_tiles > 8) return false;
  p.nb = g.N <= 8 ? 8 : 16;
  p.pb = 32 / p.nb;
  p.nwin = (3 + p.pb - 1) / p.pb;
  // one window (Cout <= 8): a stage is 38 KB and the accumulator 128 TMEM columns, so two CTAs could share an SM (two
  // stages each).  Measured (DPI_TC_WGRAD_2CTA=1, profiles/r2_two_cta_timing.txt): 5 % SLOWER per launch (4 -> 8
  // 189 -> 202 us, 64 -> 4 372 -> 398) - this kernel is bound by the tensor pipe, which two CTAs only share - so off
  static const int two_ok = [] { const char* e = getenv("DPI_TC_WGRAD_2CTA"); return (e && e[0] == '1') ? 1 : 0; }();
  const int per_sm = (two_ok && p.nwin == 1) ? 2 : 1;
  const int budget = cta_budget() * per_sm;
  if (p.c_tiles > budget) return false;
  int workers = budget / p.c_tiles;
  const int ncol = p.tiles_w * p.tiles_h;
  double best = 1e30;
  p.seg_len = g.Do; p.n_segs = 1;
  for (int want = 1; want <= g.Do; ++want) {
    const int len = (g.Do + want - 1) / want;
    const int segs = (g.Do + len - 1) / len;
    const int64_t units = (int64_t)ncol * segs;
    const int64_t rounds = (units + workers - 1) / workers;
    const double cost = (double)rounds * (len + 0.5);
    if (cost < best - 1e-9) { best = cost; p.seg_len = len; p.n_segs = segs; }
  }
  p.n_units = ncol * p.n_segs;
  if (workers > p.n_units) workers = p.n_units;
  p.workers = workers;
  const int stage_bytes = kXBytes + p.nwin * kYBytes;
  p.stages = ((per_sm == 2 ? 113 * 1024 : tc_smem_budget(true)) - 1024 - 2048 - 256) / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 2) return false;
  p.tmem_cols = p.nwin == 1 ? 128u : 256u;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  return true;
}

static int launch_kd(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial, int64_t partial_bytes,
                     const GatherGeom& g, const ParamsKd& p, int* nchunks_out, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  const int64_t need = (int64_t)p.workers * g.N * p.taps * g.C * (int64_t)sizeof(float);
  if (partial_bytes < need) {
    set_error("conv_tc_wgrad_march(kd): workspace too small (%lld < %lld)", (long long)partial_bytes, (long long)need);
    return DPI_ERR_WORKSPACE;
  }
  CUtensorMap mx, mdy;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)g.Wi * x_ld * 4, (cuuint64_t)g.Hi * g.Wi * x_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)WW, (cuuint32_t)TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_march(kd): cuTensorMapEncodeTiled(x) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  {
    // dy as (channel, d, w, h), no swizzle: the box lands densely as [h][w][plane][channel] = 128 bytes per voxel
    cuuint64_t dims[4] = {(cuuint64_t)g.N, (cuuint64_t)g.Do, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho};
    cuuint64_t strides[3] = {(cuuint64_t)g.Ho * g.Wo * dy_ld * 4, (cuuint64_t)dy_ld * 4, (cuuint64_t)g.Wo * dy_ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.nb, (cuuint32_t)p.pb, (cuuint32_t)TW, (cuuint32_t)HY};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_march(kd): cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * (kXBytes + p.nwin * kYBytes) + 2048 + 8 * (3 * kMaxStages + 4) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_kdpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("wgrad_march(kd): cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const unsigned grid = (unsigned)(p.workers * p.c_tiles);
  conv_tc_wgrad_kdpack_kernel<<<grid, kThreads, smem, st>>>(mx, mdy, partial, p);
  *nchunks_out = p.workers;
  return check_launch("conv_tc_wgrad_kdpack_kernel");
}

}  // namespace wgm

int64_t conv_tc_wgrad_march_workspace_bytes(const GatherGeom& g) {
  {
    wgm::ParamsKd pk;
    if (wgm::plan_kd(g, pk)) return (int64_t)pk.workers * g.N * pk.taps * g.C * (int64_t)sizeof(float);
  }
  wgm::Params p;
  if (!wgm::plan(g, p)) return 0;
  return (int64_t)p.workers * g.N * p.taps * g.C * (int64_t)sizeof(float);
}

int conv_tc_wgrad_march(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial,
                        int64_t partial_bytes, const GatherGeom& g, int* nchunks_out, cudaStream_t st) {
  using namespace wgm;
  {
    ParamsKd pk;
    if (plan_kd(g, pk)) return launch_kd(x, x_ld, dy, dy_ld, partial, partial_bytes, g, pk, nchunks_out, st);
  }
  Params p;
  if (!plan(g, p)) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  const int64_t need = (int64_t)p.workers * g.N * p.taps * g.C * (int64_t)sizeof(float);
  if (partial_bytes < need) {
    set_error("conv_tc_wgrad_march: workspace too small (%lld < %lld)", (long long)partial_bytes, (long long)need);
    return DPI_ERR_WORKSPACE;
  }
  CUtensorMap mx, mdy;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)g.Wi * x_ld * 4, (cuuint64_t)g.Hi * g.Wi * x_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)WW, (cuuint32_t)TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_march: cuTensorMapEncodeTiled(x) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.N, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)g.Do};
    cuuint64_t strides[3] = {(cuuint64_t)dy_ld * 4, (cuuint64_t)g.Wo * dy_ld * 4, (cuuint64_t)g.Ho * g.Wo * dy_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)HY, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_march: cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * kXBytes + 2048 + (size_t)kYRing * kYBytes + 8 * (2 * kMaxStages + 2 * kYRing + 4) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("wgrad_march: cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const unsigned grid = (unsigned)(p.workers * p.c_tiles * p.n_tiles);
  conv_tc_wgrad_march_kernel<<<grid, kThreads, smem, st>>>(mx, mdy, partial, p);
  *nchunks_out = p.workers;
  return check_launch("conv_tc_wgrad_march_kernel");
}

}  // namespace dpi
