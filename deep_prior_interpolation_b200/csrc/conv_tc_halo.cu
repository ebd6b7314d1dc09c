// tcgen05 implicit-GEMM 3x3(x3) convolution with SHARED-MEMORY HALO REUSE (forward and stride-1 dgrad).
//
// conv_tc.cu issues one TMA box load per (tap, channel chunk): every input voxel travels L2 -> SMEM 27 times.
// Here a CTA owns an output tile of 1 x 16 x 8 voxels (d, h, w) and loads, per (channel chunk, kd), ONE halo
// plane of 18 x 10 voxels; the nine (kh, kw) taps of that plane are nine *views* of the same shared-memory
// rows: the K-major SWIZZLE_128B matrix descriptor starts (kh*10 + kw) rows into the plane and strides
// 10 rows between 8-row groups (SBO = 1280 B).  The tensor core applies the 128-byte swizzle on absolute
// shared-memory address bits, so a row-shifted, non-1024-aligned descriptor reads exactly what TMA wrote
// (measured on B200: scratch/umma_probe.cu, all shifts / SBO 1024,1152,1280 exact).  L2 -> SMEM traffic drops
// from 27 x 128 rows to 3 x 180 rows per tile and channel chunk (6.4x).
//
// The nine weight tiles of the plane arrive with one rank-3 TMA load ([tap][n][c] box of the packed weights).
// Accumulator: 128 x BN fp32 in TMEM.  Channel chunk is always 32 floats (one 128-byte swizzle row); the tail
// chunk issues only ceil(rem/8) K-steps.  Warp roles as in conv_tc.cu.
#include <cuda.h>
#include <stdlib.h>
#include "conv_geom.cuh"

namespace dpi {
namespace halo {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
    if (spin > (1u << 26)) __trap();
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// descriptors as (lo, hi) halves: advancing the start address is one 32-bit add (see conv_tc_wgrad_kw.cu)
__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                             uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accum) : "memory");
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
// K-major swizzled descriptor (layout 2/4/6 = SWIZZLE_128B/64B/32B) with an arbitrary 8-row-group stride
__device__ __forceinline__ uint64_t make_k_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

constexpr int TH = 16, TW = 8;              // output tile (h, w); 128 rows
constexpr int HH = TH + 2, WW = TW + 2;     // halo plane
constexpr int kThreads = 192;

struct HaloParams {
  int Do, Ho, Wo;
  int tiles_w, tiles_h;
  int C, N, nkd, pd, transposed;
  int kc, rb, layout;           // channel chunk (8/16/32 floats), row bytes (32/64/128), UMMA layout type (6/4/2)
  int n_chunks;                 // ceil(C / kc)
  int n_iters;                  // n_chunks * nkd
  int BN, stages;
  int plane_bytes;              // 180 rows x rb, rounded up to 1024
  int b_bytes;                  // 9 * BN * rb, rounded up to 1024
  uint32_t idesc, tmem_cols;
  int64_t out_ld;
  int accumulate;
  // stride-2 forward (the down-sampling convs, mulresunet.py:224-227): output (oh, ow) reads input (2 oh + kh - 1,
  // 2 ow + kw - 1), so the plane of one kd tap is loaded as its FOUR (h, w) parity classes - class (ph, pw) = rows
  // 2 (h0 + i) - ph, columns 2 (w0 + j) - pw, a dense (16 + ph) x (8 + pw) tile through TMA element strides - and the
  // nine taps are row-shifted views of them (kh = 1 -> even rows; kh = 0 / 2 -> odd rows, shift 0 / 1; same for kw):
  // 4 boxes per plane instead of the 9 per-tap boxes of conv_tc_kernel.
  int s2;
  int sub_off[4];               // byte offset of class ph * 2 + pw inside the stage
  // stride-2 DATA GRADIENT, all eight output parity classes in one launch: dx[2u + c] = sum over the taps t with (c + 1 - t)
  // even of W[t] . dy[u + (c + 1 - t)/2], i.e. per axis tap t belongs to class c(t) = (t != 1) and reads dy at offset
  // o(t) = (t == 0).  A CTA owns a 16 x 8 tile of u positions of one u plane: stage (chunk, td) = ONE 17 x 9 dy tile of
  // plane u_d + o(td) + the nine weight tiles of that td; tap (th, tw) is the view shifted by (o(th), o(tw)) and
  // accumulates into the TMEM accumulator of class (c(td), c(th), c(tw)) - eight accumulators of BN columns.  The epilogue
  // writes the 2 x 32 x 16 dx voxels of the tile, every one exactly once (whole lines, no per-class launches).
  int s2t;
  int OD, OH, OW;               // dx extents (s2t)
};

__global__ void __launch_bounds__(kThreads)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_a1, const __grid_constant__ CUtensorMap tma_a2,
                    const __grid_constant__ CUtensorMap tma_a3, const float* __restrict__ bias, float* __restrict__ out,
                    const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)p.plane_bytes + (uint32_t)p.b_bytes;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = tmem_full_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h;
  const int d0 = t / p.tiles_h;
  const int w0 = tw * TW, h0 = th * TH;
  const int n0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
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
    if (lane == 0) {
      // ================= TMA producer: one halo plane + nine weight tiles per (chunk, kd) =================
      for (int it = 0; it < p.n_iters; ++it) {
        const int s = it % p.stages;
        const int chunk = it / p.nkd, kd = it - chunk * p.nkd;
        mbar_wait(empty_bar(s), ((uint32_t)(it / p.stages) & 1u) ^ 1u);
        const uint32_t a_dst = base + s * stage_bytes;
        if (p.s2t) {
          mbar_expect_tx(full_bar(s), (uint32_t)((TH + 1) * (TW + 1) * p.rb) + (uint32_t)(9 * p.BN * p.rb));
          const int dz = p.nkd == 3 ? d0 + (kd == 0 ? 1 : 0) : d0;
          tma_load_4d(a_dst, &tma_a, full_bar(s), chunk * p.kc, w0, h0, dz);
          tma_load_3d(a_dst + p.plane_bytes, &tma_b, full_bar(s), chunk * p.kc, n0, kd * 9);
          continue;
        }
        if (p.s2) {
          mbar_expect_tx(full_bar(s), (uint32_t)((TH * TW + TH * (TW + 1) + (TH + 1) * TW + (TH + 1) * (TW + 1)) * p.rb) +
                                          (uint32_t)(9 * p.BN * p.rb));
          const int dz = p.nkd == 3 ? 2 * d0 + kd - 1 : d0;
          tma_load_4d(a_dst + (uint32_t)p.sub_off[0], &tma_a, full_bar(s), chunk * p.kc, 2 * w0, 2 * h0, dz);
          tma_load_4d(a_dst + (uint32_t)p.sub_off[1], &tma_a1, full_bar(s), chunk * p.kc, 2 * w0 - 1, 2 * h0, dz);
          tma_load_4d(a_dst + (uint32_t)p.sub_off[2], &tma_a2, full_bar(s), chunk * p.kc, 2 * w0, 2 * h0 - 1, dz);
          tma_load_4d(a_dst + (uint32_t)p.sub_off[3], &tma_a3, full_bar(s), chunk * p.kc, 2 * w0 - 1, 2 * h0 - 1, dz);
          tma_load_3d(a_dst + p.plane_bytes, &tma_b, full_bar(s), chunk * p.kc, n0, kd * 9);
          continue;
        }
        mbar_expect_tx(full_bar(s), (uint32_t)(HH * WW * p.rb) + (uint32_t)(9 * p.BN * p.rb));
        // plane index along d: forward reads d0 + kd - pd ; dgrad reads d0 + pd - kd
        const int dz = p.transposed ? d0 + p.pd - kd : d0 + kd - p.pd;
        tma_load_4d(a_dst, &tma_a, full_bar(s), chunk * p.kc, w0 - 1, h0 - 1, dz);
        tma_load_3d(a_dst + p.plane_bytes, &tma_b, full_bar(s), chunk * p.kc, n0, kd * 9);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: whole-warp control flow, one elected lane issues (see conv_tc_march.cu) =======
    uint32_t started = 0;                          // s2t: bit a = accumulator a has been written
    for (int it = 0; it < p.n_iters; ++it) {
      const int s = it % p.stages;
      const int chunk = it / p.nkd;
      mbar_wait(full_bar(s), (uint32_t)(it / p.stages) & 1u);
      tc_fence_after();
      const int kd_it = it - chunk * p.nkd;
      const int cd = (p.nkd == 3 && kd_it != 1) ? 1 : 0;     // s2t: d class of this stage's taps
      if (elect_one()) {
        const uint32_t a0 = base + s * stage_bytes, b0 = a0 + p.plane_bytes;
        const int rem = p.C - chunk * p.kc;
        const int ksteps = rem >= p.kc ? (p.kc >> 3) : (rem + 7) >> 3;
        const uint32_t ru = (uint32_t)(p.rb >> 4);                   // row pitch in 16-byte units
        const uint64_t ad0 = make_k_desc(a0, WW * p.rb, p.layout), bd0 = make_k_desc(b0, 8 * p.rb, p.layout);
        const uint32_t bstep = (uint32_t)p.BN * ru;                   // BN rows, in 16-byte units
        const uint32_t alo0 = (uint32_t)ad0, ahi = (uint32_t)(ad0 >> 32), blo0 = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32);
        uint32_t al[9], bl[9];
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
          const int kh = tp / 3, kw = tp - 3 * kh;
          // forward: output (h, w) reads halo row (h + kh, w + kw); dgrad reads (h + 2 - kh, w + 2 - kw)
          al[tp] = alo0 + (uint32_t)(p.transposed ? ((2 - kh) * WW + (2 - kw)) : (kh * WW + kw)) * ru;
          bl[tp] = blo0 + bstep * (uint32_t)tp;
        }
        uint32_t accum = it > 0 ? 1u : 0u;
        if (p.s2t) {
          const uint64_t ad9 = make_k_desc(a0, (TW + 1) * p.rb, p.layout);
          const uint32_t ahi9 = (uint32_t)(ad9 >> 32);
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const int th = tp / 3, tw = tp - 3 * th;
            const int a = (cd * 2 + (th != 1 ? 1 : 0)) * 2 + (tw != 1 ? 1 : 0);
            const uint32_t astart = alo0 + (uint32_t)((th == 0 ? TW + 1 : 0) + (tw == 0 ? 1 : 0)) * ru;
            const uint32_t dcol = tmem_d + (uint32_t)(a * p.BN);
            umma_tf32_lh(dcol, astart, ahi9, bl[tp], bhi, p.idesc, (started >> a) & 1u);
            if (ksteps > 1) umma_tf32_lh(dcol, astart + 2, ahi9, bl[tp] + 2, bhi, p.idesc, 1u);
            if (ksteps > 2) umma_tf32_lh(dcol, astart + 4, ahi9, bl[tp] + 4, bhi, p.idesc, 1u);
            if (ksteps > 3) umma_tf32_lh(dcol, astart + 6, ahi9, bl[tp] + 6, bhi, p.idesc, 1u);
            started |= 1u << a;
          }
        } else if (p.s2) {
          // tap (kh, kw) -> parity class (kh != 1, kw != 1), row / column shift (kh == 2, kw == 2); the 8-row groups of a
          // class are its (8 + pw)-voxel rows
          const uint64_t ad8 = make_k_desc(a0, TW * p.rb, p.layout), ad9 = make_k_desc(a0, (TW + 1) * p.rb, p.layout);
          const uint32_t ahi8 = (uint32_t)(ad8 >> 32), ahi9 = (uint32_t)(ad9 >> 32);
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const int kh = tp / 3, kw = tp - 3 * kh;
            const int ph = kh != 1, pw = kw != 1;
            const uint32_t astart = alo0 + ((uint32_t)p.sub_off[ph * 2 + pw] >> 4) +
                                    (uint32_t)((kh == 2 ? TW + pw : 0) + (kw == 2 ? 1 : 0)) * ru;
            const uint32_t ah = pw ? ahi9 : ahi8;
            umma_tf32_lh(tmem_d, astart, ah, bl[tp], bhi, p.idesc, accum);
            if (ksteps > 1) umma_tf32_lh(tmem_d, astart + 2, ah, bl[tp] + 2, bhi, p.idesc, 1u);
            if (ksteps > 2) umma_tf32_lh(tmem_d, astart + 4, ah, bl[tp] + 4, bhi, p.idesc, 1u);
            if (ksteps > 3) umma_tf32_lh(tmem_d, astart + 6, ah, bl[tp] + 6, bhi, p.idesc, 1u);
            accum = 1;
          }
        } else if (ksteps == 4) {
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            umma_tf32_lh(tmem_d, al[tp], ahi, bl[tp], bhi, p.idesc, accum);
            umma_tf32_lh(tmem_d, al[tp] + 2, ahi, bl[tp] + 2, bhi, p.idesc, 1u);
            umma_tf32_lh(tmem_d, al[tp] + 4, ahi, bl[tp] + 4, bhi, p.idesc, 1u);
            umma_tf32_lh(tmem_d, al[tp] + 6, ahi, bl[tp] + 6, bhi, p.idesc, 1u);
            accum = 1;
          }
        } else if (ksteps == 1) {
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            umma_tf32_lh(tmem_d, al[tp], ahi, bl[tp], bhi, p.idesc, accum);
            accum = 1;
          }
        } else if (ksteps == 2) {
          // (straight-line also for the tail chunks: a rolled K loop costs the issuing lane more than the MMAs)
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            umma_tf32_lh(tmem_d, al[tp], ahi, bl[tp], bhi, p.idesc, accum);
            umma_tf32_lh(tmem_d, al[tp] + 2, ahi, bl[tp] + 2, bhi, p.idesc, 1u);
            accum = 1;
          }
        } else {
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            umma_tf32_lh(tmem_d, al[tp], ahi, bl[tp], bhi, p.idesc, accum);
            umma_tf32_lh(tmem_d, al[tp] + 2, ahi, bl[tp] + 2, bhi, p.idesc, 1u);
            umma_tf32_lh(tmem_d, al[tp] + 4, ahi, bl[tp] + 4, bhi, p.idesc, 1u);
            accum = 1;
          }
        }
        umma_commit(empty_bar(s));
        if (it == p.n_iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      // (the accumulators this stage touched, for every lane: whichever lane is elected next must know)
      if (p.s2t) started |= cd ? 0xF0u : 0x0Fu;
    }
  } else if (p.s2t) {
    // ================= epilogue, stride-2 data gradient: eight accumulators -> the 2 x 2 x 2 dx voxels of every u ======
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int uw = w0 + (row & 7), uh = h0 + (row >> 3);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int ncls = p.nkd == 3 ? 8 : 4;
    for (int a = 0; a < ncls; ++a) {
      const int cd = p.nkd == 3 ? (a >> 2) : 0, ch = (a >> 1) & 1, cw = a & 1;
      const int od = p.nkd == 3 ? 2 * d0 + cd : d0, oh = 2 * uh + ch, ow = 2 * uw + cw;
      const bool valid = od < p.OD && oh < p.OH && ow < p.OW;
      float* orow = out + (((int64_t)od * p.OH + oh) * p.OW + ow) * p.out_ld + n0;
      for (int c = 0; c < p.BN; c += 16) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.BN + c), v);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int n = n0 + c + i;
            if (n < p.N) {
              float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              float4* dst = reinterpret_cast<float4*>(orow + c + i);
              if (p.accumulate) {
                const float4 o4 = *dst;
                r.x += o4.x; r.y += o4.y; r.z += o4.z; r.w += o4.w;
              }
              *dst = r;
            }
          }
        }
      }
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ow = w0 + (row & 7), oh = h0 + (row >> 3);
    const bool valid = ow < p.Wo && oh < p.Ho;
    float* orow = out + (((int64_t)d0 * p.Ho + oh) * p.Wo + ow) * p.out_ld + n0;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    for (int c = 0; c < p.BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const int n = n0 + c + i;
          if (n < p.N) {
            float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (bias) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
              r.x += b4.x; r.y += b4.y; r.z += b4.z; r.w += b4.w;
            }
            float4* dst = reinterpret_cast<float4*>(orow + c + i);
            if (p.accumulate) {
              const float4 o4 = *dst;
              r.x += o4.x; r.y += o4.y; r.z += o4.z; r.w += o4.w;
            }
            *dst = r;
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

}  // namespace halo

int conv_tc_halo_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                        const GatherGeom& g, int accumulate, cudaStream_t st) {
  using namespace halo;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("DPI_TC_HALO");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return DPI_ERR_UNSUPPORTED;
  if (g.kh != 3 || g.kw != 3 || (g.kd != 3 && g.kd != 1)) return DPI_ERR_UNSUPPORTED;
  const bool s1 = g.sd == 1 && g.sh == 1 && g.sw == 1;
  // stride-2 forward: (2, 2, 2) for 3x3x3, (1, 2, 2) for 1x3x3 (DPI_TC_HALO_S2=0 leaves it to conv_tc_kernel)
  const int s2_on = [] { const char* e = getenv("DPI_TC_HALO_S2"); return (e && e[0] == '0') ? 0 : 1; }();
  const bool s2 = s2_on && !g.transposed && g.sh == 2 && g.sw == 2 && g.ph == 1 && g.pw == 1 &&
                  ((g.kd == 3 && g.sd == 2 && g.pd == 1) || (g.kd == 1 && g.sd == 1));
  // stride-2 data gradient, all parity classes in one launch (DPI_TC_HALO_S2T=0 leaves it to the per-class march launches)
  // (read on every call - a getenv is nothing next to a launch - so that the tests can exercise both paths in one process)
  const int s2t_on = [] { const char* e = getenv("DPI_TC_HALO_S2T"); return (e && e[0] == '0') ? 0 : 1; }();
  // Measured (profiles/r2_s2_halo_timing.txt): 25 -> 25 at 256x128x128 762 us (499 with one stage) against 574 us for the
  // eight per-class march launches, 51 -> 51 at 128x64x64 213 against 196, 105 -> 105 at 64x32x32 70 against 111: the
  // tile-per-CTA form with its eight-accumulator epilogue wins where the per-class launches are launch-bound, so it takes
  // the problems with at most 32 K u positions (DPI_TC_HALO_S2T_MAX_U)
  static const int64_t s2t_max_u = [] { const char* e = getenv("DPI_TC_HALO_S2T_MAX_U"); return e ? atoll(e) : 32768LL; }();
  const bool s2t = s2t_on && g.transposed && g.sh == 2 && g.sw == 2 && g.ph == 1 && g.pw == 1 &&
                   ((g.kd == 3 && g.sd == 2 && g.pd == 1) || (g.kd == 1 && g.sd == 1)) &&
                   (int64_t)g.Di * g.Hi * g.Wi <= s2t_max_u;
  if (!s1 && !s2 && !s2t) return DPI_ERR_UNSUPPORTED;
  if ((g.C & 3) || (g.N & 3)) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  HaloParams p;
  p.s2 = s2 ? 1 : 0;
  p.s2t = s2t ? 1 : 0;
  p.OD = g.Do; p.OH = g.Ho; p.OW = g.Wo;
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N; p.nkd = g.kd; p.pd = g.pd; p.transposed = g.transposed;
  p.tiles_w = (g.Wo + TW - 1) / TW;
  p.tiles_h = (g.Ho + TH - 1) / TH;
  int grid_d = g.Do;
  if (s2t) {
    // tiles live in u space = the dy grid
    p.Do = g.Di; p.Ho = g.Hi; p.Wo = g.Wi;
    p.tiles_w = (g.Wi + TW - 1) / TW;
    p.tiles_h = (g.Hi + TH - 1) / TH;
    grid_d = g.Di;
  }
  // channel chunk = shared-memory row: 32 B / 64 B / 128 B rows with the matching swizzle (all three honour
  // row-shifted descriptors: scratch/umma_probe2.cu); thin inputs (C <= 8, <= 16) no longer pay for 128-byte rows
  p.kc = g.C <= 8 ? 8 : (g.C <= 16 ? 16 : 32);
  p.rb = p.kc * 4;
  p.layout = p.kc == 32 ? 2 : (p.kc == 16 ? 4 : 6);
  const CUtensorMapSwizzle swz = p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (p.kc == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  p.n_chunks = (g.C + p.kc - 1) / p.kc;
  p.n_iters = p.n_chunks * p.nkd;
  // keeps a stage under ~100 KB (stride 2: four class tiles of 72 KB next to the weights -> 32 columns per CTA)
  int max_bn = s2 ? (p.kc == 32 ? 32 : (p.kc == 16 ? 64 : 128)) : (p.kc == 32 ? 64 : (p.kc == 16 ? 128 : 256));
  if (s2t && max_bn > 64) max_bn = 64;                                  // eight accumulators of BN columns in TMEM
  const int n_tiles = (g.N + max_bn - 1) / max_bn;
  p.BN = (((g.N + n_tiles - 1) / n_tiles) + 15) / 16 * 16;
  p.plane_bytes = (HH * WW * p.rb + 1023) / 1024 * 1024;
  for (int i = 0; i < 4; ++i) p.sub_off[i] = 0;
  if (s2t) p.plane_bytes = ((TH + 1) * (TW + 1) * p.rb + 1023) / 1024 * 1024;
  if (s2) {
    int off = 0;
    for (int cls = 0; cls < 4; ++cls) {
      p.sub_off[cls] = off;
      off += ((TH + (cls >> 1)) * (TW + (cls & 1)) * p.rb + 1023) / 1024 * 1024;
    }
    p.plane_bytes = off;
  }
  p.b_bytes = (9 * p.BN * p.rb + 1023) / 1024 * 1024;
  const int stage_bytes = p.plane_bytes + p.b_bytes;
  int stages = (106 * 1024) / stage_bytes;                              // aim at two CTAs per SM
  if (stages < 2) stages = (225 * 1024) / stage_bytes;
  if (stages > 4) stages = 4;
  if (s2t) {
    static const int cap = [] { const char* e = getenv("DPI_TC_HALO_S2T_STAGES"); return e ? atoi(e) : 0; }();
    if (cap > 0 && stages > cap) stages = cap;
  }
  if (stages > p.n_iters) stages = p.n_iters;
  if (stages < 1 || (s2 && stages < 2 && p.n_iters > 1)) return DPI_ERR_UNSUPPORTED;
  p.stages = stages;
  int cols = 32;
  while (cols < p.BN * (s2t ? (g.kd == 3 ? 8 : 4) : 1)) cols <<= 1;
  if (cols > 512) return DPI_ERR_UNSUPPORTED;
  p.tmem_cols = (uint32_t)cols;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.out_ld = out_ld;
  p.accumulate = accumulate;

  CUtensorMap ma, mb, mas[4];
  if (s2) {
    // one map per parity class: the box spans 2 x (rows, columns) input positions, the element stride keeps every second
    for (int cls = 0; cls < 4; ++cls) {
      cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
      cuuint64_t strides[3] = {(cuuint64_t)in_ld * 4, (cuuint64_t)g.Wi * in_ld * 4, (cuuint64_t)g.Hi * g.Wi * in_ld * 4};
      cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)(2 * (TW + (cls & 1))), (cuuint32_t)(2 * (TH + (cls >> 1))), 1};
      cuuint32_t es[4] = {1, 2, 2, 1};
      CUresult r = encode(&mas[cls], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("halo(s2): cuTensorMapEncodeTiled(A%d) failed: %d", cls, (int)r); return DPI_ERR_CUDA; }
    }
    ma = mas[0];
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)in_ld * 4, (cuuint64_t)g.Wi * in_ld * 4, (cuuint64_t)g.Hi * g.Wi * in_ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)(s2t ? TW + 1 : WW), (cuuint32_t)(s2t ? TH + 1 : HH), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("halo: cuTensorMapEncodeTiled(A) failed: %d", (int)r); return DPI_ERR_CUDA; }
    mas[1] = mas[2] = mas[3] = ma;
  }
  {
    // packed weights Wp[n][tap][c] viewed as (c, n, tap): one box = [9 taps][BN][32 c]
    const int taps = g.kd * 9;
    cuuint64_t dims[3] = {(cuuint64_t)g.C, (cuuint64_t)g.N, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)taps * g.C * 4, (cuuint64_t)g.C * 4};
    cuuint32_t box[3] = {(cuuint32_t)p.kc, (cuuint32_t)p.BN, 9};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(Wp), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("halo: cuTensorMapEncodeTiled(B) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 8 * (2 * p.stages + 4) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("halo: cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  dim3 grid((unsigned)(p.tiles_w * p.tiles_h * grid_d), (unsigned)n_tiles);
  conv_tc_halo_kernel<<<grid, kThreads, smem, st>>>(ma, mb, mas[1], mas[2], mas[3], bias, out, p);
  return check_launch("conv_tc_halo_kernel");
}

}  // namespace dpi
