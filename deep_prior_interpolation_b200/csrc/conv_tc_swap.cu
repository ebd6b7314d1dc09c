// ROLE-SWAPPED tcgen05 3x3x3 convolution for layers with FOUR output channels (forward and stride-1 dgrad).
//
// The MultiRes blocks open with wide-input, four-channel convs (mulresunet.py:70-79 with the default widths:
// Block3d(64 -> 4/8/13), the decoder's Block3d(67 -> 4/8/13), the final conv 25 -> 1, and the data gradient of every
// 4 -> 8 conv).  On the marching kernels (conv_tc_march.cu) the voxels are the M rows and Cout the N columns of the MMA:
// with N = 4 (padded to 16) a 128 x 16 x 8 MMA still occupies the tensor pipe for ~40 clk, so these layers ran at
// 1/8 of the MMA's width and were the largest launches of an iteration (67 -> 4 forward 997 us at 256x128x128 against
// 195 us of HBM time).
//
// Here the roles are swapped: the M rows are the (tap, cout) pairs - 27 x 4 = 108 of 128 rows, the whole filter bank
// resident in shared memory - and the N columns are the 180 voxels of an 18 x 10 input halo tile.  ONE MMA chain per
// input plane (K = Cin, N = 192: C/8 MMAs of ~96 clk instead of 27 x C/8 x 40 clk) leaves in TMEM
//     P[(tap, co)][u] = sum_c W[co][tap][c] * x[u][c]            for every input voxel u of the halo tile,
// i.e. the contribution of input voxel u to its 27 neighbours.  The eight epilogue warps move P through shared memory
// (tcgen05.ld -> st.shared, P[u][(tap, co)]: a 432-byte row per input voxel) and every thread, owning one output voxel
// (oh, ow) of the 16 x 8 tile and one pair of output channels, gathers its 27 float2 contributions:
//     out[d][oh][ow] = sum_{kd,kh,kw} P_{plane d + kd - 1}[(kd,kh,kw)][(oh + kh, ow + kw)]
// While the CTA marches along d the three output planes an input plane feeds live in three rotating float2
// accumulators per thread; the plane that received its last (kd = 2) contribution is stored (+ bias, += old value for
// an accumulating dgrad, fp64 BatchNorm partial sums as in conv_tc_march.cu).  The dgrad is the same kernel with the tap
// index mirrored (26 - tap).  Shared-memory traffic of the shuffle (608 + 432 wavefronts per 128 outputs) is what
// bounds the kernel, at about the HBM time of the input.
//
// Pipeline: warp 0 = TMA producer (one 18 x 10 halo box per plane and channel chunk, ring of stages), warp 1 = MMA
// issuer (whole-warp control flow, elected lane issues), warps 2-9 = epilogue; two accumulators of 192 TMEM columns
// alternate between the MMA of plane d+1 and the drain of plane d.
#include <cuda.h>
#include <stdlib.h>
#include "conv_geom.cuh"

namespace dpi {
namespace swp {

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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major swizzled descriptor (layout 2/4/6 = SWIZZLE_128B/64B/32B), 8-row groups `sbo_bytes` apart
__device__ __forceinline__ uint64_t make_k_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

constexpr int TH = 16, TW = 8;               // output tile (h, w): one epilogue thread per voxel
constexpr int HH = TH + 2, WW = TW + 2;      // input halo tile
constexpr int kCols = HH * WW;               // 180 input voxels = used accumulator columns
constexpr int kMmaN = 192;                   // MMA N (multiple of 16); columns 180..191 are never read
constexpr int kBufCols = 256;                // TMEM columns per accumulator buffer (two buffers = all 512)
constexpr int kCout = 4;
constexpr int kTaps = 27;
constexpr int kRows = kTaps * kCout;         // 108 used M rows = floats per input voxel in the shuffle buffer
constexpr int kThreads = 320;               // TMA warp, MMA warp, eight epilogue warps
constexpr int kMaxStages = 8;
constexpr int kShuffleBytes = kCols * kRows * 4;   // 77 760

struct Params {
  int Do, Ho, Wo;
  int tiles_w, tiles_h;
  int C;
  int kc, rb, layout, n_chunks, ks_last;
  int stages, plane_bytes, wchunk_bytes;
  int seg_len, n_segs, n_units;
  uint32_t idesc;
  int64_t out_ld;
  int accumulate;
  void* stats;                               // STATS kernels: stats workspace receiving one partial row per CTA
};

__device__ __forceinline__ void add2(float2& a, const float2 v) { a.x += v.x; a.y += v.y; }

// contributions of relation J (input plane -> output plane pz - J) to the thread's (voxel, channel pair): nine float2 of
// the shuffle buffer, summed in a fixed order
template <int J, bool TR>
__device__ __forceinline__ void gather9(float2& acc, const float* sp) {
#pragma unroll
  for (int jh = 0; jh < 3; ++jh) {
#pragma unroll
    for (int jw = 0; jw < 3; ++jw) {
      const int idx = J * 9 + jh * 3 + jw;
      const int tap = TR ? 26 - idx : idx;
      add2(acc, *reinterpret_cast<const float2*>(sp + (jh * WW + jw) * kRows + tap * kCout));
    }
  }
}

template <bool STATS, bool TR>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_swap_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                    const float* __restrict__ bias, float* __restrict__ out, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t wbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xbase = wbase + (uint32_t)p.n_chunks * (uint32_t)p.wchunk_bytes;
  const uint32_t sbase = xbase + (uint32_t)p.stages * (uint32_t)p.plane_bytes;
  const uint32_t bar_base = sbase + (uint32_t)kShuffleBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t w_full = bar_base + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (2 * kMaxStages + 1 + b); };
  auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (2 * kMaxStages + 3 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(w_full, 1);
    for (uint32_t b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * kBufCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_d;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer: the filter bank once, then one halo box per (input plane, chunk) =================
    if (elect_one()) {
      // box [kc channels][4 cout][27 taps] of the packed weights Wp[n][tap][c]: shared-memory row = tap * 4 + cout
      mbar_expect_tx(w_full, (uint32_t)p.n_chunks * (uint32_t)(kRows * p.rb));
      for (int c = 0; c < p.n_chunks; ++c)
        tma_load_3d(wbase + (uint32_t)c * (uint32_t)p.wchunk_bytes, &tma_w, w_full, c * p.kc, 0, 0);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 1;
    uint32_t dst = xbase;
    const uint32_t tx_bytes = (uint32_t)(kCols * p.rb);
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int w0 = tw * TW - 1, h0 = th * TH - 1, d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      int dz = d_lo - 1;
      for (int pz = 0; pz < L + 2; ++pz, ++dz) {
        int c0 = 0;
        for (int c = 0; c < p.n_chunks; ++c, c0 += p.kc) {
          mbar_wait(empty_bar(s), ph);
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), tx_bytes);
            tma_load_4d(dst, &tma_x, full_bar(s), c0, w0, h0, dz);
          }
          __syncwarp();
          dst += (uint32_t)p.plane_bytes;
          if (++s == p.stages) { s = 0; ph ^= 1u; dst = xbase; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole-warp control flow, one elected lane issues: back-to-back UTCHMMA) =========
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint64_t wdesc0 = make_k_desc(wbase, 8 * p.rb, p.layout);
    const uint64_t xdesc0 = make_k_desc(xbase, 8 * p.rb, p.layout);
    const uint64_t plane_u = (uint64_t)((uint32_t)p.plane_bytes >> 4);
    const uint64_t wchunk_u = (uint64_t)((uint32_t)p.wchunk_bytes >> 4);
    const int ks_full = p.kc >> 3;
    int s = 0;
    uint32_t ph = 0, pc = 0;
    uint64_t xd = xdesc0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int seg = u / (p.tiles_w * p.tiles_h);
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int pz = 0; pz < L + 2; ++pz, ++pc) {
        const uint32_t buf = pc & 1u;
        mbar_wait(tempty_bar(buf), ((pc >> 1) & 1u) ^ 1u);       // drained by the epilogue two planes ago
        tc_fence_after();
        const uint32_t dcol = tmem_d + buf * (uint32_t)kBufCols;
        uint64_t wd = wdesc0;
        for (int c = 0; c < p.n_chunks; ++c) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const int ks = c == p.n_chunks - 1 ? p.ks_last : ks_full;
            umma_tf32(dcol, wd, xd, p.idesc, c == 0 ? 0u : 1u);
            if (ks > 1) umma_tf32(dcol, wd + 2, xd + 2, p.idesc, 1u);
            if (ks > 2) umma_tf32(dcol, wd + 4, xd + 4, p.idesc, 1u);
            if (ks > 3) umma_tf32(dcol, wd + 6, xd + 6, p.idesc, 1u);
            umma_commit(empty_bar(s));
            if (c == p.n_chunks - 1) umma_commit(tfull_bar(buf));
          }
          __syncwarp();
          wd += wchunk_u;
          xd += plane_u;
          if (++s == p.stages) { s = 0; ph ^= 1u; xd = xdesc0; }
        }
      }
    }
  } else {
    // ================= epilogue: TMEM -> shuffle buffer -> per-voxel gather -> global =================
    // Eight warps, two per TMEM lane quarter (each moves 96 of the 192 columns); in the gather a thread owns one
    // output voxel and one PAIR of output channels.
    const int e = (int)threadIdx.x - 64;                 // 0..255
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = e >> 7;                             // column half this warp moves
    const int row = q * 32 + lane;                       // M row = tap * 4 + cout
    // gather role: 16 consecutive lanes = 8 voxels (ow) x 2 channel pairs, so that the 8-byte shared-memory reads of a
    // half warp fall into 32 different banks (voxel pitch 108 floats = 12 banks) and its stores cover 128 contiguous bytes
    const int ow = e & 7, pair = (e >> 3) & 1, oh = e >> 4;
    float* const S = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)));
    float* const s_wr = S + (half * (kMmaN / 2)) * kRows + row;
    const float* const s_rd = S + (oh * WW + ow) * kRows + 2 * pair;
    const bool accumulate = !STATS && p.accumulate;
    float2 b2 = make_float2(0.f, 0.f);
    if (bias) b2 = __ldg(reinterpret_cast<const float2*>(bias + 2 * pair));
    double st_s[STATS ? 2 : 1], st_q[STATS ? 2 : 1];
#pragma unroll
    for (int j = 0; j < (STATS ? 2 : 1); ++j) st_s[j] = st_q[j] = 0.0;
    uint32_t pc = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int tt = u;
      const int tw = tt % p.tiles_w; tt /= p.tiles_w;
      const int th = tt % p.tiles_h;
      const int seg = tt / p.tiles_h;
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      const int gw = tw * TW + ow, gh = th * TH + oh;
      const bool valid = gw < p.Wo && gh < p.Ho;
      float* orow = out + (((int64_t)d_lo * p.Ho + gh) * p.Wo + gw) * p.out_ld + 2 * pair;
      const int64_t plane_stride = (int64_t)p.Ho * p.Wo * p.out_ld;
      float2 a_new = make_float2(0.f, 0.f), a_mid = a_new, a_old = a_new;
      for (int pz = 0; pz < L + 2; ++pz, ++pc) {
        const uint32_t buf = pc & 1u;
        // accumulating dgrad: fetch the previous value of the plane completed by this input plane ahead of the wait
        float2 old = make_float2(0.f, 0.f);
        if (accumulate && valid && pz >= 2) old = *reinterpret_cast<const float2*>(orow);
        mbar_wait(tfull_bar(buf), (pc >> 1) & 1u);
        tc_fence_after();
        const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)kBufCols + (uint32_t)(half * (kMmaN / 2));
        // three loads of 32 columns; load k+1 is in flight while the values of load k go to shared memory
        uint32_t ra[32], rb[32];
        tmem_ld32_nowait(tbase, ra);
        tmem_ld_wait();
        tmem_ld32_nowait(tbase + 32u, rb);
        if (row < kRows) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s_wr[i * kRows] = __uint_as_float(ra[i]);
        }
        tmem_ld_wait();
        tmem_ld32_nowait(tbase + 64u, ra);
        if (row < kRows) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s_wr[(32 + i) * kRows] = __uint_as_float(rb[i]);
        }
        tmem_ld_wait();
        tc_fence_before();
        if (row < kRows) {
          // (the last 12 columns of the second half lie past the 180 voxels of the halo tile)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (half == 0 || 64 + i < kCols - kMmaN / 2) s_wr[(64 + i) * kRows] = __uint_as_float(ra[i]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));     // the MMA of plane pz + 2 may overwrite this accumulator
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // relation J: this input plane (pz) feeds output plane pz - J of the segment
        if (pz < L) gather9<0, TR>(a_new, s_rd);
        if (pz >= 1 && pz <= L) gather9<1, TR>(a_mid, s_rd);
        if (pz >= 2) {
          gather9<2, TR>(a_old, s_rd);
          if (valid) {
            float2 v = a_old;
            add2(v, b2);
            if (accumulate) add2(v, old);
            *reinterpret_cast<float2*>(orow) = v;
            if constexpr (STATS) {
              double d;
              d = (double)v.x; st_s[0] += d; st_q[0] = fma(d, d, st_q[0]);
              d = (double)v.y; st_s[1] += d; st_q[1] = fma(d, d, st_q[1]);
            }
          }
          orow += plane_stride;
        }
        a_old = a_mid;
        a_mid = a_new;
        a_new = make_float2(0.f, 0.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // every gather of this plane is done: S may be overwritten
      }
    }
    if constexpr (STATS) {
      // one partial row per CTA in the stats-workspace format of elementwise.cu (header {rows, C}; rows of [2][C]
      // doubles): xor butterfly over the 16 lanes of a warp that hold the same channel pair (lane bit 3 = pair), then the
      // eight warps in a fixed order
      double* sred = reinterpret_cast<double*>(S);
      const int we = e >> 5;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double a = st_s[j], b = st_q[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          if (o == 8) continue;
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if ((lane & 23) == 0) {                          // lanes 0 (pair 0) and 8 (pair 1)
          sred[we * 8 + 2 * pair + j] = a;               // [warp][kind][channel]
          sred[we * 8 + 4 + 2 * pair + j] = b;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int64_t* header = reinterpret_cast<int64_t*>(p.stats);
      double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(p.stats) + 16);
      if (e < 8) {
        double v = sred[e];
#pragma unroll
        for (int w = 1; w < 8; ++w) v += sred[w * 8 + e];
        partial[(size_t)blockIdx.x * 2 * kCout + e] = v;   // [2][4]: sums, then sums of squares
      }
      if (blockIdx.x == 0 && e == 0) {
        header[0] = gridDim.x;
        header[1] = kCout;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(2 * kBufCols) : "memory");
  }
}

// ---------------------------------------------------------------- host side -----------------------------------
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

// persistent CTAs: the SM count, or DPI_TC_MARCH_CTAS (the test knob of conv_tc_march.cu: several units per CTA)
static int cta_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  const char* e = getenv("DPI_TC_MARCH_CTAS");
  if (e && e[0]) {
    const int v = atoi(e);
    if (v > 0) return v;
  }
  return n;
}

constexpr int kBarBytes = 8 * (2 * kMaxStages + 6) + 16;

static bool plan(const GatherGeom& g, Params& p, size_t* smem_out) {
  if (g.N != kCout || g.kd != 3 || g.kh != 3 || g.kw != 3 || g.sd != 1 || g.sh != 1 || g.sw != 1 || g.pd != 1 ||
      g.ph != 1 || g.pw != 1 || (g.C & 3) || g.C < 4 || g.thin_c > 0)
    return false;
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C;
  p.tiles_w = (g.Wo + TW - 1) / TW;
  p.tiles_h = (g.Ho + TH - 1) / TH;
  p.kc = g.C <= 8 ? 8 : (g.C <= 16 ? 16 : 32);
  p.rb = p.kc * 4;
  p.layout = p.kc == 32 ? 2 : (p.kc == 16 ? 4 : 6);
  p.n_chunks = (g.C + p.kc - 1) / p.kc;
  p.ks_last = ((g.C - (p.n_chunks - 1) * p.kc) + 7) >> 3;
  p.plane_bytes = (kCols * p.rb + 1023) / 1024 * 1024;
  p.wchunk_bytes = 128 * p.rb;
  const int64_t avail = (int64_t)tc_smem_budget(false) - 1024 - kBarBytes - kShuffleBytes - (int64_t)p.n_chunks * p.wchunk_bytes;
  // the MMA reads 192 rows of a stage: the 12 rows past the box must still lie inside the dynamic allocation
  const int64_t tail = (int64_t)kMmaN * p.rb - p.plane_bytes;
  int stages = (int)(avail / p.plane_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2 || tail > kShuffleBytes) return false;
  p.stages = stages;
  // segments of output planes: minimise (rounds over the CTAs) x (planes a unit streams)
  const int nsm = cta_count();
  const int ncol = p.tiles_w * p.tiles_h;
  double best = 1e30;
  p.seg_len = g.Do; p.n_segs = 1;
  for (int want = 1; want <= g.Do; ++want) {
    const int len = (g.Do + want - 1) / want;
    const int segs = (g.Do + len - 1) / len;
    const int64_t units = (int64_t)ncol * segs;
    const int64_t rounds = (units + nsm - 1) / nsm;
    const double cost = (double)rounds * (len + 2 + 0.75);
    if (cost < best - 1e-9) { best = cost; p.seg_len = len; p.n_segs = segs; }
  }
  p.n_units = ncol * p.n_segs;
  // UMMA instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kMmaN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  *smem_out = (size_t)p.n_chunks * p.wchunk_bytes + (size_t)p.stages * p.plane_bytes + kShuffleBytes + kBarBytes + 1024;
  return true;
}

static CUtensorMapSwizzle swizzle_for_row_bytes(int rb) {
  return rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

template <bool STATS, bool TR>
static int launch_t(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, float* out, const Params& p,
                    size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_swap_kernel<STATS, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      set_error("swap: cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const int nsm = cta_count();
  const unsigned grid = (unsigned)(p.n_units < nsm ? p.n_units : nsm);
  conv_tc_swap_kernel<STATS, TR><<<grid, kThreads, smem, st>>>(mx, mw, bias, out, p);
  return check_launch("conv_tc_swap_kernel");
}

static bool enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DPI_TC_SWAP");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

}  // namespace swp

int conv_tc_swap_supported(const GatherGeom& g) {
  swp::Params p;
  size_t smem = 0;
  return swp::enabled() && swp::get_encode() && swp::plan(g, p, &smem) ? 1 : 0;
}

int conv_tc_swap_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                        const GatherGeom& g, int accumulate, cudaStream_t st) {
  using namespace swp;
  if (!enabled()) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  Params p;
  size_t smem = 0;
  if (!plan(g, p, &smem)) return DPI_ERR_UNSUPPORTED;
  p.out_ld = out_ld;
  p.accumulate = accumulate;
  p.stats = nullptr;
  if (!g.transposed && !accumulate) {
    // fused BatchNorm statistics requested by dpi_conv_fwd_stats (conv_simt.cu)
    StatsRequest* rq = stats_request();
    const char* e = getenv("DPI_TC_FUSED_STATS");
    if (rq && rq->ws && !rq->done && !(e && e[0] == '0')) {
      rq->done = true;
      p.stats = rq->ws;
    }
  }
  CUtensorMap mx, mw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)in_ld * 4, (cuuint64_t)g.Wi * in_ld * 4, (cuuint64_t)g.Hi * g.Wi * in_ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)WW, (cuuint32_t)HH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(p.rb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("swap: cuTensorMapEncodeTiled(x) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  {
    // packed weights Wp[n][tap][c] viewed as (c, n, tap)
    cuuint64_t dims[3] = {(cuuint64_t)g.C, (cuuint64_t)kCout, (cuuint64_t)kTaps};
    cuuint64_t strides[2] = {(cuuint64_t)kTaps * g.C * 4, (cuuint64_t)g.C * 4};
    cuuint32_t box[3] = {(cuuint32_t)p.kc, (cuuint32_t)kCout, (cuuint32_t)kTaps};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(Wp), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(p.rb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("swap: cuTensorMapEncodeTiled(w) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  if (g.transposed) return launch_t<false, true>(mx, mw, bias, out, p, smem, st);
  return p.stats ? launch_t<true, false>(mx, mw, bias, out, p, smem, st)
                 : launch_t<false, false>(mx, mw, bias, out, p, smem, st);
}

}  // namespace dpi
