// Persistent "column-marching" tcgen05 3x3(x3) convolution (forward and stride-1 dgrad) with RESIDENT WEIGHTS.
//
// conv_tc_halo.cu starts one CTA per 128-voxel tile: every CTA pays barrier/TMEM set-up, a cold TMA round trip, a
// full reload of the 27 weight tiles (up to 4x the bytes of the activations it reads) and an epilogue nothing
// overlaps with.  For the full-resolution layers (K = 27 x 4..72, 27-108 MMAs per tile) that overhead is 2-10x the
// MMA time (profiles/r1_op_profile_a.txt: dgrad 4->72 channels 1917 us against an MMA floor of 195 us).
//
// Here a CTA is persistent (grid = #SMs) and owns work units "(h,w) tile column x segment of output planes":
//   * the packed weights of ALL taps and channel chunks are loaded ONCE per CTA and stay in shared memory;
//   * the CTA marches along d: input plane dz (18 x 10 halo rows per channel chunk, one TMA box) is loaded ONCE and
//     feeds the three output planes dz-1, dz, dz+1 (kd = 2, 1, 0) - L2 -> SMEM traffic drops another 3x;
//   * four accumulators rotate in TMEM (slot = output-plane counter mod 4): while the MMA thread fills planes
//     do+1 .. do+2 the four epilogue warps drain plane do (tcgen05.ld -> +bias / += -> global);
//   * TMA producer, MMA issuer and epilogue run decoupled across unit boundaries (mbarrier rings), so the tensor
//     pipe never waits for a tile to start.
// The nine (kh,kw) taps of a plane are row-shifted K-major swizzled descriptor views exactly as in conv_tc_halo.cu.
#include <cuda.h>
#include <stdlib.h>
#include <vector>
#include "conv_geom.cuh"

namespace dpi {
namespace march {

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
constexpr int kSlots = 16;                  // upper bound of rotating TMEM accumulators (4 / 8 / 16 in use)
constexpr int kMaxStages = 8;
constexpr int kMaxBN = 128;                 // 4 slots x 128 columns = all of TMEM

// ---- BatchNorm statistics fused into the epilogue ------------------------------------------------------------------
// A forward conv that feeds a training-mode BatchNorm (base.py:162-166,211-216) also needs sum(y) and sum(y^2) per
// channel over all voxels.  Every epilogue thread owns one (h, w) position of its tile column for the whole kernel, so
// it keeps fp64 running sums of exactly the fp32 values it stores (N <= 32 channels: 128 registers); at the end of the
// kernel the 128 threads are combined in a fixed order (xor-butterfly inside a warp, then warps 0..3) into ONE partial
// row per CTA, in the stats-workspace format of elementwise.cu (header {rows, C}; rows of [2][C] doubles), so the
// existing dpi_bn_finalize consumes it and the separate dpi_channel_stats pass over y disappears.
constexpr int kStatsMaxN = 32;
__device__ __forceinline__ void stats_add(double& s, double& q, float v) {
  const double d = (double)v;
  s += d;
  q = fma(d, d, q);
}
template <int NS>
__device__ __forceinline__ void stats_flush(const double (&s)[NS], const double (&q)[NS], int N, int wq,
                                            int lane, double* sred /* shared, 4 x 64 doubles */, void* ws_raw) {
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    double a = s[j], b = q[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      sred[wq * 64 + j] = a;
      sred[wq * 64 + 32 + j] = b;
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps only
  const int t = wq * 32 + lane;
  int64_t* header = reinterpret_cast<int64_t*>(ws_raw);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(ws_raw) + 16);
  if (t < 64) {
    const int c = t & 31;
    if (c < N) {
      const double v = ((sred[t] + sred[64 + t]) + sred[128 + t]) + sred[192 + t];
      partial[(size_t)blockIdx.x * 2 * N + (size_t)(t >> 5) * N + c] = v;
    }
  }
  if (blockIdx.x == 0 && t == 0) {
    header[0] = gridDim.x;
    header[1] = N;
  }
}

// All MMAs of one (input plane, channel chunk) stage.  The issuing lane is instruction-bound (ncu: samples spread
// evenly over UTCHMMA and the uniform-datapath descriptor arithmetic around it), so the loop order is
// tap -> k-step -> output plane: the A descriptor of (tap, k) is built once and shared by the (up to) three MMAs that
// feed the three output planes, and each B descriptor advances in place - ~1.3 uniform instructions per MMA.
template <int KS, bool ALL>
__device__ __forceinline__ void issue_stage(uint64_t ad0, const uint64_t (&aoff)[9], uint64_t (&bd)[3],
                                            const uint32_t (&dcol)[3], const bool (&vj)[3], uint64_t bstep,
                                            uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const uint64_t at = ad0 + aoff[tp];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const uint64_t a = at + (uint64_t)(2 * k);
      if (ALL || vj[0]) umma_tf32(dcol[0], a, bd[0] + (uint64_t)(2 * k), idesc, (tp == 0 && k == 0) ? acc0 : 1u);
      if (ALL || vj[1]) umma_tf32(dcol[1], a, bd[1] + (uint64_t)(2 * k), idesc, 1u);
      if (ALL || vj[2]) umma_tf32(dcol[2], a, bd[2] + (uint64_t)(2 * k), idesc, 1u);
    }
    bd[0] += bstep; bd[1] += bstep; bd[2] += bstep;
  }
}

// Parity-class mode (stride-2 data gradient), steady state: the participating relations / in-plane taps are fixed by
// the class, so they are compile-time here - no data-dependent branch between two MMAs (issue_stage_masked below is the
// generic form; its branch nest costs the issuing warp an instruction-fetch stall per MMA).  Slab order as laid out by
// conv_tc_march_dgrad_s2: relation j ascending (j = 1: kd' = 1, j = 2: kd' = 0), then in-plane tap ascending.
template <int CD, int CH, int CW>
__device__ __forceinline__ void issue_stage_class(int ks, uint64_t ad0, const uint64_t (&aoff)[9], uint64_t wd_c,
                                                  uint32_t slab_u, const uint32_t (&dcol)[3], uint32_t idesc, bool chunk0) {
  constexpr int NJ = CD ? 2 : 1, NH = CH ? 2 : 1, NW = CW ? 2 : 1;
#pragma unroll
  for (int jr = 0; jr < NJ; ++jr) {
#pragma unroll
    for (int hi = 0; hi < NH; ++hi) {
#pragma unroll
      for (int wi = 0; wi < NW; ++wi) {
        const int tp = (CH ? hi : 1) * 3 + (CW ? wi : 1);
        const uint64_t a = ad0 + aoff[tp];
        const uint64_t b = wd_c + (uint64_t)((uint32_t)(jr * NH * NW + hi * NW + wi) * slab_u);
        const uint32_t first = (jr == 0 && hi == 0 && wi == 0 && chunk0) ? 0u : 1u;
        if (ks == 4) {
          umma_tf32(dcol[1 + jr], a, b, idesc, first);
          umma_tf32(dcol[1 + jr], a + 2, b + 2, idesc, 1u);
          umma_tf32(dcol[1 + jr], a + 4, b + 4, idesc, 1u);
          umma_tf32(dcol[1 + jr], a + 6, b + 6, idesc, 1u);
        } else {
          umma_tf32(dcol[1 + jr], a, b, idesc, first);
          for (int k = 1; k < ks; ++k) umma_tf32(dcol[1 + jr], a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), idesc, 1u);
        }
      }
    }
  }
}

// Straight-line form of issue_stage_thin for the steady state (3-D, chunk 0, all three relations valid): the generic
// form below is a nest of data-dependent branches, and ncu showed the issuing warp spending its time in instruction
// fetch stalls at every branch target (~230 clk per MMA).  KTHIN / KSF = K-steps of the thin taps / of the centre tile.
template <int KTHIN, int KSF>
__device__ __forceinline__ void issue_stage_thin_fast(uint64_t ad0, const uint64_t (&aoff)[9], const uint64_t (&bt)[3],
                                                      uint64_t tstep, uint64_t bc, const uint32_t (&dcol)[3],
                                                      uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const uint64_t at = ad0 + aoff[tp];
    const uint64_t bo = (uint64_t)tp * tstep;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (tp == 4 && j == 1) continue;
#pragma unroll
      for (int k = 0; k < KTHIN; ++k)
        umma_tf32(dcol[j], at + (uint64_t)(2 * k), bt[j] + bo + (uint64_t)(2 * k), idesc,
                  (j == 0 && tp == 0 && k == 0) ? acc0 : 1u);
    }
  }
  const uint64_t at = ad0 + aoff[4];
#pragma unroll
  for (int k = 0; k < KSF; ++k) umma_tf32(dcol[1], at + (uint64_t)(2 * k), bc + (uint64_t)(2 * k), idesc, 1u);
}

// Fused data gradient (thin_c > 0).  Chunk 0, every tap: the first kthin K-steps from the thin slabs (bt[j] = thin tile
// (kd(j), tap 0), advancing by tstep per tap); every chunk, centre tap of relation jc only: all K-steps of the chunk from
// its full centre tile bc.  (tap 4, relation jc) is NOT taken from the thin slab: the centre tile holds those K-steps too.
__device__ __forceinline__ void issue_stage_thin(bool chunk0, int kthin, int ks_full, int jc, uint64_t ad0,
                                                 const uint64_t (&aoff)[9], uint64_t (&bt)[3], uint64_t tstep, uint64_t bc,
                                                 const uint32_t (&dcol)[3], const bool (&vj)[3], uint32_t idesc,
                                                 uint32_t acc0) {
  if (chunk0) {
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const uint64_t at = ad0 + aoff[tp];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (vj[j] && !(tp == 4 && j == jc)) {
          for (int k = 0; k < kthin; ++k)
            umma_tf32(dcol[j], at + (uint64_t)(2 * k), bt[j] + (uint64_t)(2 * k), idesc,
                      (j == 0 && tp == 0 && k == 0) ? acc0 : 1u);
        }
      }
      bt[0] += tstep; bt[1] += tstep; bt[2] += tstep;
    }
  }
  // (static indices only: a run-time index into dcol / vj would move those arrays - shared with the regular issue
  //  paths - to local memory, which slowed EVERY march launch by 10-40 %)
  const uint64_t at = ad0 + aoff[4];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j == jc && vj[j]) {
      for (int k = 0; k < ks_full; ++k) umma_tf32(dcol[j], at + (uint64_t)(2 * k), bc + (uint64_t)(2 * k), idesc, 1u);
    }
  }
}

struct Params;
__device__ __forceinline__ void issue_stage_masked(int ks, uint64_t ad0, const uint64_t (&aoff)[9], uint64_t wd_c,
                                                   uint32_t slab_u, const uint32_t (&dcol)[3], const bool (&vj)[3],
                                                   const Params& p, bool chunk0);

struct Params {
  int Do, Ho, Wo;
  int tiles_w, tiles_h;
  int C, N, nkd, pd, transposed;
  int kc, rb, layout, n_chunks;
  int BN, stages;
  int plane_bytes;              // 180 rows x rb, rounded up to 1024
  int wslab_bytes;              // 9 * BN * rb, rounded up to 1024: the nine (kh,kw) tiles of one (chunk, kd)
  int seg_len, n_segs, n_units;
  int slot_shift;               // log2(#accumulator slots): 3 are being accumulated, the rest is epilogue slack
  uint32_t idesc, tmem_cols;
  int64_t out_ld;
  int accumulate;
  // ---- output addressing: out voxel = u * os + oc per axis (plain conv: os = 1, oc = 0, full dims = Do,Ho,Wo) ----
  int osd, os, ocd, och, ocw;
  int OD, OH, OW;
  // ---- strided-dgrad (parity class) mode: only a subset of the 27 u-space taps exists; its weight tiles are the
  //      only ones kept in shared memory (slab r = original tap orig_tap[r]) ----
  int halo;                     // 1: 18 x 10 halo planes (3x3 taps); 0: bare 16 x 8 tiles (1x1 convs)
  int masked;
  int jmask;                    // bit j: relation j (this input plane -> output plane pz - j) takes part
  int tmask;                    // bit tp: in-plane tap tp = kh'*3 + kw' takes part
  int ntp, nslab;               // valid in-plane taps; valid (j, tap) pairs = weight slabs per chunk
  int jrank[3], tprank[9];
  int orig_tap[8];
  int wregion_bytes;            // all resident weight tiles
  void* stats;                  // STATS kernels: stats workspace receiving one partial row per CTA
  int thin_c;                   // > 0 (fused dgrad, GatherGeom::thin_c): reduction channels >= thin_c are multiplied at
                                // the centre tap only - their weights are zero everywhere else.  Resident weights are
                                // then laid out THIN: nkd slabs of nine [BN x t_rb] tiles holding the first kthin K-steps
                                // of every tap (own, narrower swizzle), followed by one full [BN x rb] centre-tap tile
                                // per channel chunk - 27 x BN x 128 B of mostly zeros would not fit
  int kthin, t_rb, t_layout, t_slab_bytes, c_tile_bytes;
  int tma_store;                // 1: the epilogue stages each 16 x 8 x N output tile in shared memory and hands it to the
                                // TMA (cp.async.bulk.tensor store, or cp.reduce ... add.f32 when accumulating) instead of
                                // 128 threads writing 16-byte pieces at an N*4-byte pitch.  Measured on the output
                                // pattern alone (profiles/r2_probe_epilogue_store.txt, N = 72): 494 -> 183 us per 1.2 GB;
                                // pays from N ~ 40 up
  int ctas_per_sm;              // 2 for thin layers (BN = 16, small shared-memory footprint): grid = 2 x #SMs
  int debug;                    // DPI_TC_MARCH_DEBUG bit mask (timing experiments only, results are wrong):
                                // 1 = no plane TMA after the first ring fill, 2 = epilogue skips TMEM/global traffic,
                                // 4 = no MMAs
};

// parity-class mode: few taps, generic loops (the MMA count is tiny here).  The accumulator of a fresh output plane is
// first written by the smallest participating relation j at chunk 0.
__device__ __forceinline__ void issue_stage_masked(int ks, uint64_t ad0, const uint64_t (&aoff)[9], uint64_t wd_c,
                                                   uint32_t slab_u, const uint32_t (&dcol)[3], const bool (&vj)[3],
                                                   const Params& p, bool chunk0) {
  const int jmin = __ffs(p.jmask) - 1;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (!vj[j]) continue;
    uint32_t acc = (j == jmin && chunk0) ? 0u : 1u;
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      if (!((p.tmask >> tp) & 1)) continue;
      const uint64_t a = ad0 + aoff[tp];
      const uint64_t b = wd_c + (uint64_t)((uint32_t)(p.jrank[j] * p.ntp + p.tprank[tp]) * slab_u);
      for (int k = 0; k < ks; ++k) {
        umma_tf32(dcol[j], a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), p.idesc, acc);
        acc = 1u;
      }
    }
  }
}

// MAXBN = widest accumulator the instantiation handles.  The MAXBN = 16 instantiations (thin layers: Cout <= 16) stay
// under 170 registers so that TWO CTAs share an SM (Params::ctas_per_sm): such layers are bound by the issuing warp's
// per-plane path plus 27 MMAs of >= 39 clk, which do not overlap within one CTA (profiles/r1_march_bottleneck_isolation.txt).
template <bool STATS, int MAXBN>
__global__ void __launch_bounds__(kThreads, (MAXBN <= 16 || (MAXBN <= 32 && !STATS)) ? 2 : 1)
conv_tc_march_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_b2, const __grid_constant__ CUtensorMap tma_c,
                     const float* __restrict__ bias, float* __restrict__ out, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t wbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t abase = wbase + (uint32_t)p.wregion_bytes;
  const uint32_t bar_base = abase + (uint32_t)p.stages * (uint32_t)p.plane_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t w_full = bar_base + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 1 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 1 + kSlots + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1 + 2 * kSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(w_full, 1);
    for (int s = 0; s < (1 << p.slot_shift); ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
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

  const int nplanes_extra = p.nkd - 1;

  if (warp == 0) {
    // ================= TMA producer: weights once, then one halo plane per (input plane, chunk) =================
    // whole-warp control flow + one elected lane (as for the MMA issuer): a lean scalar path matters here too - with
    // 27 MMAs per plane the producer has ~1000 clk per stage, and runtime div/mod alone cost more than that
    if (elect_one()) {
      if (p.thin_c > 0) {
        // thin slabs (tma_b: box [8 kthin channels][BN][9 taps]) + one full centre-tap tile per chunk (tma_b2)
        mbar_expect_tx(w_full, (uint32_t)p.nkd * (uint32_t)(9 * p.BN * p.t_rb) + (uint32_t)p.n_chunks * (uint32_t)(p.BN * p.rb));
        for (int kd = 0; kd < p.nkd; ++kd)
          tma_load_3d(wbase + (uint32_t)kd * (uint32_t)p.t_slab_bytes, &tma_b, w_full, 0, 0, kd * 9);
        for (int c = 0; c < p.n_chunks; ++c)
          tma_load_3d(wbase + (uint32_t)(p.nkd * p.t_slab_bytes) + (uint32_t)c * (uint32_t)p.c_tile_bytes, &tma_b2, w_full,
                      c * p.kc, 0, (p.nkd / 2) * 9 + 4);
      } else if (!p.masked) {
        mbar_expect_tx(w_full, (uint32_t)(p.n_chunks * p.nkd) * (uint32_t)(9 * p.BN * p.rb));
        for (int c = 0; c < p.n_chunks; ++c)
          for (int kd = 0; kd < p.nkd; ++kd)
            tma_load_3d(wbase + (uint32_t)(c * p.nkd + kd) * (uint32_t)p.wslab_bytes, &tma_b, w_full, c * p.kc, 0, kd * 9);
      } else {
        // one [BN][kc] tile per existing tap (tma_b has a one-tap box in this mode)
        mbar_expect_tx(w_full, (uint32_t)(p.n_chunks * p.nslab) * (uint32_t)(p.BN * p.rb));
        for (int c = 0; c < p.n_chunks; ++c)
          for (int r = 0; r < p.nslab; ++r)
            tma_load_3d(wbase + (uint32_t)(c * p.nslab + r) * (uint32_t)p.wslab_bytes, &tma_b, w_full, c * p.kc, 0,
                        p.orig_tap[r]);
      }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 1;                                                // parity of the "stage was never used" wait
    uint32_t a_dst = abase;
    bool ring1 = true;
    const uint32_t tx_bytes = (uint32_t)((TH + 2 * p.halo) * (TW + 2 * p.halo) * p.rb);
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int w0 = tw * TW - p.halo, h0 = th * TH - p.halo, d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      int dz = d_lo - p.pd;
      for (int pz = 0; pz < L + nplanes_extra; ++pz, ++dz) {
        int c0 = 0;
        for (int c = 0; c < p.n_chunks; ++c, c0 += p.kc) {
          mbar_wait(empty_bar(s), ph);
          if (elect_one()) {
            if ((p.debug & 1) && !ring1) {
              mbar_arrive(full_bar(s));
            } else {
              mbar_expect_tx(full_bar(s), tx_bytes);
              tma_load_4d(a_dst, &tma_a, full_bar(s), c0, w0, h0, dz);
            }
          }
          __syncwarp();
          a_dst += (uint32_t)p.plane_bytes;
          if (++s == p.stages) { s = 0; ph ^= 1u; a_dst = abase; ring1 = false; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The WHOLE warp runs the control flow (all operands are warp-uniform) and one elected lane issues each batch:
    // ptxas then emits back-to-back UTCHMMA on uniform registers.  Issuing from an `if (lane == 0)` region instead
    // makes it wrap EVERY tcgen05.mma in an ELECT/BRA.ANY waterfall loop (~10 instructions, ~45-60 clk per MMA),
    // which was the real "issue floor" of the earlier kernels (scratch/umma_rate4.cu vs umma_rate3.cu).
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t ru = (uint32_t)(p.rb >> 4);                    // row pitch in 16-byte units
    const uint64_t bstep = (uint64_t)((uint32_t)p.BN * ru);       // one tap = BN rows (16-byte units)
    const uint32_t wslab_u = (uint32_t)p.wslab_bytes >> 4;
    const uint64_t wdesc0 = make_k_desc(wbase, 8 * p.rb, p.layout);
    // thin layout (fused dgrad): thin slabs with their own row pitch / swizzle, then the centre tiles
    const uint64_t tdesc0 = make_k_desc(wbase, 8 * p.t_rb, p.t_layout);
    const uint64_t tstep = (uint64_t)((uint32_t)p.BN * (uint32_t)(p.t_rb >> 4));
    const uint32_t tslab_u = (uint32_t)p.t_slab_bytes >> 4;
    const uint64_t cdesc0 = make_k_desc(wbase + (uint32_t)(p.nkd * p.t_slab_bytes), 8 * p.rb, p.layout);
    const uint32_t ctile_u = (uint32_t)p.c_tile_bytes >> 4;
    // start-address offsets (16-byte units) of the nine row-shifted views of a halo plane:
    // forward: output (h, w) reads halo row (h + kh, w + kw); dgrad reads (h + 2 - kh, w + 2 - kw)
    uint64_t aoff[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const int kh = tp / 3, kw = tp - 3 * kh;
      aoff[tp] = p.halo ? (uint64_t)((uint32_t)(p.transposed ? ((2 - kh) * WW + (2 - kw)) : (kh * WW + kw)) * ru) : 0ull;
    }
    const uint32_t smask = (1u << p.slot_shift) - 1u;
    // Between the last MMA of one stage and the first MMA of the next the tensor pipe only has its (shallow) queue
    // to chew on, so the per-stage scalar path is kept to a few instructions: stage index / phase / descriptors are
    // carried incrementally (no div/mod), everything that depends on the plane only is hoisted out of the chunk loop.
    const uint64_t adesc0 = make_k_desc(abase, (TW + 2 * p.halo) * p.rb, p.layout);
    const uint64_t plane_u = (uint64_t)((uint32_t)p.plane_bytes >> 4);
    const uint64_t wchunk_u = (uint64_t)((uint32_t)p.nkd * wslab_u);
    const int ks_full = p.kc >> 3;
    const int ks_last = ((p.C - (p.n_chunks - 1) * p.kc) + 7) >> 3;
    int s = 0;
    uint32_t ph = 0;
    uint64_t ad_s = adesc0;
    uint32_t oc_base = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int seg = u / (p.tiles_w * p.tiles_h);
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int pz = 0; pz < L + nplanes_extra; ++pz) {
        if (pz < L) {
          // first touch of output plane pz: its TMEM slot must have been drained by the epilogue
          const uint32_t oc = oc_base + (uint32_t)pz;
          mbar_wait(tempty_bar((int)(oc & smask)), ((oc >> p.slot_shift) & 1u) ^ 1u);
          tc_fence_after();
        }
        // relation j: this input plane feeds output plane pz - j with weight slab kd = j (forward) / nkd-1-j (dgrad)
        uint64_t bd[3];
        uint32_t dcol[3];
        bool vj[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int do_rel = pz - j;
          vj[j] = j < p.nkd && do_rel >= 0 && do_rel < L && (!p.masked || ((p.jmask >> j) & 1));
          const int kd = p.transposed ? p.nkd - 1 - j : j;
          dcol[j] = tmem_d + ((oc_base + (uint32_t)do_rel) & smask) * (uint32_t)p.BN;
          bd[j] = wdesc0 + (uint64_t)((uint32_t)kd * wslab_u);
        }
        const bool all = vj[0] && vj[1] && vj[2] && !p.masked;
        const int dc = pz - nplanes_extra;                          // output plane completed by this input plane
        const uint32_t tfull_done = tfull_bar((int)((oc_base + (uint32_t)dc) & smask));
        for (int c = 0; c < p.n_chunks; ++c) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const int ksteps = c == p.n_chunks - 1 ? ks_last : ks_full;
            const uint32_t acc0 = c == 0 ? 0u : 1u;
            uint64_t bdc[3] = {bd[0], bd[1], bd[2]};
            if (p.debug & 4) {
            } else if (p.thin_c > 0) {
              uint64_t bt[3];
#pragma unroll
              for (int j = 0; j < 3; ++j)
                bt[j] = tdesc0 + (uint64_t)((uint32_t)(p.transposed ? p.nkd - 1 - j : j) * tslab_u);
              const uint64_t bc = cdesc0 + (uint64_t)((uint32_t)c * ctile_u);
              if (c == 0 && all && p.nkd == 3 && p.kthin == 1 && ksteps == 4)
                issue_stage_thin_fast<1, 4>(ad_s, aoff, bt, tstep, bc, dcol, p.idesc, acc0);
              else if (c == 0 && all && p.nkd == 3 && p.kthin == 2 && ksteps == 4)
                issue_stage_thin_fast<2, 4>(ad_s, aoff, bt, tstep, bc, dcol, p.idesc, acc0);
              else if (c == 0 && all && p.nkd == 3 && p.kthin == 1 && ksteps == 2)
                issue_stage_thin_fast<1, 2>(ad_s, aoff, bt, tstep, bc, dcol, p.idesc, acc0);
              else
                issue_stage_thin(c == 0, p.kthin, ksteps, p.nkd == 3 ? 1 : 0, ad_s, aoff, bt, tstep, bc, dcol, vj, p.idesc,
                                 acc0);
            } else if (p.masked) {
              const uint64_t wd_c = wdesc0 + (uint64_t)((uint32_t)(c * p.nslab) * wslab_u);
              if (!p.halo) {
                // 1x1 convs: one tap, one relation - straight-line per K-step count
                const uint32_t a0 = c == 0 ? 0u : 1u;
                umma_tf32(dcol[0], ad_s, wd_c, p.idesc, a0);
                if (ksteps > 1) umma_tf32(dcol[0], ad_s + 2, wd_c + 2, p.idesc, 1u);
                if (ksteps > 2) umma_tf32(dcol[0], ad_s + 4, wd_c + 4, p.idesc, 1u);
                if (ksteps > 3) umma_tf32(dcol[0], ad_s + 6, wd_c + 6, p.idesc, 1u);
              } else if (p.os == 2 && p.nkd == 3 && vj[1] && (p.ocd == 0 || vj[2])) {
                // stride-2 data gradient, all participating relations valid: the class-specialised straight-line form
                switch ((p.ocd << 2) | (p.och << 1) | p.ocw) {
                  case 0: issue_stage_class<0, 0, 0>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  case 1: issue_stage_class<0, 0, 1>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  case 2: issue_stage_class<0, 1, 0>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  case 3: issue_stage_class<0, 1, 1>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  case 4: issue_stage_class<1, 0, 0>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  case 5: issue_stage_class<1, 0, 1>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  case 6: issue_stage_class<1, 1, 0>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                  default: issue_stage_class<1, 1, 1>(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, p.idesc, c == 0); break;
                }
              } else {
                issue_stage_masked(ksteps, ad_s, aoff, wd_c, wslab_u, dcol, vj, p, c == 0);
              }
            } else if (all) {
              if (ksteps == 4) issue_stage<4, true>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
              else if (ksteps == 1) issue_stage<1, true>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
              else if (ksteps == 2) issue_stage<2, true>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
              else issue_stage<3, true>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
            } else {
              if (ksteps == 4) issue_stage<4, false>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
              else if (ksteps == 1) issue_stage<1, false>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
              else if (ksteps == 2) issue_stage<2, false>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
              else issue_stage<3, false>(ad_s, aoff, bdc, dcol, vj, bstep, p.idesc, acc0);
            }
            umma_commit(empty_bar(s));
            if (c == p.n_chunks - 1 && dc >= 0) umma_commit(tfull_done);
          }
          __syncwarp();
          bd[0] += wchunk_u; bd[1] += wchunk_u; bd[2] += wchunk_u;
          ad_s += plane_u;
          if (++s == p.stages) { s = 0; ph ^= 1u; ad_s = adesc0; }
        }
      }
      oc_base += (uint32_t)L;
    }
  } else {
    // ================= epilogue: drain accumulator slots in output-plane order =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t oc = 0;
    const bool accumulate = !STATS && p.accumulate;      // a stats-emitting launch is a forward conv: never accumulates
    // TMA-store form: two staging tiles [128 rows][N floats] behind the barriers; thread 64 (first epilogue thread)
    // issues the bulk stores
    const bool use_tma = !STATS && p.tma_store;
    const bool epi_leader = threadIdx.x == 64;
    const uint32_t stage_u32 = (tmem_slot + 16u + 127u) & ~127u;
    float* const stage = reinterpret_cast<float*>(smem_raw + (stage_u32 - smem_u32(smem_raw)));
    uint32_t sbuf = 0;
    constexpr int kNS = STATS ? (MAXBN < kStatsMaxN ? MAXBN : kStatsMaxN) : 1;
    double st_s[kNS], st_q[kNS];
#pragma unroll
    for (int j = 0; j < kNS; ++j) st_s[j] = st_q[j] = 0.0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      const int ow = (tw * TW + (row & 7)) * p.os + p.ocw, oh = (th * TH + (row >> 3)) * p.os + p.och;
      const bool valid_hw = ow < p.OW && oh < p.OH;
      float* orow = out + (((int64_t)(d_lo * p.osd + p.ocd) * p.OH + oh) * p.OW + ow) * p.out_ld;
      const int64_t plane_stride = (int64_t)p.osd * p.OH * p.OW * p.out_ld;
      for (int dr = 0; dr < L; ++dr, ++oc, orow += plane_stride) {
        const bool valid = valid_hw && (d_lo + dr) * p.osd + p.ocd < p.OD;
        const uint32_t slot = oc & ((1u << p.slot_shift) - 1u);
        // dgrad accumulation: fetch the previous gradient values BEFORE waiting for the accumulator, so their
        // global-memory latency overlaps the MMAs instead of serialising the epilogue (was 13 000 clk per plane
        // for the 4 -> 72 channel dgrad)
        float4 old[STATS ? 1 : MAXBN / 4];
        if (use_tma) {
          // the staging tile about to be overwritten was handed to the TMA two planes ago: wait until it has been READ
          if (epi_leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        } else if (!STATS && accumulate && valid) {
#pragma unroll
          for (int i = 0; i < MAXBN / 4; ++i)
            if (4 * i < p.N) old[STATS ? 0 : i] = *reinterpret_cast<const float4*>(orow + 4 * i);
        }
        float* const srow = stage + (size_t)sbuf * (size_t)(TH * TW) * (size_t)p.N + (size_t)row * (size_t)p.N;
        mbar_wait(tfull_bar((int)slot), (oc >> p.slot_shift) & 1u);
        tc_fence_after();
        const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.BN;
#pragma unroll
        for (int cc = 0; cc < MAXBN / 16; ++cc) {
          const int c = cc * 16;
          if (c < p.BN && !(p.debug & 2)) {
            uint32_t r[16];
            tmem_ld16_nowait(tbase + (uint32_t)c, r);
            tmem_ld_wait();
            if (valid || use_tma) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const int n = c + i;
                if (n < p.N) {
                  float4 v = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                         __uint_as_float(r[i + 3]));
                  if (bias) {   // (hoisting these loads into registers, as the packed kernel does, made this one slower)
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                    v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                  }
                  if (use_tma) {
                    *reinterpret_cast<float4*>(srow + n) = v;     // (rows outside the tensor are clipped by the TMA)
                    continue;
                  }
                  if (!STATS && accumulate) {
                    const float4 o4 = old[STATS ? 0 : cc * 4 + i / 4];
                    v.x += o4.x; v.y += o4.y; v.z += o4.z; v.w += o4.w;
                  }
                  *reinterpret_cast<float4*>(orow + n) = v;
                  if constexpr (STATS) if (cc < kNS / 16) {
                    constexpr int kMask = kNS - 1;             // (static indices: cc, i are unrolled)
                    stats_add(st_s[(cc * 16 + i) & kMask], st_q[(cc * 16 + i) & kMask], v.x);
                    stats_add(st_s[(cc * 16 + i + 1) & kMask], st_q[(cc * 16 + i + 1) & kMask], v.y);
                    stats_add(st_s[(cc * 16 + i + 2) & kMask], st_q[(cc * 16 + i + 2) & kMask], v.z);
                    stats_add(st_s[(cc * 16 + i + 3) & kMask], st_q[(cc * 16 + i + 3) & kMask], v.w);
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar((int)slot));
        if (use_tma) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (epi_leader) {
            const uint32_t src = stage_u32 + sbuf * (uint32_t)(TH * TW * 4) * (uint32_t)p.N;
            const int cw = tw * TW, ch = th * TH, cd = d_lo + dr;
            if (accumulate)
              asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tma_c)), "r"(src), "r"(0), "r"(cw), "r"(ch), "r"(cd) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tma_c)), "r"(src), "r"(0), "r"(cw), "r"(ch), "r"(cd) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          sbuf ^= 1u;
        }
      }
    }
    if (use_tma && epi_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if constexpr (STATS) {
      // every MMA of this CTA has completed (the last accumulator was drained), so the weight region is free
      double* sred = reinterpret_cast<double*>(smem_raw + (wbase - smem_u32(smem_raw)));
      stats_flush(st_s, st_q, p.N, q, lane, sred, p.stats);
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
// Packed variant: the three kd taps of a plane share ONE MMA.
//
// Input plane dz feeds output planes dz+1 (kd = 0), dz (kd = 1) and dz-1 (kd = 2) from the SAME A operand, so the three
// [BN x K] weight tiles are laid side by side in shared memory ([tap][kd-slot][BN rows]) and issued as one MMA with
// N = 3*BN whose accumulator columns are three neighbouring output-plane slots.  3x fewer MMAs, each 48 clk (N = 48)
// instead of 3 x 39 clk: the layers with Cout <= 16 were pinned at N/128 of the tensor peak by the per-MMA floor.
// To keep the slots of planes dz-1, dz, dz+1 contiguous without a wrapping ring, output planes are processed in groups
// of six or eight (PackedParams::group) that own a bank of as many slots (two banks alternate: the epilogue drains one while the other fills); the planes
// at a group border contribute with a narrower MMA (N = BN or 2*BN) and are loaded once more for the next group (8
// plane loads per 6 output planes).  One MMA has one accumulate flag for all its columns, so every MMA accumulates and
// the epilogue ZEROES a slot (tcgen05.st) after draining it.
constexpr int kBanks = 2;

struct PackedParams {
  int Do, Ho, Wo;
  int tiles_w, tiles_h;
  int C, N, transposed;
  int kc, rb, layout, n_chunks;
  int BN, stages;
  int plane_bytes;
  int wchunk_bytes;             // 27 tiles of BN x rb, rounded up to 1024
  int seg_len, n_segs, n_units;
  uint32_t idesc[3];            // N = BN, 2*BN, 3*BN
  uint32_t tmem_cols;
  int64_t out_ld;
  int accumulate;
  void* stats;                  // STATS kernels: stats workspace receiving one partial row per CTA
  int thin_c;                   // as in Params
  int ctas_per_sm;              // as in Params
  int group;                    // output planes per accumulator bank (6 or 8; two banks <= kSlots slots)
  long long* prof;              // DPI_TC_MARCH_PROF=1: per-CTA cycle counters of the three roles (debugging aid)
};

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// all MMAs of one (input plane, channel chunk) stage of the packed kernel: nine taps x KS K-steps, no branches
template <int KS>
__device__ __forceinline__ void issue_packed(uint32_t dcol, uint64_t ad_s, const uint64_t (&aoff)[9], uint64_t bd, uint64_t tstep,
                                             uint32_t idesc) {
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const uint64_t at = ad_s + aoff[tp];
    const uint64_t bt = bd + (uint64_t)tp * tstep;
#pragma unroll
    for (int k = 0; k < KS; ++k) umma_tf32(dcol, at + (uint64_t)(2 * k), bt + (uint64_t)(2 * k), idesc, 1u);
  }
}

// fused data gradient (thin_c > 0): KT K-steps at every tap, all KS at the centre tap
template <int KT, int KS>
__device__ __forceinline__ void issue_packed_thin(uint32_t dcol, uint64_t ad_s, const uint64_t (&aoff)[9], uint64_t bd,
                                                  uint64_t tstep, uint32_t idesc) {
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const uint64_t at = ad_s + aoff[tp];
    const uint64_t bt = bd + (uint64_t)tp * tstep;
#pragma unroll
    for (int k = 0; k < (tp == 4 ? KS : KT); ++k) umma_tf32(dcol, at + (uint64_t)(2 * k), bt + (uint64_t)(2 * k), idesc, 1u);
  }
}

template <bool STATS, int MAXBN>
__global__ void __launch_bounds__(kThreads, MAXBN <= 16 ? 2 : 1)
conv_tc_march_packed_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                            const float* __restrict__ bias, float* __restrict__ out, const PackedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t wbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t abase = wbase + (uint32_t)p.n_chunks * (uint32_t)p.wchunk_bytes;
  const uint32_t bar_base = abase + (uint32_t)p.stages * (uint32_t)p.plane_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t w_full = bar_base + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 1 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 1 + kSlots + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1 + 2 * kSlots);
  const int kNSlots = p.group * kBanks;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(w_full, 1);
    for (int s = 0; s < kNSlots; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
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
    if (elect_one()) {
      // weight tile (chunk c, tap tp, slot q): slot q pairs with the output plane that lies 2-q planes BEHIND the
      // input plane, i.e. relation j = 2-q; forward kd = j, dgrad kd = 2-j
      mbar_expect_tx(w_full, (uint32_t)(p.n_chunks * 27) * (uint32_t)(p.BN * p.rb));
      for (int c = 0; c < p.n_chunks; ++c)
        for (int tp = 0; tp < 9; ++tp)
          for (int q = 0; q < 3; ++q) {
            const int kd = p.transposed ? q : 2 - q;
            tma_load_3d(wbase + (uint32_t)c * (uint32_t)p.wchunk_bytes + (uint32_t)((tp * 3 + q) * p.BN * p.rb), &tma_b,
                        w_full, c * p.kc, 0, kd * 9 + tp);
          }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 1;
    uint32_t a_dst = abase;
    long long prof_acc[1] = {0};
    const uint32_t tx_bytes = (uint32_t)(HH * WW * p.rb);
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int w0 = tw * TW - 1, h0 = th * TH - 1, d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int g0 = 0; g0 < L; g0 += p.group) {
        const int n = min(p.group, L - g0);
        for (int pz = g0; pz < g0 + n + 2; ++pz) {
          const int dz = d_lo - 1 + pz;
          int c0 = 0;
          for (int c = 0; c < p.n_chunks; ++c, c0 += p.kc) {
            const long long tp0 = p.prof ? clock64() : 0;
            mbar_wait(empty_bar(s), ph);
            if (p.prof) prof_acc[0] += clock64() - tp0;
            if (elect_one()) {
              mbar_expect_tx(full_bar(s), tx_bytes);
              tma_load_4d(a_dst, &tma_a, full_bar(s), c0, w0, h0, dz);
            }
            __syncwarp();
            a_dst += (uint32_t)p.plane_bytes;
            if (++s == p.stages) { s = 0; ph ^= 1u; a_dst = abase; }
          }
        }
      }
    }
    if (p.prof && lane == 0) p.prof[blockIdx.x * 8 + 0] = prof_acc[0];
  } else if (warp == 1) {
    // ================= MMA issuer (whole-warp control flow, elected lane issues) =================
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t ru = (uint32_t)(p.rb >> 4);
    const uint64_t bn_u = (uint64_t)((uint32_t)p.BN * ru);          // one weight tile, in 16-byte units
    const uint64_t wchunk_u = (uint64_t)((uint32_t)p.wchunk_bytes >> 4);
    const uint64_t wdesc0 = make_k_desc(wbase, 8 * p.rb, p.layout);
    uint64_t aoff[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const int kh = tp / 3, kw = tp - 3 * kh;
      aoff[tp] = (uint64_t)((uint32_t)(p.transposed ? ((2 - kh) * WW + (2 - kw)) : (kh * WW + kw)) * ru);
    }
    const uint64_t adesc0 = make_k_desc(abase, WW * p.rb, p.layout);
    const uint64_t plane_u = (uint64_t)((uint32_t)p.plane_bytes >> 4);
    const int ks_full = p.kc >> 3;
    const int ks_last = ((p.C - (p.n_chunks - 1) * p.kc) + 7) >> 3;
    int s = 0;
    uint32_t ph = 0;
    uint64_t ad_s = adesc0;
    uint32_t par = 0;                          // bit s: parity of the next use of accumulator slot s
    uint32_t gi = 0;
    long long pw_tempty = 0, pw_full = 0, pw_issue = 0, pw_steps = 0;
    const long long pt_begin = p.prof ? clock64() : 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int seg = u / (p.tiles_w * p.tiles_h);
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      for (int g0 = 0; g0 < L; g0 += p.group, ++gi) {
        const int n = min(p.group, L - g0);
        const int slot0 = (int)(gi & 1u) * p.group;
        for (int pz = g0; pz < g0 + n + 2; ++pz) {
          if (pz < g0 + n) {
            // first contribution to output plane pz: its slot must have been drained and zeroed
            const int sl = slot0 + (pz - g0);
            const long long t0 = p.prof ? clock64() : 0;
            mbar_wait(tempty_bar(sl), (par >> sl) & 1u);
            tc_fence_after();
            if (p.prof) pw_tempty += clock64() - t0;
          }
          const int o_lo = max(pz - 2, g0), o_hi = min(pz, g0 + n - 1);
          const int nq = o_hi - o_lo + 1;
          const int q_lo = 2 - (pz - o_lo);
          const uint32_t dcol = tmem_d + (uint32_t)((slot0 + (o_lo - g0)) * p.BN);
          const uint32_t idesc = p.idesc[0] + ((uint32_t)((nq - 1) * p.BN >> 3) << 17);   // N = nq * BN
          const int o_done = pz - 2;                               // output plane completed by this input plane
          const uint32_t tfull_done = tfull_bar(slot0 + (o_done - g0));
          uint64_t bd_c = wdesc0 + (uint64_t)q_lo * bn_u;
          for (int c = 0; c < p.n_chunks; ++c) {
            const long long t1 = p.prof ? clock64() : 0;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const long long t2 = p.prof ? clock64() : 0;
            if (p.prof) { pw_full += t2 - t1; ++pw_steps; }
            if (elect_one()) {
              const int ksteps = c == p.n_chunks - 1 ? ks_last : ks_full;
              int ks_thin = ksteps;
              if (p.thin_c > 0) {
                ks_thin = (p.thin_c - c * p.kc + 7) >> 3;
                ks_thin = ks_thin < 0 ? 0 : (ks_thin > ksteps ? ksteps : ks_thin);
              }
              if (p.thin_c <= 0) {
                // straight-line issue per K-step count: the rolled form (a data-dependent branch nest per tap) cost the
                // issuing lane ~190 clk per MMA in instruction-fetch stalls - 1690 clk per plane-step for 9 MMAs
                // (DPI_TC_MARCH_PROF=1; profiles/r2_probe_umma_commit_cost.txt: 9 MMAs + 2 commits = 717 clk)
                if (ksteps == 4) issue_packed<4>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
                else if (ksteps == 1) issue_packed<1>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
                else if (ksteps == 2) issue_packed<2>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
                else issue_packed<3>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
              } else if (ks_thin == 2 && ksteps == 4) {
                issue_packed_thin<2, 4>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
              } else if (ks_thin == 1 && ksteps == 2) {
                issue_packed_thin<1, 2>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
              } else if (ks_thin == 1 && ksteps == 4) {
                issue_packed_thin<1, 4>(dcol, ad_s, aoff, bd_c, 3 * bn_u, idesc);
              } else {
                uint64_t bd = bd_c;
#pragma unroll
                for (int tp = 0; tp < 9; ++tp, bd += 3 * bn_u) {
                  const uint64_t at = ad_s + aoff[tp];
                  if (tp != 4) {
                    // fused dgrad: the 1x1 partner's channels only exist at the centre in-plane tap (and, inside its
                    // weight tiles, at the centre kd slot - the other slots hold zeros)
                    for (int k = 0; k < ks_thin; ++k) umma_tf32(dcol, at + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
                  } else {
                    for (int k = 0; k < ksteps; ++k) umma_tf32(dcol, at + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
                  }
                }
              }
              umma_commit(empty_bar(s));
              if (c == p.n_chunks - 1 && o_done >= g0) umma_commit(tfull_done);
            }
            __syncwarp();
            if (p.prof) pw_issue += clock64() - t2;
            bd_c += wchunk_u;
            ad_s += plane_u;
            if (++s == p.stages) { s = 0; ph ^= 1u; ad_s = adesc0; }
          }
        }
        par ^= ((1u << n) - 1u) << slot0;
      }
    }
    if (p.prof && lane == 0) {
      p.prof[blockIdx.x * 8 + 1] = pw_tempty; p.prof[blockIdx.x * 8 + 2] = pw_full; p.prof[blockIdx.x * 8 + 3] = pw_issue;
      p.prof[blockIdx.x * 8 + 4] = pw_steps; p.prof[blockIdx.x * 8 + 5] = clock64() - pt_begin;
    }
  } else {
    // ================= epilogue: drain + zero accumulator slots in output-plane order =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_d + ((uint32_t)(q * 32) << 16);
    const bool accumulate = !STATS && p.accumulate;
    constexpr int kNS = STATS ? (MAXBN < kStatsMaxN ? MAXBN : kStatsMaxN) : 1;
    double st_s[kNS], st_q[kNS];
#pragma unroll
    for (int j = 0; j < kNS; ++j) st_s[j] = st_q[j] = 0.0;
    // the bias lives in registers: loaded inside the plane loop (as it was) each __ldg sat on the drain -> zero -> "slot
    // empty" path the MMA warp waits for, 16 % of this kernel's stall samples (profiles/r2_ncu_packed_thin_8to16.txt)
    float4 bias_r[MAXBN / 4];
#pragma unroll
    for (int k = 0; k < MAXBN / 4; ++k)
      bias_r[k] = (bias && 4 * k < p.N) ? __ldg(reinterpret_cast<const float4*>(bias + 4 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    // initial state: every slot zero and "empty"
    for (int sl = 0; sl < kNSlots; ++sl) {
      for (int c = 0; c < p.BN; c += 16) tmem_st16_zero(lane_base + (uint32_t)(sl * p.BN + c));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(sl));
    }
    uint32_t par = 0, gi = 0;
    long long pe_wait = 0, pe_drain = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int t = u;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int seg = t / p.tiles_h;
      const int d_lo = seg * p.seg_len;
      const int L = min(p.seg_len, p.Do - d_lo);
      const int ow = tw * TW + (row & 7), oh = th * TH + (row >> 3);
      const bool valid = ow < p.Wo && oh < p.Ho;
      float* orow = out + (((int64_t)d_lo * p.Ho + oh) * p.Wo + ow) * p.out_ld;
      const int64_t plane_stride = (int64_t)p.Ho * p.Wo * p.out_ld;
      for (int g0 = 0; g0 < L; g0 += p.group, ++gi) {
        const int n = min(p.group, L - g0);
        const int slot0 = (int)(gi & 1u) * p.group;
        for (int i = 0; i < n; ++i, orow += plane_stride) {
          const int sl = slot0 + i;
          float4 old[STATS ? 1 : MAXBN / 4];
          if (!STATS && accumulate && valid) {
#pragma unroll
            for (int k = 0; k < MAXBN / 4; ++k)
              if (4 * k < p.N) old[STATS ? 0 : k] = *reinterpret_cast<const float4*>(orow + 4 * k);
          }
          const long long te0 = p.prof ? clock64() : 0;
          mbar_wait(tfull_bar(sl), (par >> sl) & 1u);
          tc_fence_after();
          const long long te1 = p.prof ? clock64() : 0;
          if (p.prof) pe_wait += te1 - te0;
          const uint32_t tbase = lane_base + (uint32_t)(sl * p.BN);
#pragma unroll
          for (int cc = 0; cc < MAXBN / 16; ++cc) {
            const int c = cc * 16;
            if (c < p.BN) {
              uint32_t r[16];
              tmem_ld16_nowait(tbase + (uint32_t)c, r);
              tmem_ld_wait();
              tmem_st16_zero(tbase + (uint32_t)c);
              if (valid) {
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                  const int nn = c + k;
                  if (nn < p.N) {
                    float4 v = make_float4(__uint_as_float(r[k]), __uint_as_float(r[k + 1]), __uint_as_float(r[k + 2]),
                                           __uint_as_float(r[k + 3]));
                    {
                      const float4 b4 = bias_r[cc * 4 + k / 4];
                      v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                    }
                    if (!STATS && accumulate) {
                      const float4 o4 = old[STATS ? 0 : cc * 4 + k / 4];
                      v.x += o4.x; v.y += o4.y; v.z += o4.z; v.w += o4.w;
                    }
                    *reinterpret_cast<float4*>(orow + nn) = v;
                    if constexpr (STATS) {
                      constexpr int kMask = kNS - 1;
                      stats_add(st_s[(cc * 16 + k) & kMask], st_q[(cc * 16 + k) & kMask], v.x);
                      stats_add(st_s[(cc * 16 + k + 1) & kMask], st_q[(cc * 16 + k + 1) & kMask], v.y);
                      stats_add(st_s[(cc * 16 + k + 2) & kMask], st_q[(cc * 16 + k + 2) & kMask], v.z);
                      stats_add(st_s[(cc * 16 + k + 3) & kMask], st_q[(cc * 16 + k + 3) & kMask], v.w);
                    }
                  }
                }
              }
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(sl));
          if (p.prof) pe_drain += clock64() - te1;
        }
        par ^= ((1u << n) - 1u) << slot0;
      }
    }
    if (p.prof && threadIdx.x == 64) { p.prof[blockIdx.x * 8 + 6] = pe_wait; p.prof[blockIdx.x * 8 + 7] = pe_drain; }
    if constexpr (STATS) {
      double* sred = reinterpret_cast<double*>(smem_raw + (wbase - smem_u32(smem_raw)));
      stats_flush(st_s, st_q, p.N, q, lane, sred, p.stats);
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

// number of persistent CTAs: the SM count, or DPI_TC_MARCH_CTAS (test knob: forces several units per CTA on small
// problems so the cross-unit pipelining is exercised)
static int sm_count() {
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

// set by choose_parts() around its planning passes: reject plans that squeeze a wide input into 8-channel chunks
static thread_local bool g_plan_strict = false;
// two CTAs per SM: each gets half of the 228 KB (1 KB per CTA is reserved by the system)
constexpr int kSmemLimitTwo = 113 * 1024;
// DPI_TC_MARCH_2CTA: bit 0 = 3x3(x3) convs with C <= 16, bit 1 = packed march with one 32-channel chunk, bit 2 = 1x1,
// bit 3 = data gradients with a 32-column accumulator and C <= 8
static int two_ctas_mask() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DPI_TC_MARCH_2CTA");
    on = (e && e[0]) ? atoi(e) : 15;
  }
  return on;
}

// Tiling / shared-memory plan for a u-space of (Ud,Hu,Wu) output voxels.  nslab = 0: plain conv, all 27 (or 9)
// weight tiles resident; nslab > 0: parity-class mode with that many tiles per channel chunk.
static bool plan(int Ud, int Hu, int Wu, int C, int N, int nkd, int pd, int transposed, int nslab, Params& p,
                 size_t* smem_out, int halo = 1, int thin_c = 0, int want_tma = 0) {
  if ((C & 3) || (N & 3) || C < 4 || Ud < 1 || Hu < 1 || Wu < 1) return false;
  p.Do = Ud; p.Ho = Hu; p.Wo = Wu;
  p.C = C; p.N = N; p.nkd = nkd; p.pd = pd; p.transposed = transposed;
  p.tiles_w = (Wu + TW - 1) / TW;
  p.tiles_h = (Hu + TH - 1) / TH;
  p.BN = (N + 15) / 16 * 16;
  if (p.BN * 4 > 512) return false;
  p.slot_shift = p.BN * 16 <= 512 ? 4 : (p.BN * 8 <= 512 ? 3 : 2);
  // channel chunk = shared-memory row (32/64/128 B with the matching swizzle); widest one whose resident weights
  // leave room for >= 3 plane stages
  int kc_max = C <= 8 ? 8 : (C <= 16 ? 16 : 32);
  {
    // experiment knob: wider shared-memory rows than the channel count needs (the tail is TMA zero fill)
    static const int kc_floor = [] { const char* e = getenv("DPI_TC_MARCH_KC_MIN"); return e ? atoi(e) : 0; }();
    if (kc_floor > kc_max && (kc_floor == 16 || kc_floor == 32)) kc_max = kc_floor;
  }
  const int bar_bytes = 8 * (2 * kMaxStages + 2 + 2 * kSlots) + 16;
  // TMA-store epilogue: two staging tiles of 128 rows x N floats (+ alignment slack)
  const int64_t stage_bytes = want_tma ? 2LL * TH * TW * N * 4 + 256 : 0;
  // thin plain convs (one 16-column accumulator per plane): two CTAs per SM, half the shared memory each
  p.ctas_per_sm = 1;
  int64_t smem_limit = tc_smem_budget(false);
  if (p.BN == 16 && thin_c == 0 && !want_tma &&
      (((two_ctas_mask() & 1) && halo == 1 && nslab == 0 && C <= 16) || ((two_ctas_mask() & 4) && halo == 0 && nslab == 1))) {
    p.ctas_per_sm = 2;
    smem_limit = kSmemLimitTwo;
  }
  // ... and data gradients with a 32-column accumulator and a single K-step per tap (the 25 -> 1 conv: dy has one channel):
  // eight accumulator slots instead of sixteen, so that two CTAs' TMEM (2 x 256 columns) fits
  if ((two_ctas_mask() & 8) && p.BN == 32 && transposed && thin_c == 0 && !want_tma && halo == 1 && nslab == 0 && C <= 8) {
    p.ctas_per_sm = 2;
    smem_limit = kSmemLimitTwo;
    p.slot_shift = 3;
  }
  bool ok = false;
  for (int kc = kc_max; kc >= 8 && !ok; kc >>= 1) {
    p.kc = kc;
    p.rb = kc * 4;
    p.n_chunks = (C + kc - 1) / kc;
    p.plane_bytes = ((TH + 2 * halo) * (TW + 2 * halo) * p.rb + 1023) / 1024 * 1024;
    int64_t wbytes;
    if (thin_c > 0) {
      // thin slabs hold the first kthin K-steps (8 channels each) of every tap: rows of 32 / 64 / 128 bytes
      int kthin = (thin_c + 7) / 8;
      if (kthin == 3) kthin = 4;
      if (kthin > 4 || kthin * 8 > kc || nslab != 0 || halo != 1) return false;
      p.kthin = kthin;
      p.t_rb = kthin * 32;
      p.t_layout = kthin == 4 ? 2 : (kthin == 2 ? 4 : 6);
      p.t_slab_bytes = (9 * p.BN * p.t_rb + 1023) / 1024 * 1024;
      p.c_tile_bytes = (p.BN * p.rb + 1023) / 1024 * 1024;
      p.wslab_bytes = p.t_slab_bytes;
      wbytes = (int64_t)nkd * p.t_slab_bytes + (int64_t)p.n_chunks * p.c_tile_bytes;
    } else if (nslab == 0) {
      p.wslab_bytes = (9 * p.BN * p.rb + 1023) / 1024 * 1024;
      wbytes = (int64_t)p.n_chunks * nkd * p.wslab_bytes;
    } else {
      p.wslab_bytes = (p.BN * p.rb + 1023) / 1024 * 1024;
      wbytes = (int64_t)p.n_chunks * nslab * p.wslab_bytes;
    }
    const int64_t avail = smem_limit - 1024 - bar_bytes - wbytes - stage_bytes;
    if (avail < 3LL * p.plane_bytes) continue;
    if (g_plan_strict && kc == 8 && C > 16 && halo == 1 && nslab == 0) continue;
    int stages = (int)(avail / p.plane_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    p.stages = stages;
    p.wregion_bytes = (int)wbytes;
    ok = true;
  }
  if (!ok) {
    // no room for the staging tiles: the per-thread epilogue needs none
    return want_tma ? plan(Ud, Hu, Wu, C, N, nkd, pd, transposed, nslab, p, smem_out, halo, thin_c, 0) : false;
  }
  p.tma_store = want_tma;
  p.layout = p.kc == 32 ? 2 : (p.kc == 16 ? 4 : 6);
  // segments of output planes: minimise (rounds over the CTAs) x (planes a unit streams)
  const int nsm = sm_count() * p.ctas_per_sm;
  const int ncol = p.tiles_w * p.tiles_h;
  double best = 1e30;
  p.seg_len = Ud; p.n_segs = 1;
  for (int want = 1; want <= Ud; ++want) {
    const int len = (Ud + want - 1) / want;
    const int segs = (Ud + len - 1) / len;
    const int64_t units = (int64_t)ncol * segs;
    const int64_t rounds = (units + nsm - 1) / nsm;
    const double cost = (double)rounds * (len + nkd - 1 + 0.75);
    if (cost < best - 1e-9) { best = cost; p.seg_len = len; p.n_segs = segs; }
  }
  p.n_units = ncol * p.n_segs;
  int cols = 32;
  while (cols < (p.BN << p.slot_shift)) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.halo = halo;
  p.thin_c = thin_c;
  if (thin_c <= 0) { p.kthin = 0; p.t_rb = 32; p.t_layout = 6; p.t_slab_bytes = 0; p.c_tile_bytes = 0; }
  p.masked = 0; p.jmask = 7; p.tmask = 0x1ff; p.ntp = 9; p.nslab = nslab;
  p.osd = p.os = 1; p.ocd = p.och = p.ocw = 0;
  p.OD = Ud; p.OH = Hu; p.OW = Wu;
  *smem_out = (size_t)p.wregion_bytes + (size_t)p.stages * p.plane_bytes + bar_bytes + 1024 + (size_t)stage_bytes;
  return true;
}

// outputs narrower than this keep the per-thread epilogue (DPI_TC_TMA_STORE_MIN_N; 0 switches the TMA stores off)
static int tma_store_wanted(int N) {
  static int min_n = -1;
  if (min_n < 0) {
    const char* e = getenv("DPI_TC_TMA_STORE_MIN_N");
    min_n = (e && e[0]) ? atoi(e) : 40;
  }
  return (min_n > 0 && N >= min_n) ? 1 : 0;
}

// the output tensor [OD][OH][OW][out_ld] as a TMA store target: one box = a 16 x 8 tile of N channels
static int encode_out_map(EncodeTiledFn encode, float* out, int64_t out_ld, int N, int OW, int OH, int OD, CUtensorMap* mc) {
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)OD};
  cuuint64_t strides[3] = {(cuuint64_t)out_ld * 4, (cuuint64_t)OW * out_ld * 4, (cuuint64_t)OH * OW * out_ld * 4};
  cuuint32_t box[4] = {(cuuint32_t)N, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = encode(mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("march: cuTensorMapEncodeTiled(out) failed: %d", (int)r); return DPI_ERR_CUDA; }
  return DPI_OK;
}

static CUtensorMapSwizzle swizzle_for_row_bytes(int rb) {
  return rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// packed weights Wp[n][tap][c] viewed as (c, n, tap): one box = [box_taps][BN][box_c channels]
static int encode_weight_map(EncodeTiledFn encode, const float* Wp, const GatherGeom& g, int BN, int box_c, int box_taps,
                             CUtensorMap* mb) {
  const int taps = g.kd * g.kh * g.kw;
  const int wc = g.wC > 0 ? g.wC : g.C;         // channel pitch of the pack (C may be a sub-range of it)
  cuuint64_t dims[3] = {(cuuint64_t)g.C, (cuuint64_t)g.N, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)taps * wc * 4, (cuuint64_t)wc * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)BN, (cuuint32_t)box_taps};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode(mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(Wp), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(box_c * 4), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("march: cuTensorMapEncodeTiled(B) failed: %d", (int)r); return DPI_ERR_CUDA; }
  return DPI_OK;
}

static int encode_maps(EncodeTiledFn encode, const float* in, int64_t in_ld, const float* Wp, const GatherGeom& g,
                       const Params& p, int w_box_taps, CUtensorMap* ma, CUtensorMap* mb, CUtensorMap* mb2 = nullptr) {
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)in_ld * 4, (cuuint64_t)g.Wi * in_ld * 4, (cuuint64_t)g.Hi * g.Wi * in_ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)(TW + 2 * p.halo), (cuuint32_t)(TH + 2 * p.halo), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(p.kc * 4), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("march: cuTensorMapEncodeTiled(A) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  int rc;
  if (mb2 && p.thin_c > 0) {
    // thin layout: nine-tap slabs of the first kthin K-steps, and one-tap full-width tiles for the centre tap
    rc = encode_weight_map(encode, Wp, g, p.BN, 8 * p.kthin, 9, mb);
    if (!rc) rc = encode_weight_map(encode, Wp, g, p.BN, p.kc, 1, mb2);
    return rc;
  }
  rc = encode_weight_map(encode, Wp, g, p.BN, p.kc, w_box_taps, mb);
  if (!rc && mb2) *mb2 = *mb;
  return rc;
}

// A pending request for fused BatchNorm statistics (set by dpi_conv_fwd_stats around the dispatch): taken by the first
// eligible launch - forward, not accumulating, N <= kStatsMaxN, one launch covering the whole output.
static void* take_stats_request(int N, int transposed, int accumulate) {
  StatsRequest* rq = stats_request();
  if (!rq || !rq->ws || rq->done || transposed || accumulate || N > kStatsMaxN) return nullptr;
  const char* e = getenv("DPI_TC_FUSED_STATS");
  if (e && e[0] == '0') return nullptr;
  rq->done = true;
  return rq->ws;
}

template <bool STATS, int MAXBN>
static int launch_t(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const CUtensorMap& mc,
                    const float* bias, float* out, const Params& p, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_march_kernel<STATS, MAXBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("march: cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const int nsm = sm_count() * p.ctas_per_sm;
  const unsigned grid = (unsigned)(p.n_units < nsm ? p.n_units : nsm);
  conv_tc_march_kernel<STATS, MAXBN><<<grid, kThreads, smem, st>>>(ma, mb, mb2, mc, bias, out, p);
  return check_launch("conv_tc_march_kernel");
}

static int launch(EncodeTiledFn encode, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2,
                  const float* bias, float* out, Params& p, size_t smem, cudaStream_t st, bool may_emit_stats = false) {
  p.stats = may_emit_stats ? take_stats_request(p.N, p.transposed, p.accumulate) : nullptr;
  if (p.stats || p.os != 1) p.tma_store = 0;
  CUtensorMap mc = ma;                      // (placeholder when the TMA-store epilogue is off)
  if (p.tma_store) {
    const int rc = encode_out_map(encode, out, p.out_ld, p.N, p.OW, p.OH, p.OD, &mc);
    if (rc) return rc;
  }
  if (p.BN <= 16)
    return p.stats ? launch_t<true, 16>(ma, mb, mb2, mc, bias, out, p, smem, st)
                   : launch_t<false, 16>(ma, mb, mb2, mc, bias, out, p, smem, st);
  if (p.BN <= 32 && p.ctas_per_sm == 2 && !p.stats) return launch_t<false, 32>(ma, mb, mb2, mc, bias, out, p, smem, st);
  return p.stats ? launch_t<true, kMaxBN>(ma, mb, mb2, mc, bias, out, p, smem, st)
                 : launch_t<false, kMaxBN>(ma, mb, mb2, mc, bias, out, p, smem, st);
}

// plan of the packed variant; false when the shape is not eligible (then the plain march is used)
static bool plan_packed(const GatherGeom& g, PackedParams& p, size_t* smem_out) {
  const char* e = getenv("DPI_TC_MARCH_PACKED");
  if (e && e[0] == '0') return false;
  if (g.kd != 3 || (g.C & 3) || (g.N & 3) || g.C < 4) return false;
  p.BN = (g.N + 15) / 16 * 16;
  if (p.BN > 32) return false;                      // 2 x group slots x BN columns of TMEM, N = 3*BN <= 96
  // groups of eight planes: 10 plane loads per 8 outputs instead of 8 per 6; 16 slots x BN <= 512 columns (256 with BN = 16,
  // so two CTAs per SM still fit).  DPI_TC_MARCH_PACKED_GROUP=6 restores the smaller groups
  static const int group = [] { const char* e = getenv("DPI_TC_MARCH_PACKED_GROUP"); const int v = e ? atoi(e) : 8; return v == 6 ? 6 : 8; }();
  p.group = group;
  p.prof = nullptr;
  // With a single K-step per tap (C <= 8) a plane is only 27 MMAs; with one CTA per SM and groups of six planes the plain
  // march was faster there.  With two CTAs per SM (which hide the per-plane scalar path) and groups of eight the packed
  // form wins: 4 -> 8 forward 219 -> 182 us, 8 -> 13 forward 193 -> 180, iteration 25.29 -> 25.24 ms, 64^3 4.40 -> 4.35 ms.
  // DPI_TC_MARCH_PACKED_THIN=0 keeps C <= 8 on the plain march.
  {
    static const int thin_ok = [] { const char* e = getenv("DPI_TC_MARCH_PACKED_THIN"); return (e && e[0] == '0') ? 0 : 1; }();
    if ((g.C + 7) / 8 < 2 && !thin_ok) return false;
  }
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N; p.transposed = g.transposed;
  p.thin_c = 0;
  p.tiles_w = (g.Wo + TW - 1) / TW;
  p.tiles_h = (g.Ho + TH - 1) / TH;
  int kc_max = g.C <= 8 ? 8 : (g.C <= 16 ? 16 : 32);
  {
    // experiment knob (as in plan()): wider shared-memory rows than the channel count needs
    static const int kc_floor = [] { const char* e = getenv("DPI_TC_MARCH_KC_MIN"); return e ? atoi(e) : 0; }();
    if (kc_floor > kc_max && (kc_floor == 16 || kc_floor == 32)) kc_max = kc_floor;
  }
  const int bar_bytes = 8 * (2 * kMaxStages + 2 + 2 * kSlots) + 16;
  // thin layers (BN = 16, one narrow channel chunk): two CTAs per SM, as in plan()
  p.ctas_per_sm = 1;
  int64_t smem_limit = tc_smem_budget(false);
  if (p.BN == 16 && g.thin_c == 0 && (((two_ctas_mask() & 1) && g.C <= 16) || ((two_ctas_mask() & 2) && g.C <= 32))) {
    p.ctas_per_sm = 2;
    smem_limit = kSmemLimitTwo;
  }
  bool ok = false;
  for (int kc = kc_max; kc >= 8 && !ok; kc >>= 1) {
    p.kc = kc;
    p.rb = kc * 4;
    p.n_chunks = (g.C + kc - 1) / kc;
    p.plane_bytes = (HH * WW * p.rb + 1023) / 1024 * 1024;
    p.wchunk_bytes = (27 * p.BN * p.rb + 1023) / 1024 * 1024;
    const int64_t wbytes = (int64_t)p.n_chunks * p.wchunk_bytes;
    const int64_t avail = smem_limit - 1024 - bar_bytes - wbytes;
    // two stages are enough here: a wide-chunk stage is >= 36 MMAs of 48 clk, longer than a TMA round trip, while
    // narrower chunks would multiply the per-stage scalar path (C = 72: 3 chunks of 32 beat 5 chunks of 16)
    if (avail < 2LL * p.plane_bytes) continue;
    if (g_plan_strict && kc == 8 && g.C > 16) continue;
    int stages = (int)(avail / p.plane_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    p.stages = stages;
    ok = true;
  }
  if (!ok) return false;
  p.layout = p.kc == 32 ? 2 : (p.kc == 16 ? 4 : 6);
  // segments: minimise (rounds over the CTAs) x (planes a unit streams = L + 2 per group of six)
  const int nsm = sm_count() * p.ctas_per_sm;
  const int ncol = p.tiles_w * p.tiles_h;
  double best = 1e30;
  p.seg_len = g.Do; p.n_segs = 1;
  for (int want = 1; want <= g.Do; ++want) {
    const int len = (g.Do + want - 1) / want;
    const int segs = (g.Do + len - 1) / len;
    const int64_t units = (int64_t)ncol * segs;
    const int64_t rounds = (units + nsm - 1) / nsm;
    const double cost = (double)rounds * (len + 2 * ((len + p.group - 1) / p.group) + 0.75);
    if (cost < best - 1e-9) { best = cost; p.seg_len = len; p.n_segs = segs; }
  }
  p.n_units = ncol * p.n_segs;
  int cols = 32;
  while (cols < p.group * kBanks * p.BN) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  for (int i = 0; i < 3; ++i)
    p.idesc[i] = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(((i + 1) * p.BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  *smem_out = (size_t)p.n_chunks * p.wchunk_bytes + (size_t)p.stages * p.plane_bytes + bar_bytes + 1024;
  return true;
}

template <bool STATS, int MAXBN>
static int launch_packed_t(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias, float* out,
                           const PackedParams& p, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_march_packed_kernel<STATS, MAXBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("march(packed): cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const int nsm = sm_count() * p.ctas_per_sm;
  const unsigned grid = (unsigned)(p.n_units < nsm ? p.n_units : nsm);
  static const bool prof_on = [] { const char* e = getenv("DPI_TC_MARCH_PROF"); return e && e[0] == '1'; }();
  if (prof_on) {
    // debugging aid: cycle counters of the producer / MMA / epilogue roles, averaged over the CTAs (synchronises!)
    PackedParams q = p;
    cudaMalloc(&q.prof, (size_t)grid * 8 * sizeof(long long));
    cudaMemsetAsync(q.prof, 0, (size_t)grid * 8 * sizeof(long long), st);
    conv_tc_march_packed_kernel<STATS, MAXBN><<<grid, kThreads, smem, st>>>(ma, mb, bias, out, q);
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)grid * 8);
    cudaMemcpy(h.data(), q.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(q.prof);
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned i = 0; i < grid; ++i) for (int k = 0; k < 8; ++k) a[k] += (double)h[i * 8 + k] / grid;
    fprintf(stderr, "march(packed) prof C=%d N=%d grid=%u: steps/CTA %.0f, MMA warp total %.0f clk = %.0f per step: wait tempty %.0f, "
            "wait full %.0f, issue region %.0f; producer wait empty %.0f per step; epilogue wait tfull %.0f, drain %.0f per step\n",
            p.C, p.N, grid, a[4], a[5], a[5] / a[4], a[1] / a[4], a[2] / a[4], a[3] / a[4], a[0] / a[4], a[6] / a[4], a[7] / a[4]);
    return check_launch("conv_tc_march_packed_kernel");
  }
  conv_tc_march_packed_kernel<STATS, MAXBN><<<grid, kThreads, smem, st>>>(ma, mb, bias, out, p);
  return check_launch("conv_tc_march_packed_kernel");
}

static int launch_packed(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias, float* out,
                         PackedParams& p, size_t smem, cudaStream_t st, bool allow_stats = true) {
  p.stats = allow_stats ? take_stats_request(p.N, p.transposed, p.accumulate) : nullptr;
  if (p.BN <= 16)
    return p.stats ? launch_packed_t<true, 16>(ma, mb, bias, out, p, smem, st)
                   : launch_packed_t<false, 16>(ma, mb, bias, out, p, smem, st);
  return p.stats ? launch_packed_t<true, 32>(ma, mb, bias, out, p, smem, st)
                 : launch_packed_t<false, 32>(ma, mb, bias, out, p, smem, st);
}

static bool enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DPI_TC_MARCH");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

static int debug_bits() {
  const char* e = getenv("DPI_TC_MARCH_DEBUG");
  return e ? atoi(e) : 0;
}

}  // namespace march

// An output wider than one accumulator (N > 128 columns x 4 slots = all of TMEM) is computed as 2..4 launches over
// output-channel ranges: each re-reads the (narrow) input and owns its columns of the output (weight rows
// [n0, n0 + Ns) of Wp[n][tap][c], bias + n0, out + n0).  Before, such shapes fell to the tile-per-CTA kernels: 325 us
// for the 8 -> 144 channel data gradient at half resolution against 49 us of HBM time.
static int part_width(int N, int parts) { return ((N + parts - 1) / parts + 3) / 4 * 4; }
constexpr int kMaxParts = 4;

static int march_gather_one(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                            const GatherGeom& g, int accumulate, cudaStream_t st, bool dry_run, bool allow_stats);

// smallest number of output-channel ranges for which every range is plannable (0: none up to kMaxParts)
template <class PlanOne>
static int choose_parts(const GatherGeom& g, PlanOne plan_one) {
  // Outputs that no single launch can hold (N > 128) are split into up to four ranges; narrower ones whose weights do
  // not fit next to the plane stages into TWO at most (the half-resolution 51 -> 32 conv and its data gradient: 325 /
  // 333 us on the tile-per-CTA kernel against 2 x ~60 us here).  Splitting further was measured too: it helps some
  // layers (276 -> 17 dgrad 247 -> 157 us) and hurts others (105 -> 64 forward 99 -> 240 us, every range re-reads the
  // 112-channel input), and costs small patches a launch each (64^3: +8 %).
  static const int narrow_parts = [] { const char* e = getenv("DPI_TC_MARCH_SPLIT_NARROW"); return (e && e[0] == '0') ? 1 : 2; }();
  // (small problems are launch-bound: a 64^3 patch got 2 % slower with the extra launches, so narrow outputs are only
  //  split from 128 K output voxels up)
  const bool big = (int64_t)g.Do * g.Ho * g.Wo >= 131072;
  const int max_parts = g.N > 128 ? kMaxParts : (big ? narrow_parts : 1);
  // first pass: only plans with shared-memory rows of >= 64 bytes for C > 16 (march::g_plan_strict) - a plan that fits
  // the weights only with 32-byte rows (51 -> 32 forward: seven 8-channel chunks per plane, 322 us) loses against two
  // ranges with 128-byte rows; second pass: anything that fits
  for (int strict = (narrow_parts > 1 && big) ? 1 : 0; strict >= 0; --strict) {
    march::g_plan_strict = strict != 0;
    for (int parts = (g.N + 127) / 128; parts <= max_parts; ++parts) {
      const int Ns = part_width(g.N, parts);
      if (Ns < 4 || (parts > 1 && Ns < 8)) break;
      bool ok = true;
      for (int n0 = 0; n0 < g.N && ok; n0 += Ns) {
        GatherGeom gp = g;
        gp.N = g.N - n0 < Ns ? g.N - n0 : Ns;
        ok = plan_one(gp);
      }
      if (ok) return parts;           // (g_plan_strict stays as it was for the launches that follow)
    }
  }
  march::g_plan_strict = false;
  return 0;
}

static int march_gather_parts(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                              int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st);

int conv_tc_march_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                         const GatherGeom& g, int accumulate, cudaStream_t st) {
  const int rc = march_gather_parts(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
  march::g_plan_strict = false;       // (choose_parts leaves the mode of its successful pass set for the launches)
  return rc;
}

static int march_gather_parts(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out,
                              int64_t out_ld, const GatherGeom& g, int accumulate, cudaStream_t st) {
  const int parts = choose_parts(g, [&](const GatherGeom& gp) {
    return march_gather_one(in, in_ld, Wp, bias, out, out_ld, gp, accumulate, st, true, false) == DPI_OK;
  });
  if (parts == 0) {
    // Wide input, narrow output whose resident weights do not fit (137 -> 8 at half resolution: 380 us on the
    // tile-per-CTA kernel): the REDUCTION is split in two channel ranges - two launches over the same output, the second
    // accumulating.  (The BatchNorm statistics of such a forward conv are left to the separate pass.)
    static const int ksplit = [] { const char* e = getenv("DPI_TC_MARCH_SPLIT_K"); return (e && e[0] == '0') ? 0 : 1; }();
    if (!ksplit || g.wC > 0 || g.thin_c > 0 || g.N > 32 || g.C < 96 || (int64_t)g.Do * g.Ho * g.Wo < 131072) return DPI_ERR_UNSUPPORTED;
    const int c0 = (g.C / 2 + 31) / 32 * 32;
    GatherGeom ga = g, gb = g;
    ga.C = c0; ga.wC = g.C;
    gb.C = g.C - c0; gb.wC = g.C;
    if (gb.C < 4) return DPI_ERR_UNSUPPORTED;
    march::g_plan_strict = false;
    if (march_gather_one(in, in_ld, Wp, bias, out, out_ld, ga, accumulate, st, true, false) != DPI_OK ||
        march_gather_one(in + c0, in_ld, Wp + c0, nullptr, out, out_ld, gb, 1, st, true, false) != DPI_OK)
      return DPI_ERR_UNSUPPORTED;
    int rc = march_gather_one(in, in_ld, Wp, bias, out, out_ld, ga, accumulate, st, false, false);
    if (rc) return rc;
    return march_gather_one(in + c0, in_ld, Wp + c0, nullptr, out, out_ld, gb, 1, st, false, false);
  }
  if (parts == 1) return march_gather_one(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st, false, true);
  const int taps = g.kd * g.kh * g.kw, Ns = part_width(g.N, parts);
  for (int n0 = 0; n0 < g.N; n0 += Ns) {
    GatherGeom gp = g;
    gp.N = g.N - n0 < Ns ? g.N - n0 : Ns;
    // (a split forward conv leaves the BatchNorm statistics to the separate pass: one partial row per CTA covers all
    //  channels of a launch, not a column range)
    const int rc = march_gather_one(in, in_ld, Wp + (int64_t)n0 * taps * g.C, bias ? bias + n0 : nullptr, out + n0, out_ld, gp,
                                    accumulate, st, false, false);
    if (rc) return rc;
  }
  return DPI_OK;
}

static int march_gather_one(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                            const GatherGeom& g, int accumulate, cudaStream_t st, bool dry_run, bool allow_stats) {
  using namespace march;
  if (!enabled()) return DPI_ERR_UNSUPPORTED;
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.kh != 3 || g.kw != 3 || (g.kd != 3 && g.kd != 1)) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  {
    // narrow outputs (Cout <= 32, 3-D): the three kd taps share one MMA
    PackedParams pp;
    size_t psmem = 0;
    if (plan_packed(g, pp, &psmem)) {
      if (dry_run) return DPI_OK;
      pp.out_ld = out_ld;
      pp.accumulate = accumulate;
      pp.thin_c = g.thin_c;
      Params shape;                      // only kc / BN are read by encode_maps
      shape.kc = pp.kc; shape.BN = pp.BN; shape.halo = 1;
      CUtensorMap ma, mb;
      const int rc = encode_maps(encode, in, in_ld, Wp, g, shape, 1, &ma, &mb);
      if (rc) return rc;
      return launch_packed(ma, mb, bias, out, pp, psmem, st, allow_stats);
    }
  }
  Params p;
  size_t smem = 0;
  if (!plan(g.Do, g.Ho, g.Wo, g.C, g.N, g.kd, g.pd, g.transposed, 0, p, &smem, 1, g.thin_c, tma_store_wanted(g.N)))
    return DPI_ERR_UNSUPPORTED;
  if (dry_run) return DPI_OK;
  p.out_ld = out_ld;
  p.accumulate = accumulate;
  p.debug = debug_bits();
  CUtensorMap ma, mb, mb2;
  const int rc = encode_maps(encode, in, in_ld, Wp, g, p, 9, &ma, &mb, &mb2);
  if (rc) return rc;
  return launch(encode, ma, mb, mb2, bias, out, p, smem, st, allow_stats);
}

int conv_tc_march_supported(const GatherGeom& g) {
  using namespace march;
  if (!enabled() || !get_encode()) return 0;
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.kh != 3 || g.kw != 3 || (g.kd != 3 && g.kd != 1)) return 0;
  const int parts = choose_parts(g, [&](const GatherGeom& gp) {
    PackedParams pp;
    size_t smem = 0;
    if (plan_packed(gp, pp, &smem)) return true;
    Params p;
    return plan(gp.Do, gp.Ho, gp.Wo, gp.C, gp.N, gp.kd, gp.pd, gp.transposed, 0, p, &smem, 1, gp.thin_c);
  });
  g_plan_strict = false;
  // (a narrow output that only fits as two ranges is not offered for the fused data gradient: measured on the
  //  half-resolution 51 -> 32 + 1x1 pair, fused 298 us against 146 + 55 us for the two separate launches)
  if (parts > 1 && g.N <= 128 && g.thin_c > 0) return 0;
  return parts > 0 ? 1 : 0;
}

// 1x1(x1) convolutions (the shortcut / ResPath convs, mulresunet.py:82,105), forward and dgrad: HBM-bound, so what
// matters is a pipeline that never drains - the same persistent march with bare 16 x 8 tiles instead of halo planes,
// one resident weight tile per channel chunk and one tap.
static int march_1x1_one(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                         const GatherGeom& g, int accumulate, cudaStream_t st, bool dry_run);

int conv_tc_march_1x1(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                      const GatherGeom& g, int accumulate, cudaStream_t st) {
  const int parts = choose_parts(g, [&](const GatherGeom& gp) {
    return march_1x1_one(in, in_ld, Wp, bias, out, out_ld, gp, accumulate, st, true) == DPI_OK;
  });
  march::g_plan_strict = false;       // (the strict mode only concerns 3x3 plans)
  if (parts == 0) return DPI_ERR_UNSUPPORTED;
  if (parts == 1) return march_1x1_one(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st, false);
  const int Ns = part_width(g.N, parts);
  for (int n0 = 0; n0 < g.N; n0 += Ns) {
    GatherGeom gp = g;
    gp.N = g.N - n0 < Ns ? g.N - n0 : Ns;
    const int rc = march_1x1_one(in, in_ld, Wp + (int64_t)n0 * g.C, bias ? bias + n0 : nullptr, out + n0, out_ld, gp, accumulate,
                                 st, false);
    if (rc) return rc;
  }
  return DPI_OK;
}

static int march_1x1_one(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                         const GatherGeom& g, int accumulate, cudaStream_t st, bool dry_run) {
  using namespace march;
  if (!enabled()) return DPI_ERR_UNSUPPORTED;
  {
    const char* e = getenv("DPI_TC_MARCH_1X1");
    if (e && e[0] == '0') return DPI_ERR_UNSUPPORTED;
  }
  if (g.kd != 1 || g.kh != 1 || g.kw != 1 || g.sd != 1 || g.sh != 1 || g.sw != 1) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  Params p;
  size_t smem = 0;
  if (!plan(g.Do, g.Ho, g.Wo, g.C, g.N, 1, 0, g.transposed, 1, p, &smem, 0, 0, tma_store_wanted(g.N))) return DPI_ERR_UNSUPPORTED;
  if (dry_run) return DPI_OK;
  p.masked = 1;
  p.jmask = 1; p.tmask = 1 << 4; p.ntp = 1;
  for (int j = 0; j < 3; ++j) p.jrank[j] = 0;
  for (int t = 0; t < 9; ++t) p.tprank[t] = 0;
  p.orig_tap[0] = 0;
  p.out_ld = out_ld;
  p.accumulate = accumulate;
  p.debug = 0;
  CUtensorMap ma, mb;
  const int rc = encode_maps(encode, in, in_ld, Wp, g, p, 1, &ma, &mb);
  if (rc) return rc;
  return launch(encode, ma, mb, mb, bias, out, p, smem, st, true);
}

// Data gradient of a stride-2 3x3(x3) convolution (the down-sampling convs, mulresunet.py:224-227) as one march per
// output parity class: dx[2u + c] = sum over the taps t with (c + 1 - t) even of W[t] . dy[u + (c + 1 - t)/2].
// Per axis class 0 has one tap (t = 1, offset 0) and class 1 two (t = 0, offset +1; t = 2, offset 0): in the
// u-space of a class this is the transposed gather above restricted to the taps k' = (t - c + 1)/2 in {0, 1}, with
// only those weight tiles resident and the outputs written with stride 2.  27 tap applications per 8 outputs and no
// zero-stuffed copy of dy.
int conv_tc_march_dgrad_s2(const float* dy, int64_t dy_ld, const float* Wt, float* dx, int64_t dx_ld,
                           const GatherGeom& g, int accumulate, cudaStream_t st) {
  using namespace march;
  if (!enabled()) return DPI_ERR_UNSUPPORTED;
  if (!g.transposed || g.sh != 2 || g.sw != 2 || g.kh != 3 || g.kw != 3) return DPI_ERR_UNSUPPORTED;
  if (!((g.kd == 3 && g.sd == 2) || (g.kd == 1 && g.sd == 1))) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  const int ncd = g.sd;                       // parity classes along d
  // all classes must be plannable before anything is launched (otherwise the caller falls back for the whole op)
  Params ps[8];
  size_t smems[8];
  int ncls = 0;
  for (int cd = 0; cd < ncd; ++cd)
    for (int ch = 0; ch < 2; ++ch)
      for (int cw = 0; cw < 2; ++cw) {
        const int Ud = g.sd == 2 ? (g.Do - cd + 1) / 2 : g.Do, Hu = (g.Ho - ch + 1) / 2, Wu = (g.Wo - cw + 1) / 2;
        if (Ud < 1 || Hu < 1 || Wu < 1) continue;     // no output voxel of this parity
        // participating u-space taps per axis: class 0 -> k' = 1; class 1 -> k' = 0, 1.  Relation j = nkd-1-kd'.
        int kds[2], nkdv = 0, khs[2], nkh = 0, kws[2], nkw = 0;
        if (g.kd == 1) { kds[nkdv++] = 0; }
        else if (cd == 0) { kds[nkdv++] = 1; }
        else { kds[nkdv++] = 0; kds[nkdv++] = 1; }
        if (ch == 0) { khs[nkh++] = 1; } else { khs[nkh++] = 0; khs[nkh++] = 1; }
        if (cw == 0) { kws[nkw++] = 1; } else { kws[nkw++] = 0; kws[nkw++] = 1; }
        const int ntp = nkh * nkw, nslab = nkdv * ntp;
        Params& p = ps[ncls];
        if (!plan(Ud, Hu, Wu, g.C, g.N, g.kd, g.pd, 1, nslab, p, &smems[ncls])) return DPI_ERR_UNSUPPORTED;
        p.masked = 1;
        p.jmask = 0; p.tmask = 0; p.ntp = ntp;
        for (int j = 0; j < 3; ++j) p.jrank[j] = 0;
        for (int t = 0; t < 9; ++t) p.tprank[t] = 0;
        // slabs ordered by relation j ascending, then in-plane tap ascending
        int r = 0, jr = 0;
        for (int j = 0; j < g.kd; ++j) {
          const int kdp = g.kd - 1 - j;
          bool has = false;
          for (int i = 0; i < nkdv; ++i) has |= kds[i] == kdp;
          if (!has) continue;
          p.jmask |= 1 << j;
          p.jrank[j] = jr++;
          const int td = g.kd == 1 ? 0 : cd - 1 + 2 * kdp;
          int tr = 0;
          for (int tp = 0; tp < 9; ++tp) {
            const int khp = tp / 3, kwp = tp % 3;
            bool hh = false, hw = false;
            for (int i = 0; i < nkh; ++i) hh |= khs[i] == khp;
            for (int i = 0; i < nkw; ++i) hw |= kws[i] == kwp;
            if (!hh || !hw) continue;
            p.tmask |= 1 << tp;
            p.tprank[tp] = tr++;
            const int th = ch - 1 + 2 * khp, tw = cw - 1 + 2 * kwp;
            p.orig_tap[r++] = (td * 3 + th) * 3 + tw;
          }
        }
        p.osd = g.sd; p.os = 2; p.ocd = cd; p.och = ch; p.ocw = cw;
        p.OD = g.Do; p.OH = g.Ho; p.OW = g.Wo;
        p.out_ld = dx_ld;
        p.accumulate = accumulate;
        p.debug = 0;
        ++ncls;
      }
  for (int i = 0; i < ncls; ++i) {
    CUtensorMap ma, mb;
    int rc = encode_maps(encode, dy, dy_ld, Wt, g, ps[i], 1, &ma, &mb);
    if (rc) return rc;
    rc = launch(encode, ma, mb, mb, nullptr, dx, ps[i], smems[i], st);
    if (rc) return rc;
  }
  return DPI_OK;
}

}  // namespace dpi
