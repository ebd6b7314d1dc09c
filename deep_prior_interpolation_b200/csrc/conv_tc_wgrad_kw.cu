// tcgen05 weight gradient with the three kw taps PACKED INTO THE M DIMENSION (stride-1 convs with kw = 3).
//
// conv_tc_wgrad.cu spends one MMA per (tap, 8 voxels); the single issuing thread (>= 44 clk per kind::tf32 MMA)
// is the bottleneck for the full-resolution layers.  Here one MMA covers three taps:
//   * per (kd, kh) and 32-channel block, ONE TMA box brings the x rows of the tile with a 1-voxel halo along w:
//     16 groups (a (d,h) pair each) x 10 rows x 32 channels;
//   * A (MN-major, 32-byte-atom 128-byte swizzle) is described with LBO = 128 B = ONE ROW: "channel block" j of
//     the M = 128 lanes is the same 32 channels read one voxel further along w, i.e. tap kw = j.  Lanes
//     [32*kw, 32*kw+32) therefore accumulate dW[:, (kd,kh,kw), c-block]; lanes 96..127 (kw = 3) are ignored.
//     K step g (8 voxels) starts at row 10*g of the box (verified: MN-major descriptors honour arbitrary row
//     offsets and LBO, scratch/umma_probe.cu cfg0);
//   * B = the dy tile (16 groups x 8 rows), exactly as in conv_tc_wgrad.cu.
// MMAs per 128-voxel tile: 9 * CB * 16 instead of 27 * 16 (CB = 32-channel blocks of the input, <= 3), and the
// L2 -> SMEM traffic for x drops from 27 x 128 to 9 x 160 rows per block.
// Accumulators (one per (kd,kh) pair and channel block, BN columns each) stay in TMEM across all voxel tiles of
// the CTA's chunk; split-K partials go to workspace[chunk][n][tap][c] and are reduced in a fixed order.
#include <cuda.h>
#include "conv_geom.cuh"

namespace dpi {
namespace wgk {

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
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// Same MMA with the descriptors passed as (lo, hi) 32-bit halves: only the 14-bit start-address field (lo word)
// changes between K steps, so advancing a descriptor is ONE 32-bit add off the critical path instead of a
// dependent 64-bit add-with-carry; the single issuing thread is latency-bound (ncu: ~20 clk per dependent
// instruction), and every instruction removed per MMA shows up directly in the MMA issue rate.
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
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

constexpr int TW = 8, WW = TW + 2;            // tile width and haloed width
constexpr int kXBlk = 16 * WW * 128;          // 16 groups x 10 rows x 128 B = 20480
constexpr int kDyBlk = 128 * 128;             // 128 voxels x 128 B
constexpr int kThreads = 192;

struct Params {
  int Do, Ho, Wo;
  int tiles_w, tiles_h, tiles_d, n_vtiles;
  int BD, BH;                     // BD * BH = 16 groups of 8 voxels along w
  int C, N, taps, kd, kh, pd, ph;
  int npairs;                     // kd * kh
  int CB, NB, BN;                 // channel blocks per c tile (<= 3), n blocks (<= 2), MMA N
  int PG;                         // (kd,kh) pairs per CTA
  int c_tiles, n_tiles, pair_groups, nchunks, vtiles_per_chunk;
  int stages;
  uint32_t idesc, tmem_cols;
};

__global__ void __launch_bounds__(kThreads)
conv_tc_wgrad_kw_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_dy,
                        float* __restrict__ partial, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_stage_bytes = (uint32_t)p.CB * kXBlk, dy_buf_bytes = (uint32_t)p.NB * kDyBlk;
  const uint32_t dy_base = base + (uint32_t)p.stages * x_stage_bytes;
  const uint32_t bar_base = dy_base + 2u * dy_buf_bytes + 2048u;   // slack: the kw = 3 pseudo-tap reads 1 row past a box
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto dy_full = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto dy_empty = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 + b); };
  const uint32_t done_bar = bar_base + 8u * (2 * p.stages + 4);
  const uint32_t tmem_slot = done_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int pg = b % p.pair_groups; b /= p.pair_groups;
  const int nt = b % p.n_tiles; b /= p.n_tiles;
  const int ct = b % p.c_tiles;
  const int chunk = b / p.c_tiles;
  const int pair0 = pg * p.PG;
  const int npair = min(p.PG, p.npairs - pair0);
  const int c_base = ct * p.CB * 32, n_base = nt * p.NB * 32;
  const int cb_here = min(p.CB, (p.C - c_base + 31) / 32);       // channel blocks that exist in this c tile
  const int vt_begin = chunk * p.vtiles_per_chunk;
  const int vt_end = min(p.n_vtiles, vt_begin + p.vtiles_per_chunk);
  const int nvt = vt_end - vt_begin;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(dy_full(i), 1); mbar_init(dy_empty(i), 1); }
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
    if (lane == 0 && nvt > 0) {
      // ================= TMA producer =================
      int it = 0;
      for (int v = 0; v < nvt; ++v) {
        int t = vt_begin + v;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int td = t / p.tiles_h;
        const int w0 = tw * TW, h0 = th * p.BH, d0 = td * p.BD;
        const int db = v & 1;
        mbar_wait(dy_empty(db), ((uint32_t)(v >> 1) & 1u) ^ 1u);
        mbar_expect_tx(dy_full(db), dy_buf_bytes);
        for (int j = 0; j < p.NB; ++j)
          tma_load_4d(dy_base + db * dy_buf_bytes + j * kDyBlk, &tma_dy, dy_full(db), n_base + 32 * j, w0, h0, d0);
        for (int pp = 0; pp < npair; ++pp, ++it) {
          const int pair = pair0 + pp;
          const int tkh = pair % p.kh, tkd = pair / p.kh;
          const int s = it % p.stages;
          mbar_wait(empty_bar(s), ((uint32_t)(it / p.stages) & 1u) ^ 1u);
          mbar_expect_tx(full_bar(s), (uint32_t)cb_here * kXBlk);
          for (int j = 0; j < cb_here; ++j)
            tma_load_4d(base + s * x_stage_bytes + j * kXBlk, &tma_x, full_bar(s), c_base + 32 * j, w0 - 1,
                        h0 + tkh - p.ph, d0 + tkd - p.pd);
        }
      }
    }
  } else if (warp == 1) {
    if (nvt > 0) {
      // ================= MMA issuer: whole-warp control flow, one elected lane issues (see conv_tc_march.cu) =====
      int it = 0;
      for (int v = 0; v < nvt; ++v) {
        const int db = v & 1;
        mbar_wait(dy_full(db), (uint32_t)(v >> 1) & 1u);
        tc_fence_after();
        const uint64_t bd0 = make_mn_desc(dy_base + db * dy_buf_bytes, kDyBlk, 512);
        const uint32_t blo = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32);
        const uint32_t acc0 = v > 0 ? 1u : 0u;
        for (int pp = 0; pp < npair; ++pp, ++it) {
          const int s = it % p.stages;
          mbar_wait(full_bar(s), (uint32_t)(it / p.stages) & 1u);
          tc_fence_after();
          if (elect_one()) {
            for (int j = 0; j < cb_here; ++j) {
              // LBO = 128 B: M block kw = the same channels one row (voxel) further along w
              const uint64_t ad0 = make_mn_desc(base + s * x_stage_bytes + j * kXBlk, 128, 512);
              const uint32_t dcol = tmem_d + (uint32_t)((pp * p.CB + j) * p.BN);
              const uint32_t alo = (uint32_t)ad0, ahi = (uint32_t)(ad0 >> 32);
              umma_tf32_lh(dcol, alo, ahi, blo, bhi, p.idesc, acc0);
#pragma unroll
              for (int g = 1; g < 16; ++g)          // K step g: x rows 10g.., dy rows 8g..
                umma_tf32_lh(dcol, alo + (uint32_t)(g * WW * 8), ahi, blo + (uint32_t)(g * 64), bhi, p.idesc, 1u);
            }
            umma_commit(empty_bar(s));
            if (pp == npair - 1) {
              umma_commit(dy_empty(db));
              if (v == nvt - 1) umma_commit(done_bar);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (nvt > 0) {
    // ================= epilogue: lanes [32*kw, 32*kw+32) of accumulator (pair, block) -> dW[:, (kd,kh,kw), c] ===
    const int kw = warp & 3;                 // TMEM lane quarter == kw tap
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = partial + (int64_t)chunk * p.N * p.taps * p.C;
    for (int pp = 0; pp < npair; ++pp) {
      const int pair = pair0 + pp;
      for (int j = 0; j < p.CB; ++j) {
        const int c = c_base + j * 32 + lane;
        for (int nn = 0; nn < p.BN; nn += 16) {
          float v[16];
          tmem_ld16(tmem_d + ((uint32_t)(kw * 32) << 16) + (uint32_t)((pp * p.CB + j) * p.BN + nn), v);
          if (kw < 3 && j < cb_here && c < p.C) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = n_base + nn + i;
              if (n < p.N) dst[((int64_t)n * p.taps + pair * 3 + kw) * p.C + c] = v[i];
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

static bool plan(const GatherGeom& g, Params& p) {
  if (g.transposed || g.sd != 1 || g.sh != 1 || g.sw != 1 || g.kw != 3) return false;
  if ((g.C & 3) || (g.N & 3)) return false;
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N; p.kd = g.kd; p.kh = g.kh; p.pd = g.pd; p.ph = g.ph;
  p.taps = g.kd * g.kh * 3;
  p.npairs = g.kd * g.kh;
  int bh = 1;
  while (bh < g.Ho && bh < 16) bh <<= 1;
  p.BH = bh;
  p.BD = 16 / bh;
  p.tiles_w = (g.Wo + TW - 1) / TW;
  p.tiles_h = (g.Ho + p.BH - 1) / p.BH;
  p.tiles_d = (g.Do + p.BD - 1) / p.BD;
  p.n_vtiles = p.tiles_w * p.tiles_h * p.tiles_d;
  const int cblocks = (g.C + 31) / 32, nblocks = (g.N + 31) / 32;
  p.NB = nblocks < 2 ? nblocks : 2;
  p.n_tiles = (nblocks + p.NB - 1) / p.NB;
  const int n_in_tile = g.N < p.NB * 32 ? g.N : p.NB * 32;
  p.BN = (n_in_tile + 15) / 16 * 16;
  p.CB = cblocks < 3 ? cblocks : 3;
  while (p.CB > 1 && p.CB * p.BN > 512) --p.CB;
  p.c_tiles = (cblocks + p.CB - 1) / p.CB;
  p.PG = 512 / (p.CB * p.BN);
  if (p.PG < 1) return false;
  if (p.PG > p.npairs) p.PG = p.npairs;
  p.pair_groups = (p.npairs + p.PG - 1) / p.PG;
  int cols = 32;
  while (cols < p.PG * p.CB * p.BN) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  const int base_ctas = p.c_tiles * p.n_tiles * p.pair_groups;
  int want = (148 + base_ctas - 1) / base_ctas;
  if (want > p.n_vtiles) want = p.n_vtiles;
  if (want < 1) want = 1;
  p.vtiles_per_chunk = (p.n_vtiles + want - 1) / want;
  p.nchunks = (p.n_vtiles + p.vtiles_per_chunk - 1) / p.vtiles_per_chunk;
  const int x_stage = p.CB * kXBlk, dy_buf = p.NB * kDyBlk;
  p.stages = (196 * 1024 - 2 * dy_buf) / x_stage;
  if (p.stages > 5) p.stages = 5;
  if (p.stages < 2) return false;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  return true;
}

}  // namespace wgk

int64_t conv_tc_wgrad_kw_workspace_bytes(const GatherGeom& g) {
  wgk::Params p;
  if (!wgk::plan(g, p)) return 0;
  return (int64_t)p.nchunks * g.N * p.taps * g.C * (int64_t)sizeof(float);
}

int conv_tc_wgrad_kw(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial, int64_t partial_bytes,
                     const GatherGeom& g, int* nchunks_out, cudaStream_t st) {
  using namespace wgk;
  Params p;
  if (!plan(g, p)) return DPI_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return DPI_ERR_UNSUPPORTED;
  const int64_t need = (int64_t)p.nchunks * g.N * p.taps * g.C * (int64_t)sizeof(float);
  if (partial_bytes < need) {
    set_error("conv_tc_wgrad_kw: workspace too small (%lld < %lld)", (long long)partial_bytes, (long long)need);
    return DPI_ERR_WORKSPACE;
  }
  CUtensorMap mx, mdy;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)g.Wi * x_ld * 4, (cuuint64_t)g.Hi * g.Wi * x_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)WW, (cuuint32_t)p.BH, (cuuint32_t)p.BD};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_kw: cuTensorMapEncodeTiled(x) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.N, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)g.Do};
    cuuint64_t strides[3] = {(cuuint64_t)dy_ld * 4, (cuuint64_t)g.Wo * dy_ld * 4, (cuuint64_t)g.Ho * g.Wo * dy_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)p.BH, (cuuint32_t)p.BD};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_kw: cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.CB * kXBlk + 2 * (size_t)p.NB * kDyBlk + 2048 + 8 * (2 * p.stages + 8) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_kw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("wgrad_kw: cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const unsigned grid = (unsigned)(p.nchunks * p.c_tiles * p.n_tiles * p.pair_groups);
  conv_tc_wgrad_kw_kernel<<<grid, kThreads, smem, st>>>(mx, mdy, partial, p);
  *nchunks_out = p.nchunks;
  return check_launch("conv_tc_wgrad_kw_kernel");
}

}  // namespace dpi
