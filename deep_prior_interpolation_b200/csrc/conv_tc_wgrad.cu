// tcgen05 / TMEM / TMA weight-gradient of the stride-1 convolutions (DPI_PREC_TF32).
//
//     dW[n][tap][c] = sum_v dy[v][n] * x[v + off(tap)][c]
//
// GEMM view per CTA: the reduction dimension K is the VOXEL index, so both operands are "MN-major" for the
// tensor core (channels are contiguous in the channels-last tensors, voxels are the strided dimension):
//     A = x tile, shifted by the tap   (M = input channels, up to 128 = 4 blocks of 32 at LBO)
//     B = dy tile                      (N = output channels, blocks of 32 at LBO)
//     D[c][n] (+)= sum over the 128 voxels of the tile, 8 voxels per tcgen05.mma (kind::tf32)
// For TF32 the only MN-major shared-memory layout the tensor core accepts is the 128-byte swizzle with 32-byte
// atoms (UMMA layout type SWIZZLE_128B_BASE32B, TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 4-row K groups at
// SBO = 512 B (verified on B200 with scratch/umma_probe.cu).
//
// Every tap of a tap-group owns its own accumulator (BN TMEM columns); the accumulators stay in TMEM while the
// CTA walks over all voxel tiles of its chunk (split-K over voxel chunks), and are written once at the end to
// workspace[chunk][n][tap][c]; wgrad_reduce_kernel (conv_simt.cu) sums the chunks in a fixed order, so the
// result is bit-reproducible.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
#include <cuda.h>
#include <stdlib.h>
#include "conv_geom.cuh"

namespace dpi {
namespace wg {

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
    if (spin > (1u << 27)) __trap();     // a protocol bug must not hang the GPU
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
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
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
// MN-major descriptor, layout SWIZZLE_128B_BASE32B (=1): LBO = byte distance between 32-float MN blocks,
// SBO = byte distance between 4-row K groups (512 for dense 128-byte rows)
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

struct WgParams {
  int Do, Ho, Wo;                 // output (dy) spatial dims
  int tiles_w, tiles_h, tiles_d, n_vtiles;
  int BD, BH, BW;
  int C, N, taps, kd, kh, kw, pd, ph, pw, sd, sh, sw;
  int CB, NB;                     // 32-channel blocks of the c tile / n tile
  int BN;                         // n tile width used by the MMA (multiple of 16)
  int TG;                         // taps per CTA (accumulators resident in TMEM)
  int c_tiles, n_tiles, tap_groups, nchunks, vtiles_per_chunk;
  int stages;
  uint32_t idesc, tmem_cols;
};

constexpr int kWgThreads = 192;
constexpr int kBlkBytes = 128 * 128;    // one 32-channel block of a 128-voxel tile

__global__ void __launch_bounds__(kWgThreads)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_dy,
                     float* __restrict__ partial, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_stage_bytes = (uint32_t)p.CB * kBlkBytes, dy_buf_bytes = (uint32_t)p.NB * kBlkBytes;
  const uint32_t dy_base = base + (uint32_t)p.stages * x_stage_bytes;
  const uint32_t bar_base = dy_base + 2u * dy_buf_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto dy_full = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto dy_empty = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 + b); };
  const uint32_t done_bar = bar_base + 8u * (2 * p.stages + 4);
  const uint32_t tmem_slot = done_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int tg = b % p.tap_groups; b /= p.tap_groups;
  const int nt = b % p.n_tiles; b /= p.n_tiles;
  const int ct = b % p.c_tiles;
  const int chunk = b / p.c_tiles;
  const int tap0 = tg * p.TG;
  const int ntap = min(p.TG, p.taps - tap0);
  const int c_base = ct * p.CB * 32, n_base = nt * p.NB * 32;
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
        const int w0 = tw * p.BW, h0 = th * p.BH, d0 = td * p.BD;
        const int db = v & 1;
        mbar_wait(dy_empty(db), ((uint32_t)(v >> 1) & 1u) ^ 1u);
        mbar_expect_tx(dy_full(db), dy_buf_bytes);
        for (int j = 0; j < p.NB; ++j)
          tma_load_4d(dy_base + db * dy_buf_bytes + j * kBlkBytes, &tma_dy, dy_full(db), n_base + 32 * j, w0, h0, d0);
        for (int tp = 0; tp < ntap; ++tp, ++it) {
          const int tap = tap0 + tp;
          const int tkw = tap % p.kw, tkh = (tap / p.kw) % p.kh, tkd = tap / (p.kw * p.kh);
          const int s = it % p.stages;
          mbar_wait(empty_bar(s), ((uint32_t)(it / p.stages) & 1u) ^ 1u);
          mbar_expect_tx(full_bar(s), x_stage_bytes);
          for (int j = 0; j < p.CB; ++j)
            tma_load_4d(base + s * x_stage_bytes + j * kBlkBytes, &tma_x, full_bar(s), c_base + 32 * j,
                        w0 * p.sw + tkw - p.pw, h0 * p.sh + tkh - p.ph, d0 * p.sd + tkd - p.pd);
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
        const uint32_t dy0 = dy_base + db * dy_buf_bytes;
        const uint64_t bd0 = make_mn_desc(dy0, kBlkBytes, 512);
        for (int tp = 0; tp < ntap; ++tp, ++it) {
          const int s = it % p.stages;
          mbar_wait(full_bar(s), (uint32_t)(it / p.stages) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t x0 = base + s * x_stage_bytes;
            const uint32_t dcol = tmem_d + (uint32_t)(tp * p.BN);
            // M = 128 lanes = 4 channel blocks at LBO; with a single block LBO = 0 aliases it (the surplus lanes are
            // never stored) so the tensor core never reads past the stage.  Descriptors are built once per stage and
            // advanced by 1024 B (= 64 in the 16-byte start-address field) per 8-voxel K step.
            const uint64_t ad = make_mn_desc(x0, p.CB == 1 ? 0u : (uint32_t)kBlkBytes, 512);
            umma_tf32(dcol, ad, bd0, p.idesc, v > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < 16; ++k)            // 16 x 8 voxels = the 128-voxel tile
              umma_tf32(dcol, ad + 64u * k, bd0 + 64u * k, p.idesc, 1u);
            umma_commit(empty_bar(s));
            if (tp == ntap - 1) {
              umma_commit(dy_empty(db));
              if (v == nvt - 1) umma_commit(done_bar);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (nvt > 0) {
    // ================= epilogue: TMEM -> workspace[chunk][n][tap][c] =================
    const int q = warp & 3;
    const int c = c_base + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = partial + (int64_t)chunk * p.N * p.taps * p.C;
    for (int tp = 0; tp < ntap; ++tp) {
      for (int nn = 0; nn < p.BN; nn += 16) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * p.BN + nn), v);
        if (c < p.C && q * 32 < p.CB * 32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n_base + nn + i;
            if (n < p.N) dst[((int64_t)n * p.taps + tap0 + tp) * p.C + c] = v[i];
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

__global__ void zero_partial_rows(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
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

static int pow2_at_least(int x, int lo) {
  int v = lo;
  while (v < x) v <<= 1;
  return v;
}

}  // namespace wg

// Plans the split; shared by the workspace query and the launch.
static bool wgrad_tc_plan(const GatherGeom& g, wg::WgParams& p) {
  if (g.sd > 2 || g.sh > 2 || g.sw > 2 || g.transposed) return false;
  p.sd = g.sd; p.sh = g.sh; p.sw = g.sw;
  if ((g.C & 3) || (g.N & 3)) return false;
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N; p.kd = g.kd; p.kh = g.kh; p.kw = g.kw; p.pd = g.pd; p.ph = g.ph; p.pw = g.pw;
  p.taps = g.kd * g.kh * g.kw;
  p.BW = wg::pow2_at_least(g.Wo < 16 ? g.Wo : 16, 1);
  if (p.BW > 16) p.BW = 16;
  p.BH = wg::pow2_at_least(g.Ho < 128 / p.BW ? g.Ho : 128 / p.BW, 1);
  if (p.BH > 128 / p.BW) p.BH = 128 / p.BW;
  p.BD = 128 / (p.BW * p.BH);
  p.tiles_w = (g.Wo + p.BW - 1) / p.BW;
  p.tiles_h = (g.Ho + p.BH - 1) / p.BH;
  p.tiles_d = (g.Do + p.BD - 1) / p.BD;
  p.n_vtiles = p.tiles_w * p.tiles_h * p.tiles_d;
  const int cblocks = (g.C + 31) / 32, nblocks = (g.N + 31) / 32;
  p.CB = cblocks < 4 ? cblocks : 4;
  p.NB = nblocks < 2 ? nblocks : 2;
  p.c_tiles = (cblocks + p.CB - 1) / p.CB;
  p.n_tiles = (nblocks + p.NB - 1) / p.NB;
  const int n_in_tile = g.N < p.NB * 32 ? g.N : p.NB * 32;
  p.BN = (n_in_tile + 15) / 16 * 16;
  p.TG = 512 / p.BN;
  if (p.TG > p.taps) p.TG = p.taps;
  p.tap_groups = (p.taps + p.TG - 1) / p.TG;
  p.tmem_cols = (uint32_t)wg::pow2_at_least(p.TG * p.BN, 32);
  const int base_ctas = p.c_tiles * p.n_tiles * p.tap_groups;
  int want = (148 + base_ctas - 1) / base_ctas;            // one CTA per SM (TMEM + ~190 KB smem each)
  if (want > p.n_vtiles) want = p.n_vtiles;
  if (want < 1) want = 1;
  p.vtiles_per_chunk = (p.n_vtiles + want - 1) / want;
  p.nchunks = (p.n_vtiles + p.vtiles_per_chunk - 1) / p.vtiles_per_chunk;
  const int x_stage = p.CB * wg::kBlkBytes, dy_buf = p.NB * wg::kBlkBytes;
  p.stages = (200 * 1024 - 2 * dy_buf) / x_stage;
  if (p.stages > 6) p.stages = 6;
  if (p.stages < 2) return false;
  // a_major = b_major = MN (bits 15, 16)
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  return true;
}

int64_t conv_tc_wgrad_workspace_bytes(const GatherGeom& g) {
  wg::WgParams p;
  int64_t kwb = conv_tc_wgrad_kw_workspace_bytes(g);
  const int64_t mb = conv_tc_wgrad_march_workspace_bytes(g);
  if (mb > kwb) kwb = mb;
  if (!wgrad_tc_plan(g, p)) return kwb;
  const int64_t v1 = (int64_t)p.nchunks * g.N * p.taps * g.C * (int64_t)sizeof(float);
  return v1 > kwb ? v1 : kwb;
}

// returns DPI_ERR_UNSUPPORTED when the shape is not covered; *nchunks_out = number of partial slabs written
int conv_tc_wgrad(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* partial, int64_t partial_bytes,
                  const GatherGeom& g, int* nchunks_out, cudaStream_t st) {
  {
    // 3x3(x3), stride 1, few channel blocks: nine taps per MMA, marching along d (conv_tc_wgrad_march.cu)
    const int rc = conv_tc_wgrad_march(x, x_ld, dy, dy_ld, partial, partial_bytes, g, nchunks_out, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  {
    // kw = 3, stride 1: three taps per MMA (conv_tc_wgrad_kw.cu)
    static int kw_enabled = -1;
    if (kw_enabled < 0) { const char* e = getenv("DPI_TC_WGRAD_KW"); kw_enabled = (e && e[0] == '0') ? 0 : 1; }
    if (kw_enabled) {
      const int rc = conv_tc_wgrad_kw(x, x_ld, dy, dy_ld, partial, partial_bytes, g, nchunks_out, st);
      if (rc != DPI_ERR_UNSUPPORTED) return rc;
    }
  }
  wg::WgParams p;
  if (!wgrad_tc_plan(g, p)) return DPI_ERR_UNSUPPORTED;
  static int device_ok = -1;
  if (device_ok < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    device_ok = dpi_device_supports_tcgen05(dev);
  }
  wg::EncodeTiledFn encode = wg::get_encode();
  if (!device_ok || !encode) return DPI_ERR_UNSUPPORTED;
  const int64_t need = (int64_t)p.nchunks * g.N * p.taps * g.C * (int64_t)sizeof(float);
  if (partial_bytes < need) {
    set_error("conv_tc_wgrad: workspace too small (%lld < %lld)", (long long)partial_bytes, (long long)need);
    return DPI_ERR_WORKSPACE;
  }
  CUtensorMap mx, mdy;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)g.Wi * x_ld * 4, (cuuint64_t)g.Hi * g.Wi * x_ld * 4};
    // strided conv: the box spans B*s input positions, the TMA element stride keeps every s-th one
    cuuint32_t box[4] = {32, (cuuint32_t)(p.BW * g.sw), (cuuint32_t)(p.BH * g.sh), (cuuint32_t)(p.BD * g.sd)};
    cuuint32_t es[4] = {1, (cuuint32_t)g.sw, (cuuint32_t)g.sh, (cuuint32_t)g.sd};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(x) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.N, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)g.Do};
    cuuint64_t strides[3] = {(cuuint64_t)dy_ld * 4, (cuuint64_t)g.Wo * dy_ld * 4, (cuuint64_t)g.Ho * g.Wo * dy_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BD};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return DPI_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.CB * wg::kBlkBytes + 2 * (size_t)p.NB * wg::kBlkBytes + 8 * (2 * p.stages + 8) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(wg::conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  const unsigned grid = (unsigned)(p.nchunks * p.c_tiles * p.n_tiles * p.tap_groups);
  wg::conv_tc_wgrad_kernel<<<grid, wg::kWgThreads, smem, st>>>(mx, mdy, partial, p);
  *nchunks_out = p.nchunks;
  return check_launch("conv_tc_wgrad_kernel");
}

}  // namespace dpi
