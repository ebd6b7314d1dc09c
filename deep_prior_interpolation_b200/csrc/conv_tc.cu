// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (DPI_PREC_TF32).
//
// Forward and stride-1 data-gradient of the 3x3x3 / 1x1x1 (and 2-D 3x3 / 1x1) convolutions as ONE kernel:
//     out[v][n] = sum_{tap,c} in[v + off(tap)][c] * Wp[n][tap][c]
// GEMM view per CTA:  M = 128 output voxels (a BD x BH x BW spatial box), N = BN output channels, K = taps*C.
//
//   * A tiles come straight from the channels-last activation through a rank-4 TMA tensor map
//     (C, W, H, D): one box load per (tap, channel chunk) at the tap-shifted coordinate.  Out-of-bounds
//     elements (the zero padding of nn.Conv3d, the ragged last tile, the channel tail) are zero-filled by TMA,
//     so there is no im2col buffer, no halo copy and no bounds check in the inner loop.
//   * B tiles (packed weights [N][taps*C], K-major) come through a rank-2 tensor map.
//   * Both land in shared memory in the canonical K-major swizzled layout (SWIZZLE_32B/64B/128B by channel
//     chunk 8/16/32) that tcgen05.mma consumes through 64-bit matrix descriptors.
//   * One elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8 per instruction), fp32
//     accumulator in TMEM; tcgen05.commit hands shared-memory stages back to the TMA producer and signals
//     the epilogue warps, which read TMEM with tcgen05.ld (32 lanes x 32 bit), add the bias and store.
//   * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
//
// Shapes not covered (stride 2, C or N not multiples of 4) return DPI_ERR_UNSUPPORTED and the caller uses the
// CUDA-core kernel of conv_simt.cu.
#include <cuda.h>
#include "conv_geom.cuh"

namespace dpi {

// ---------------------------------------------------------------- PTX wrappers -------------------------------
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
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
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

// K-major shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: UMMA::SmemDescriptor):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major, set to 1)
//   [32,46) stride byte offset >> 4 (distance between 8-row groups)   [46,48) version = 1   [61,64) layout type
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

struct TcParams {
  int Do, Ho, Wo;
  int tiles_w, tiles_h, tiles_d;
  int BD, BH, BW;
  int C, N;
  int kd, kh, kw, pd, ph, pw, transposed;
  int sd, sh, sw;                 // conv stride per axis
  int ncd, nch, ncw;              // parity classes per axis (= stride for a strided dgrad, else 1); blockIdx.z
  int KC, G, chunks_per_tap;
  int BN, stages;
  int a_sub_bytes, b_sub_bytes;
  int sbo_bytes, layout_type;
  uint32_t idesc, tmem_cols;
  int64_t out_ld;
  int accumulate;
};

constexpr int kTcThreads = 192;

__global__ void __launch_bounds__(kTcThreads)
conv_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const float* __restrict__ bias, float* __restrict__ out, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)p.G * (p.a_sub_bytes + p.b_sub_bytes);
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;       // 8-byte aligned (multiples of 1024)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 1);
  auto a_addr = [&](int s, int j) { return base + s * stage_bytes + j * p.a_sub_bytes; };
  auto b_addr = [&](int s, int j) { return base + s * stage_bytes + p.G * p.a_sub_bytes + j * p.b_sub_bytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h;
  const int td = t / p.tiles_h;
  const int w0 = tw * p.BW, h0 = th * p.BH, d0 = td * p.BD;
  const int n0 = blockIdx.y * p.BN;
  // Strided data-gradient = one stride-1 gather per output parity class: class (cd,ch,cw) owns the outputs
  // 2u+c and only the taps with (c + p - t) divisible by the stride contribute, reading dy[u + (c+p-t)/s].
  const int cls = blockIdx.z;
  const int cw = cls % p.ncw, ch = (cls / p.ncw) % p.nch, cd = cls / (p.ncw * p.nch);
  __shared__ int s_tap[27], s_off[27];
  // every thread counts the contributing taps itself (block-uniform arithmetic, so the MMA warp's loop bounds stay
  // in uniform registers); thread 0 also records them
  int ntaps_cls = 0;
  for (int tkd = 0; tkd < p.kd; ++tkd)
    for (int tkh = 0; tkh < p.kh; ++tkh)
      for (int tkw = 0; tkw < p.kw; ++tkw) {
        int dd, dh, dw;
        bool ok = true;
        if (!p.transposed) {
          dd = tkd - p.pd; dh = tkh - p.ph; dw = tkw - p.pw;
        } else {
          const int nd = cd + p.pd - tkd, nh = ch + p.ph - tkh, nw = cw + p.pw - tkw;
          ok = (nd % p.sd == 0) && (nh % p.sh == 0) && (nw % p.sw == 0);
          dd = nd / p.sd; dh = nh / p.sh; dw = nw / p.sw;
        }
        if (ok) {
          if (threadIdx.x == 0) {
            s_tap[ntaps_cls] = (tkd * p.kh + tkh) * p.kw + tkw;
            s_off[ntaps_cls] = ((dd + 8) << 8) | ((dh + 8) << 4) | (dw + 8);
          }
          ++ntaps_cls;
        }
      }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_d;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));
  const int n_sub = ntaps_cls * p.chunks_per_tap;
  const int n_iters = (n_sub + p.G - 1) / p.G;
  // coordinate scale of the A tensor: a strided FORWARD conv reads x at u*s + (t - p) (TMA element strides do
  // the striding inside the box); a dgrad reads dy at u + off
  const int csw = p.transposed ? 1 : p.sw, csh = p.transposed ? 1 : p.sh, csd = p.transposed ? 1 : p.sd;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      const uint32_t sub_bytes = 128u * p.KC * 4u + (uint32_t)p.BN * p.KC * 4u;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int nsub = min(p.G, n_sub - it * p.G);
        mbar_expect_tx(full_bar(s), sub_bytes * nsub);
        for (int j = 0; j < nsub; ++j) {
          const int sub = it * p.G + j;
          const int ti = sub / p.chunks_per_tap;
          const int c0 = (sub - ti * p.chunks_per_tap) * p.KC;
          const int tap = s_tap[ti], off = s_off[ti];
          const int dw = (off & 15) - 8, dh = ((off >> 4) & 15) - 8, dd = (off >> 8) - 8;
          tma_load_4d(a_addr(s, j), &tma_a, full_bar(s), c0, w0 * csw + dw, h0 * csh + dh, d0 * csd + dd);
          tma_load_2d(b_addr(s, j), &tma_b, full_bar(s), tap * p.C + c0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: whole-warp control flow, one elected lane issues (see conv_tc_march.cu) =======
    const int kk = p.KC / 8;
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const int nsub = min(p.G, n_sub - it * p.G);
        uint32_t accum = it > 0 ? 1u : 0u;
        for (int j = 0; j < nsub; ++j) {
          const uint32_t a0 = a_addr(s, j), b0 = b_addr(s, j);
          const uint64_t ad = make_kmajor_desc(a0, p.sbo_bytes, p.layout_type);
          const uint64_t bd = make_kmajor_desc(b0, p.sbo_bytes, p.layout_type);
          // (straight-line per K-step count: a rolled loop costs the issuing lane far more than the MMAs themselves,
          //  see conv_tc_march.cu issue_packed)
          umma_tf32(tmem_d, ad, bd, p.idesc, accum);
          if (kk == 4) {
            umma_tf32(tmem_d, ad + 2u, bd + 2u, p.idesc, 1u);
            umma_tf32(tmem_d, ad + 4u, bd + 4u, p.idesc, 1u);
            umma_tf32(tmem_d, ad + 6u, bd + 6u, p.idesc, 1u);
          } else if (kk == 2) {
            umma_tf32(tmem_d, ad + 2u, bd + 2u, p.idesc, 1u);
          }
          accum = 1;
        }
        umma_commit(empty_bar(s));       // frees the stage once these MMAs have read it
        if (it == n_iters - 1) umma_commit(tmem_full_bar);        // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int w = row % p.BW, h = (row / p.BW) % p.BH, d = row / (p.BW * p.BH);
    const int ow = p.transposed ? (w0 + w) * p.sw + cw : w0 + w;
    const int oh = p.transposed ? (h0 + h) * p.sh + ch : h0 + h;
    const int od = p.transposed ? (d0 + d) * p.sd + cd : d0 + d;
    const bool valid = ow < p.Wo && oh < p.Ho && od < p.Do;
    float* orow = out + (((int64_t)od * p.Ho + oh) * p.Wo + ow) * p.out_ld + n0;
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
    tmem_dealloc(tmem_d, p.tmem_cols);
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

static int pow2_at_least(int x, int lo) {
  int v = lo;
  while (v < x) v <<= 1;
  return v;
}

int conv_tc_gather(const float* in, int64_t in_ld, const float* Wp, const float* bias, float* out, int64_t out_ld,
                   const GatherGeom& g, int accumulate, cudaStream_t st) {
  if ((g.C & 3) || (g.N & 3) || g.C < 4 || g.sd > 2 || g.sh > 2 || g.sw > 2) return DPI_ERR_UNSUPPORTED;
  static int device_ok = -1;
  if (device_ok < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    device_ok = dpi_device_supports_tcgen05(dev);
  }
  EncodeTiledFn encode = get_encode();
  if (!device_ok || !encode) {
    set_error("tcgen05 path unavailable on this device/driver");
    return DPI_ERR_UNSUPPORTED;
  }
  if (g.kd == 1 && g.kh == 1 && g.kw == 1) {
    const int rc = conv_tc_march_1x1(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  if (g.transposed && g.sh == 2) {
    // data gradient of a stride-2 conv: all eight output parity classes in one launch (conv_tc_halo.cu)
    const int rc = conv_tc_halo_gather(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  if (g.transposed && g.sh == 2) {
    // ... or one march per output parity class
    const int rc = conv_tc_march_dgrad_s2(in, in_ld, Wp, out, out_ld, g, accumulate, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  if (g.N == 4) {
    // four output channels: role-swapped kernel (conv_tc_swap.cu)
    const int rc = conv_tc_swap_gather(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  {
    // 3x3(x3) kernels whose weights fit in shared memory: persistent column march (conv_tc_march.cu)
    const int rc = conv_tc_march_gather(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  {
    // other 3x3(x3) kernels: shared-memory halo reuse (conv_tc_halo.cu); everything else: per-tap loads below
    const int rc = conv_tc_halo_gather(in, in_ld, Wp, bias, out, out_ld, g, accumulate, st);
    if (rc != DPI_ERR_UNSUPPORTED) return rc;
  }
  TcParams p;
  p.Do = g.Do; p.Ho = g.Ho; p.Wo = g.Wo;
  p.C = g.C; p.N = g.N;
  p.kd = g.kd; p.kh = g.kh; p.kw = g.kw; p.pd = g.pd; p.ph = g.ph; p.pw = g.pw; p.transposed = g.transposed;
  p.sd = g.sd; p.sh = g.sh; p.sw = g.sw;
  p.ncd = g.transposed ? g.sd : 1; p.nch = g.transposed ? g.sh : 1; p.ncw = g.transposed ? g.sw : 1;
  // tiles live in "u space": the outputs themselves, or for a strided dgrad the outputs of one parity class
  const int Uw = (g.Wo + p.ncw - 1) / p.ncw, Uh = (g.Ho + p.nch - 1) / p.nch, Ud = (g.Do + p.ncd - 1) / p.ncd;
  // spatial box of 128 output voxels
  p.BW = pow2_at_least(Uw < 16 ? Uw : 16, 1);
  if (p.BW > 16) p.BW = 16;
  p.BH = pow2_at_least(Uh < 128 / p.BW ? Uh : 128 / p.BW, 1);
  if (p.BH > 128 / p.BW) p.BH = 128 / p.BW;
  p.BD = 128 / (p.BW * p.BH);
  p.tiles_w = (Uw + p.BW - 1) / p.BW;
  p.tiles_h = (Uh + p.BH - 1) / p.BH;
  p.tiles_d = (Ud + p.BD - 1) / p.BD;
  // channel chunk = shared-memory row.  Channels past C are TMA out-of-bounds elements: they cost shared memory
  // and MMA K-steps but no memory traffic, while narrow rows cost TMA efficiency (32-byte rows made the 1x1 convs
  // with C = 72/144/280 run at half their HBM bound) - so: the widest row that C fills at least once
  const int best_kc = g.C > 16 ? 32 : (g.C > 8 ? 16 : 8);
  p.KC = best_kc;
  p.G = 32 / p.KC;
  p.chunks_per_tap = (g.C + p.KC - 1) / p.KC;
  const int taps = g.kd * g.kh * g.kw;
  const int n_iters_max = (taps * p.chunks_per_tap + p.G - 1) / p.G;
  const int n_tiles = (g.N + 255) / 256;
  p.BN = (((g.N + n_tiles - 1) / n_tiles) + 15) / 16 * 16;
  p.a_sub_bytes = 128 * p.KC * 4;
  p.b_sub_bytes = (p.BN * p.KC * 4 + 1023) / 1024 * 1024;
  p.sbo_bytes = 8 * p.KC * 4;
  p.layout_type = p.KC == 32 ? 2 : (p.KC == 16 ? 4 : 6);
  const CUtensorMapSwizzle swz = p.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (p.KC == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const int stage_bytes = p.G * (p.a_sub_bytes + p.b_sub_bytes);
  p.stages = 200 * 1024 / stage_bytes;
  if (p.stages > 6) p.stages = 6;
  if (p.stages > n_iters_max) p.stages = n_iters_max < 1 ? 1 : n_iters_max;
  if (p.stages < 1) return DPI_ERR_UNSUPPORTED;
  p.tmem_cols = (uint32_t)pow2_at_least(p.BN, 32);
  // instruction descriptor (UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13), K-major A/B,
  // N>>3 at [17,23), M>>4 at [24,29)
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.out_ld = out_ld;
  p.accumulate = accumulate;

  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di};
    cuuint64_t strides[3] = {(cuuint64_t)in_ld * 4, (cuuint64_t)g.Wi * in_ld * 4, (cuuint64_t)g.Hi * g.Wi * in_ld * 4};
    // strided forward: the box spans BW*s input positions and the TMA element stride keeps every s-th one
    const cuuint32_t esw = g.transposed ? 1 : g.sw, esh = g.transposed ? 1 : g.sh, esd = g.transposed ? 1 : g.sd;
    cuuint32_t box[4] = {(cuuint32_t)p.KC, (cuuint32_t)p.BW * esw, (cuuint32_t)p.BH * esh, (cuuint32_t)p.BD * esd};
    cuuint32_t es[4] = {1, esw, esh, esd};
    CUresult r = encode(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d D=%d ld=%lld box=%d,%d,%d,%d)", (int)r, g.C, g.Wi,
                g.Hi, g.Di, (long long)in_ld, p.KC, p.BW, p.BH, p.BD);
      return DPI_ERR_CUDA;
    }
  }
  {
    const cuuint64_t Ktot = (cuuint64_t)taps * g.C;
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)g.N};
    cuuint64_t strides[1] = {Ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)p.KC, (cuuint32_t)p.BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Wp), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B) failed: %d (K=%llu N=%d box=%d,%d)", (int)r, (unsigned long long)Ktot, g.N,
                p.KC, p.BN);
      return DPI_ERR_CUDA;
    }
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 8 * (2 * p.stages + 2) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%zu) failed", smem);
      cudaGetLastError();
      return DPI_ERR_CUDA;
    }
    smem_set = smem;
  }
  dim3 grid((unsigned)(p.tiles_w * p.tiles_h * p.tiles_d), (unsigned)n_tiles, (unsigned)(p.ncd * p.nch * p.ncw));
  conv_tc_kernel<<<grid, kTcThreads, smem, st>>>(ma, mb, bias, out, p);
  return check_launch("conv_tc_kernel");
}

}  // namespace dpi
