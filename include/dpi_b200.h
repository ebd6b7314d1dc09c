/*
 * dpi_b200.h — C ABI of the B200-native deep-prior-interpolation hot path.
 *
 * The reference (polimi-ispl/deep_prior_interpolation) has no FFI of its own: every FLOP of its
 * hot loop is a PyTorch library call.  Each entry point below replaces one of those call sites
 * (cited as <file>:<line> of the reference) with a hand-written sm_100a kernel.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - activations are channels-last fp32:  element (d,h,w,c) lives at  ((d*H + h)*W + w)*ld + c,
 *     `ld` (floats) is the channel pitch of the underlying buffer, so a channel slice of a wider
 *     buffer (the skip-concat of architectures/base.py:325-362) is addressed by pointer offset;
 *     channel counts and pitches are multiples of 4 (16 B), pad channels are kept at 0;
 *   - 2-D networks use D == 1 and kd == 1;
 *   - conv weights are in the packed layout  [N][taps][C]  (N = output channels of the GEMM,
 *     taps = kd*kh*kw in (kd,kh,kw) C-order, C = reduction channels), produced by
 *     dpi_pack_conv_weights from the state_dict layout [Cout][Cin][kd][kh][kw];
 *   - all functions are stream-ordered, never allocate, never synchronise, and return
 *     DPI_OK (0) or a negative DPI_ERR_* code; dpi_last_error_string() describes the last failure
 *     of the calling thread;
 *   - `stream` is a cudaStream_t passed as void*.
 */
#ifndef DPI_B200_H_
#define DPI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPI_OK 0
#define DPI_ERR_INVALID_ARG (-1)
#define DPI_ERR_CUDA (-2)
#define DPI_ERR_UNSUPPORTED (-3)
#define DPI_ERR_WORKSPACE (-4)

/* activation codes — architectures/base.py:97-114 (get_activation) */
#define DPI_ACT_NONE 0
#define DPI_ACT_LEAKY_RELU 1 /* slope 0.2 */
#define DPI_ACT_RELU 2
#define DPI_ACT_ELU 3
#define DPI_ACT_TANH 4
#define DPI_ACT_SIGMOID 5
/* OR-ed into an `act` / upsample `mode` / loss `kind` argument: round what the kernel stores to TF32 (RNA),
 * because a tcgen05 kind::tf32 MMA reads it next and would otherwise truncate the mantissa */
#define DPI_ACT_ROUND_TF32 0x100

/* conv precision */
#define DPI_PREC_FP32 0 /* CUDA-core FFMA implicit GEMM, exact fp32 */
#define DPI_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM */

/* loss kinds — main.py:24-27 */
#define DPI_LOSS_MAE 0
#define DPI_LOSS_MSE 1

/* upsample modes — mulresunet.py:168,242 (nn.Upsample scale_factor=2) */
#define DPI_UP_NEAREST 0
#define DPI_UP_LINEAR 1 /* bi/tri-linear, align_corners=False */

const char* dpi_last_error_string(void);
int dpi_version(void);
/* number of kernel launches issued through this library by the calling process so far */
int64_t dpi_launch_count(void);
/* 1 if the device `dev` can run the tcgen05 path (compute capability 10.x) */
int dpi_device_supports_tcgen05(int dev);

/* ---------------------------------------------------------------- convolution ------------- */
/* Geometry shared by the three conv entry points: the FORWARD convolution
 *   y[do,ho,wo,n] = bias[n] + sum_{kd,kh,kw,c} x[do*s+kd-pd, ho*s+kh-ph, wo*s+kw-pw, c] * w[n][tap][c]
 * with zero padding p = (k-1)/2 and output size (D+2p-k)/s+1  (nn.Conv3d / nn.Conv2d,
 * architectures/base.py:117-126,169-180).  Spatial stride applies to axes with k>1 or all axes of
 * a 3-D net; for 2-D nets (D==1,kd==1) the D axis is untouched. */
typedef struct {
  int32_t D, H, W;       /* input spatial size */
  int32_t Cin, Cout;     /* physical (padded) channel counts */
  int32_t kd, kh, kw;    /* kernel size (1 or 3 per axis) */
  int32_t stride;        /* 1 or 2 */
} dpi_conv_geom;

/* forward: x [D,H,W,x_ld] -> y [Do,Ho,Wo,y_ld]; w packed [Cout][taps][Cin]; bias may be NULL.
 * Replaces nn.Conv3d/Conv2d forward (base.py:123,176). */
int dpi_conv_fwd(const float* x, int64_t x_ld, const float* w, const float* bias, float* y,
                 int64_t y_ld, const dpi_conv_geom* g, int precision, void* stream);

/* dpi_conv_fwd that ALSO leaves the per-channel sum / sum-of-squares partial rows of y in `stats_ws` (the workspace
 * dpi_bn_finalize reads, sized by dpi_stats_workspace_bytes for Cout channels), for a conv that feeds a training-mode BatchNorm
 * (base.py:162-166,211-216: conv -> nn.BatchNorm*d).  The tcgen05 march kernels accumulate them in their epilogue
 * (fp64, fixed combination order, one row per CTA); every other path runs dpi_channel_stats over y after the conv.
 * stats_ws == NULL is dpi_conv_fwd. */
int dpi_conv_fwd_stats(const float* x, int64_t x_ld, const float* w_packed, const float* bias, float* y,
                       int64_t y_ld, const dpi_conv_geom* geom, int precision, void* stats_ws, void* stream);
/* data gradient: dy [Do,Ho,Wo,dy_ld] -> dx [D,H,W,dx_ld]; wt packed [Cin][taps][Cout]
 * (the transposed pack of the same weights); accumulate!=0 adds into dx.
 * Replaces the dgrad half of total_loss.backward() (main.py:162). */
int dpi_conv_dgrad(const float* dy, int64_t dy_ld, const float* wt, float* dx, int64_t dx_ld,
                   const dpi_conv_geom* g, int accumulate, int precision, void* stream);
/* Fused data gradient of TWO convolutions that read the same input x (Block3d.conv3x3 + Block3d.shortcut,
 * mulresunet.py:85-92; ResPath3d.conv3x3 + ResPath3d.conv1x1, mulresunet.py:108-109): dy holds the output gradients
 * of both side by side ([nvox][dy_ld]: the 3x3(x3) conv's channels first, then the 1x1 conv's), wt their transposed
 * weights [Cin][taps][Cout_a + Cout_b] with the 1x1 weights at the centre tap and zeros elsewhere, g->Cout =
 * Cout_a + Cout_b.  dx (+)= dgrad_a + dgrad_b in ONE pass: no read-modify-write of dx between the two.
 * full_tap_channels = Cout_a tells the tcgen05 kernels which reduction channels to skip at the non-centre taps. */
int dpi_conv_dgrad_fused(const float* dy, int64_t dy_ld, const float* wt, float* dx, int64_t dx_ld,
                         const dpi_conv_geom* g, int full_tap_channels, int accumulate, int precision, void* stream);
/* 1 when the fused form runs on the kernels that skip the zero blocks (else the two separate dgrads are cheaper) */
int dpi_conv_dgrad_fused_supported(const dpi_conv_geom* g, int full_tap_channels);

/* weight gradient: dw packed [Cout][taps][Cin] = sum_v dy[v][n] * x[src(v,tap)][c]; split over
 * voxel chunks into `workspace` and reduced in a fixed order (bit-reproducible).
 * Replaces the wgrad half of total_loss.backward() (main.py:162). */
int64_t dpi_conv_wgrad_workspace_bytes(const dpi_conv_geom* g);
int dpi_conv_wgrad(const float* x, int64_t x_ld, const float* dy, int64_t dy_ld, float* dw,
                   const dpi_conv_geom* g, void* workspace, int64_t workspace_bytes, int precision,
                   void* stream);

/* pack state_dict weights [Cout_l][Cin_l][taps] into w_fwd [Cout_p][taps][Cin_p] and (optional)
 * w_dgrad [Cin_p][taps][Cout_p]; cout_map/cin_map give for every physical channel its logical
 * index or -1 (pad -> zero).  round_tf32 applies cvt.rna.tf32 to the packed copies.  When
 * bias_packed != NULL the bias [Cout_l] is scattered to physical order too (0 for pads / NULL). */
int dpi_pack_conv_weights(const float* w, const int32_t* cout_map, const int32_t* cin_map,
                          int Cout_l, int Cin_l, int Cout_p, int Cin_p, int taps, float* w_fwd,
                          float* w_dgrad, const float* bias, float* bias_packed, int round_tf32,
                          void* stream);
/* inverse gather for the gradient: dw_packed [Cout_p][taps][Cin_p] -> dw [Cout_l][Cin_l][taps] */
int dpi_unpack_conv_wgrad(const float* dw_packed, const int32_t* cout_map, const int32_t* cin_map,
                          int Cout_l, int Cin_l, int Cout_p, int Cin_p, int taps, float* dw,
                          void* stream);
/* Batched forms: one launch packs (or un-packs) every conv layer of a network.  `jobs_dev` is a DEVICE array of
 * dpi_pack_job records - pointers as in the single-layer calls, unused ones NULL; Cout_p*taps*Cin_p < 2^31 per job. */
typedef struct dpi_pack_job {
  const float* w;              /* state_dict weight [Cout_l][Cin_l][taps]            (pack) */
  const float* bias;           /* [Cout_l] or NULL                                   (pack) */
  float* w_fwd;                /* [Cout_p][taps][Cin_p] or NULL                      (pack) */
  float* w_dgrad;              /* [Cin_p][taps][Cout_p] or NULL                      (pack) */
  float* bias_packed;          /* [Cout_p] or NULL                                   (pack) */
  const float* dw_packed;      /* [Cout_p][taps][Cin_p]                              (unpack) */
  float* dw;                   /* gradient in state_dict layout [Cout_l][Cin_l][taps] (unpack) */
  const int32_t* cout_map;
  const int32_t* cin_map;
  int32_t Cout_l, Cin_l, Cout_p, Cin_p, taps, reserved;
  /* optional: this layer's slice of the transposed weights of a FUSED data gradient (dpi_conv_dgrad_fused),
   * [Cin_p][cat_taps][cat_ld]: tap t of this layer goes to tap cat_tap0 + t, its output channels to
   * [cat_off, cat_off + Cout_p); entries no layer writes must be zero (the buffer is zero-filled once) */
  float* w_dgrad_cat;
  int32_t cat_ld, cat_off, cat_taps, cat_tap0;
} dpi_pack_job;
int dpi_pack_conv_weights_batched(const dpi_pack_job* jobs_dev, int njobs, int round_tf32, void* stream);
int dpi_unpack_conv_wgrad_batched(const dpi_pack_job* jobs_dev, int njobs, void* stream);
/* per-channel sum of a channels-last tensor scattered through map: out[map[c]] = sum_v x[v][c]
 * (bias gradient of convs that are not followed by a BatchNorm). */
int dpi_bias_grad(const float* dy, int64_t ld, int64_t nvox, int C, const int32_t* map, float* db,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- batch norm + pointwise -- */
/* bytes of the partial-sum workspace the statistics-producing kernels need for C channels */
int64_t dpi_stats_workspace_bytes(int C);

/* per-channel (sum, sum of squares) of x over nvox voxels -> stats workspace (fp64 partials) */
int dpi_channel_stats(const float* x, int64_t ld, int64_t nvox, int C, void* stats_ws, void* stream);

/* training-mode BatchNorm statistics (nn.BatchNorm3d/2d, base.py:164,214; mulresunet.py:80-81,104,225):
 * reduces the stats workspace in a fixed order, produces mean/invstd and the fused affine
 *   scale[p] = gamma[map[p]]*invstd[p],  shift[p] = beta[map[p]]   (0 for pads), to be applied as
 *   y = (x - mean[p])*scale[p] + shift[p]  (same operation order as PyTorch's CPU kernel),
 * and updates running_mean / running_var (unbiased) / num_batches_tracked like PyTorch. */
int dpi_bn_finalize(const void* stats_ws, int64_t nvox, int C, const int32_t* map,
                    const float* gamma, const float* beta, float* running_mean, float* running_var,
                    int64_t* num_batches_tracked, float momentum, float eps, float* mean,
                    float* invstd, float* scale, float* shift, void* stream);
/* dpi_bn_finalize over the statistics of a tensor whose channel ranges were produced by DIFFERENT kernels: part i
 * (channels [cbegin[i], cbegin[i+1]) of the C channels) has its own workspace, with rows laid out for the width of that
 * part.  The branch outputs of a MultiRes block (torch.cat of mulresunet.py:89 followed by BatchNorm, :90) are written by
 * three BatchNorm + activation passes that each leave the statistics of their own output (dpi_affine_act with a stats
 * workspace): the separate statistics pass over the concatenation (dpi_channel_stats_parts) is not launched. */
typedef struct dpi_stats_parts {
  const void* ws[4];
  int32_t cbegin[5];
  int32_t n;
} dpi_stats_parts;
int dpi_bn_finalize_parts(const dpi_stats_parts* stats, int64_t nvox, int C, const int32_t* map,
                          const float* gamma, const float* beta, float* running_mean, float* running_var,
                          int64_t* num_batches_tracked, float momentum, float eps, float* mean,
                          float* invstd, float* scale, float* shift, void* stream);


/* y = act((x-mean)*scale + shift)   (mean/scale/shift NULL -> 0/1/0); optional statistics of y */
int dpi_affine_act(const float* x, int64_t x_ld, const float* mean, const float* scale,
                   const float* shift, int act, float* y, int64_t y_ld, int64_t nvox, int C,
                   void* stats_ws_or_null, void* stream);
/* y = act(p + ((q-mean)*scale + shift))  — the residual adds of mulresunet.py:92-93,109-110 */
int dpi_add_affine_act(const float* p, int64_t p_ld, const float* q, int64_t q_ld, const float* mean,
                       const float* scale, const float* shift, int act, float* y, int64_t y_ld,
                       int64_t nvox, int C, void* stats_ws_or_null, void* stream);

/* A channels-last tensor whose channel ranges live in SEPARATE dense buffers: part i holds the physical channels
 * [cbegin[i], cbegin[i+1]) of the tensor, 1 <= n <= 4, cbegin[0] = 0, cbegin[n] = C; widths, offsets and pitches
 * are multiples of 4.  This is how the branch outputs of a MultiRes block (torch.cat of mulresunet.py:31,89) reach
 * the residual add: a concat buffer with 4/8/16-channel slices at a 112-byte pitch made every slice access move
 * (nearly) the whole buffer through L2/HBM (ncu: 473 MB read for a 67 MB operand). */
typedef struct dpi_parts {
  const float* ptr[4];
  int64_t ld[4];
  int32_t cbegin[5];
  int32_t n;
} dpi_parts;
/* multi-part forms of dpi_channel_stats / dpi_add_affine_act (q in parts) / dpi_bn_bwd_reduce (x in parts) /
 * dpi_bn_bwd_apply (x and dx in parts; bit i of accumulate_mask: dx part i is accumulated into; dp, when not NULL,
 * additionally receives g = dy*act'(out), the gradient of the other addend p of y = act(p + BN(q))) */
int dpi_channel_stats_parts(const dpi_parts* x, int64_t nvox, int C, void* stats_ws, void* stream);
int dpi_add_affine_act_parts(const float* p, int64_t p_ld, const dpi_parts* q, const float* mean,
                             const float* scale, const float* shift, int act, float* y, int64_t y_ld,
                             int64_t nvox, int C, void* stats_ws_or_null, void* stream);
int dpi_bn_bwd_reduce_parts(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act,
                            const dpi_parts* x, const float* mean, const float* invstd, int64_t nvox, int C,
                            void* stats_ws, void* stream);
int dpi_bn_bwd_apply_parts(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act,
                           const dpi_parts* x, const float* mean, const float* invstd, const float* scale,
                           const float* c1, const float* c2, const dpi_parts* dx, int accumulate_mask,
                           float* dp_or_null, int64_t dp_ld, int64_t nvox, int C, void* stream);

/* g = dy * act'(out)  (derivative expressed through the activation OUTPUT) */
int dpi_act_bwd(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act, float* g,
                int64_t g_ld, int64_t nvox, int C, int accumulate, void* stream);

/* BatchNorm backward, pass 1: with g = dy*act'(out) and xhat = (x-mean)*invstd accumulate sum(g), sum(g*xhat)
 * per channel into the stats workspace.  `out` may be NULL: with scale/shift (the forward's per-channel
 * gamma*invstd and beta) the activation output is re-derived from x as act((x-mean)*scale+shift), which saves
 * one tensor read; with scale == shift == NULL the activation is skipped (g = dy). */
int dpi_bn_bwd_reduce(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act,
                      const float* x, int64_t x_ld, const float* mean, const float* invstd,
                      const float* scale_or_null, const float* shift_or_null,
                      int64_t nvox, int C, void* stats_ws, void* stream);
/* pass 1b: dgamma[map[p]] = sum(g*xhat), dbeta[map[p]] = sum(g); c1 = sum(g)/M, c2 = sum(g*xhat)/M */
int dpi_bn_bwd_finalize(const void* stats_ws, int64_t nvox, int C, const int32_t* map, float* dgamma,
                        float* dbeta, float* c1, float* c2, void* stream);
/* pass 2: dx (+)= scale*(g - c1 - xhat*c2); out == NULL with shift given: activation output re-derived from x */
int dpi_bn_bwd_apply(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act,
                     const float* x, int64_t x_ld, const float* mean, const float* invstd,
                     const float* scale, const float* shift_or_null, const float* c1, const float* c2,
                     float* dx, int64_t dx_ld, int64_t nvox, int C, int accumulate, void* stream);

/* Pass 2 with the NEXT unit's pass 1 fused in.  In the backward pass of a MultiRes block (mulresunet.py:85-96:
 * y = norm2(act(norm1(cat) + shortcut))) the gradient a BatchNorm-backward apply writes is exactly the incoming gradient
 * of the unit that follows, so the kernel that produces it also accumulates that unit's sums and its separate reduce
 * pass (dpi_bn_bwd_reduce*) is not launched:
 *   kind 1: g' = dx * act'(x)  - this unit's input x is the next unit's activation output (norm2 -> act(norm1 + shortcut));
 *   kind 2: g' = dp [* act'(out') with out' = act((x'-mean')*scale'+shift') when scale'/shift' are given] - the second
 *           output of dpi_bn_bwd_apply_parts, the gradient of the other addend (the block's shortcut conv + BN).
 * The stats workspace receives per-CTA rows of (sum g', sum g'*xhat') with xhat' = (x'-mean')*invstd' of the next
 * BatchNorm; dpi_bn_bwd_finalize + dpi_bn_bwd_apply* of the next unit follow unchanged.
 *   kind 3: no sums.  The next unit is a residual add WITHOUT BatchNorm, out = act(p + q) (ResPath*, mulresunet.py:108-112):
 *           its backward g = dx * act'(x) is the gradient of BOTH addends, so `dx` receives g (pass q.grad) and so does the
 *           tensor x' names (part 0: pass p.grad - it is WRITTEN); the add's two dpi_act_bwd passes are not launched.
 *           mean / invstd / scale / shift / stats_ws are ignored. */
typedef struct dpi_bn_next_reduce {
  int32_t kind;
  int32_t act;              /* activation code of the next unit */
  dpi_parts x;              /* x' : input of the next BatchNorm (its conv output / the concatenated branches) */
  const float* mean;        /* next BatchNorm: per-channel mean', invstd' (dpi_bn_finalize) */
  const float* invstd;
  const float* scale;       /* kind 2, next unit with an activation: gamma'*invstd' and beta'; else NULL */
  const float* shift;
  void* stats_ws;
} dpi_bn_next_reduce;
int dpi_bn_bwd_apply_next(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act,
                          const float* x, int64_t x_ld, const float* mean, const float* invstd,
                          const float* scale, const float* shift_or_null, const float* c1, const float* c2,
                          float* dx, int64_t dx_ld, int64_t nvox, int C, int accumulate,
                          const dpi_bn_next_reduce* next, void* stream);
int dpi_bn_bwd_apply_parts_next(const float* dy, int64_t dy_ld, const float* out, int64_t out_ld, int act,
                                const dpi_parts* x, const float* mean, const float* invstd, const float* scale,
                                const float* c1, const float* c2, const dpi_parts* dx, int accumulate_mask,
                                float* dp_or_null, int64_t dp_ld, int64_t nvox, int C,
                                const dpi_bn_next_reduce* next, void* stream);

/* ---------------------------------------------------------------- upsample / layout ------- */
/* x2 upsample (mulresunet.py:168,242) written into a channel slice; only the first (Do,Ho,Wo)
 * outputs are produced, which is the centre-crop of Concat/Concat3D (base.py:302-319,342-357).
 * up_d == 0 leaves the D axis untouched (2-D nets). */
int dpi_upsample2x_fwd(const float* x, int64_t x_ld, int D, int H, int W, float* y, int64_t y_ld,
                       int Do, int Ho, int Wo, int C, int mode, int up_d, void* stream);
int dpi_upsample2x_bwd(const float* dy, int64_t dy_ld, int Do, int Ho, int Wo, float* dx,
                       int64_t dx_ld, int D, int H, int W, int C, int mode, int up_d,
                       int accumulate, void* stream);
/* Grid-attention gate of the attention MultiRes U-Net (architectures/attention.py:86-113, `return x * psi`):
 *   y[v][c] = x[v][c] * psi[v][0]
 * psi is a channels-last tensor padded to 4 channels whose channel 0 holds the (already up-sampled) attention map;
 * y may be a channel slice of the decoder's concat buffer (`concat([att(g, x), up(g)])`, attention.py:258-261).
 * `flags`: DPI_ACT_ROUND_TF32 or 0. */
int dpi_gate_mul_fwd(const float* x, int64_t x_ld, const float* psi, int64_t psi_ld, float* y, int64_t y_ld,
                     int64_t nvox, int C, int flags, void* stream);
/* its backward (the x * psi node of total_loss.backward(), main.py:162):
 *   dx[v][c] (+)= dy[v][c] * psi[v][0];   dpsi[v][0] = sum_c dy[v][c] * x[v][c],  dpsi[v][1..3] = 0
 * (fixed summation order, no atomics) */
int dpi_gate_mul_bwd(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld, const float* psi, int64_t psi_ld,
                     float* dx, int64_t dx_ld, float* dpsi, int64_t dpsi_ld, int64_t nvox, int C, int accumulate_dx,
                     void* stream);
/* y[v][c] (+)= x[v][c] for a channel slice */
int dpi_copy_slice(const float* x, int64_t x_ld, float* y, int64_t y_ld, int64_t nvox, int C,
                   int accumulate, void* stream);
/* NCDHW (PyTorch, [C_l][nvox]) <-> channels-last [nvox][ld]; map gives logical channel per
 * physical channel (-1 -> 0 on the way in, skipped on the way out) */
int dpi_nchw_to_cl(const float* src, int C_l, int64_t nvox, const int32_t* map, float* dst,
                   int64_t ld, int C_p, void* stream);
int dpi_cl_to_nchw(const float* src, int64_t ld, int C_p, const int32_t* map, float* dst, int C_l,
                   int64_t nvox, void* stream);

/* ---------------------------------------------------------------- loop body pieces --------- */
/* input perturbation (main.py:148-150): out = z + sigma * eps, eps ~ N(0,1) from Philox4x32-10
 * keyed by (seed, offset) when eps == NULL, else the supplied eps tensor */
int dpi_noise_axpy(const float* z, const float* eps, float* out, int64_t n, float sigma,
                   uint64_t seed, uint64_t offset, int round_tf32, void* stream);
/* same with the Philox offset (= iteration index, counter_dev[0]) and an additive seed (counter_dev[1])
 * read from device memory, for CUDA-graph replay */
int dpi_noise_axpy_dev(const float* z, float* out, int64_t n, float sigma, uint64_t seed,
                       const uint64_t* counter_dev, int round_tf32, void* stream);
/* `--data_forgetting_factor F` (main.py:86-97 setup, main.py:153-155 in the loop): while the iteration index
 * counter_dev[0] is below F,  zin[v][c] += weights_dev[iteration] * data[v][c % Cd]  for the C logical channels of the
 * channels-last network input; a no-op from iteration F on (the launch stays in the replayed graph). */
int dpi_add_data_dev(float* zin, int64_t ld, int C, int64_t nvox, const float* data, int64_t data_ld, int Cd,
                     const float* weights_dev, int F, const uint64_t* counter_dev, int round_tf32, void* stream);
/* depth-wise "same" FIR along the middle axis of a contiguous [outer][T][inner] tensor, zero beyond the ends:
 * y[o][t][i] = sum_m taps[m] * x[o][t + ntaps/2 - m][i]  -  ConvolveKernel_1d / LowPassButterworth of
 * utils/processing.py:34-79 as applied to the input noise once per patch (main.py:66-84); ntaps odd, y != x */
int dpi_fir_axis(const float* x, float* y, int64_t outer, int64_t T, int64_t inner, const float* taps_dev, int ntaps,
                 void* stream);
int dpi_fill_normal(float* out, int64_t n, float mean, float std, uint64_t seed, uint64_t offset,
                    void* stream);
/* end-of-iteration bookkeeping on the device (main.py:165-182): appends {loss,snr,pcorr,lr} to
 * history[counter], keeps the best output (loss <= loss_min, or iteration 0) in `best`, then
 * increments counter_dev[0] and the Adam step hyper_dev[1].  best_state = {loss_min, copy_flag}. */
int dpi_iteration_end(const double* scalars, double* hyper_dev, uint64_t* counter_dev, double* history,
                      int64_t max_iters, double* best_state, const float* out, float* best, int64_t n,
                      void* stream);

int64_t dpi_loss_workspace_bytes(void);
/* masked sampling-operator loss (main.py:161), its gradient, and the sums behind u.snr/u.pcorr
 * (utils/metrics.py:6-44) in one pass.  n = elements of out (pads included, they are 0),
 * n_logical = elements the reference averages over.  scalars_out (device, 8 doubles):
 *   [0] loss  [1] snr_db  [2] pcorr  [3] nan_flag  [4] sum(img^2)  [5] sum((img-out)^2) */
int dpi_masked_loss(const float* out, const float* img, const float* mask, int64_t n,
                    int64_t n_logical, int kind, float* dout, void* workspace,
                    int64_t workspace_bytes, double* scalars_out, void* stream);

/* torch.optim.Adam.step (main.py:200,213) over one flat parameter buffer */
int dpi_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int64_t step, void* stream);
/* y = a*y + b*x over n floats (n % 4 == 0): sums the flat gradients of the batch rows a rank holds in the
 * shared-network mode (replaces the per-patch networks of main.py:274-295) and pre-scales them by
 * 1 / #rows so that the all-reduce(sum) over ranks yields the gradient of the global-mean loss */
int dpi_axpby(float* y, const float* x, float a, float b, int64_t n, void* stream);
/* same, with lr and step read from device memory (CUDA-graph replay): hyper = {lr, step} */
int dpi_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n,
                      const double* hyper_dev, double beta1, double beta2, double eps,
                      double weight_decay, void* stream);

/* ---------------------------------------------------------------- patches ------------------ */
/* PatchExtractor.extract (utils/patch_extractor.py:299-362) + `* gain` (data.py:82): gathers
 * n_patches = prod((n-p)/s+1) windows in C order from a float64 volume; bit-exact. */
int dpi_patch_extract_f64(const double* vol, const int32_t* vol_shape3, const int32_t* patch_shape3,
                          const int32_t* stride3, double gain, double* patches, void* stream);
/* PatchExtractor.reconstruct (utils/patch_extractor.py:370-428) + `/ gain` (data.py:116):
 * float64 overlap-add in patch order, divide by the hit count, cast to float32, divide by gain in
 * float32; bit-exact with the NumPy loops. */
int dpi_patch_reassemble_f32(const float* patches, const int32_t* vol_shape3,
                             const int32_t* patch_shape3, const int32_t* stride3, float gain,
                             float* vol, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPI_B200_H_ */
