/*
 * diqt.h -- C ABI of libdiqt_b200.so, the sm_100a kernel library under the DiffusionIQT
 * sampling hot path (Imagen.sample -> p_sample_loop -> p_sample -> Unet.forward).
 *
 * The reference has no FFI layer of its own: every arithmetic op on the path is a PyTorch
 * library call inside /root/reference/imagen_pytorch3D.py.  Each entry point below names the
 * reference call site(s) it replaces.  The Python shells in diffusioniqt_b200/ (same class
 * names, constructor / forward / sample signatures and state_dict keys as the reference) bind
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only: raw device pointers, ints, floats, a cudaStream_t passed as void*.
 *   - every call is asynchronous on `stream`, never allocates, never synchronises, and is legal
 *     inside CUDA-graph stream capture (diqt_conv_plan_create is the exception: call it once,
 *     outside capture).
 *   - return 0 on success, negative DIQT_E* on failure; diqt_last_error() gives the text
 *     (thread-local).
 *   - activations are channels-last volumes  [n][d0][d1][d2][c]  with a row pitch `ld`
 *     (elements between consecutive voxels, >= c) so that the skip concat
 *     (imagen_pytorch3D.py:1653) is a view, not a copy.  (d0,d1,d2) are the reference's
 *     tensor dims (2,3,4).
 *   - dtype: DIQT_F32 (exact mode, CUDA-core fp32) or DIQT_BF16 (tcgen05 tensor cores, fp32
 *     accumulation in TMEM; statistics, gates, sampler state stay fp32).
 */
#ifndef DIQT_H_
#define DIQT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIQT_ABI_VERSION 1

enum { DIQT_F32 = 0, DIQT_BF16 = 1 };

enum {
  DIQT_OK = 0,
  DIQT_EINVAL = -1,      /* bad argument / unsupported shape            */
  DIQT_ECUDA = -2,       /* a CUDA runtime / driver call failed         */
  DIQT_EUNSUPPORTED = -3 /* valid request this build has no kernel for  */
};

/* conv modes */
enum {
  DIQT_CONV_K3 = 0,   /* 3x3x3, stride 1, zero padding 1:  Block.project, imagen_pytorch3D.py:550-553 */
  DIQT_CONV_K1 = 1,   /* 1x1x1: ResnetBlock.res_conv :597, last-level post_downsample :1388            */
  DIQT_CONV_DOWN = 2, /* pixel-unshuffle(2) + 1x1x1: Downsample :489-496 (a 2x2x2 stride-2 conv)      */
  DIQT_CONV_UP = 3    /* 1x1x1 + Mish + PixelShuffle3D(2): PixelShuffleUpsample :459-487, :416-439    */
};

/* conv kernel families */
enum {
  DIQT_IMPL_AUTO = 0,
  DIQT_IMPL_SIMT = 1, /* CUDA-core implicit GEMM, fp32 accumulate; any channel count; both dtypes */
  DIQT_IMPL_TC = 2,   /* tcgen05 + TMA implicit GEMM, bf16 only, C_in % 64 == 0, C_out % 64 == 0   */
  DIQT_IMPL_ZM = 3    /* tcgen05 "z-march" 3x3x3 64->64 bf16: haloed planes reused by 9 taps, dz taps stacked to N=192 */
};

int diqt_abi_version(void);
const char* diqt_last_error(void);
/* number of kernels this library has launched from the calling process (for bench.py's gpu_launches) */
uint64_t diqt_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Convolutions (implicit GEMM).  Replaces nn.Conv3d at imagen_pytorch3D.py:467, 495, 551-553,
 * 597, 1388 and the einops / PixelShuffle3D data movement around them (:416-439, :493).
 *
 * Weights are repacked once by diqt_conv_pack_weights from the reference's
 * (C_out, C_in, k, k, k) fp32 tensor into the layout the chosen kernel family wants.
 * ------------------------------------------------------------------------------------------ */
/* keep the z-march conv on single CTAs even where the CTA-pair kernel (tcgen05.mma.cta_group::2) applies: parity tests, A/B */
#define DIQT_CONV_FLAG_NO_CTA_PAIR 1

typedef struct diqt_conv_desc {
  int32_t mode;          /* DIQT_CONV_*                                                     */
  int32_t dtype;         /* DIQT_F32 / DIQT_BF16 (activations in and out)                   */
  int32_t impl;          /* DIQT_IMPL_*                                                     */
  int32_t n;             /* volumes                                                         */
  int32_t d0, d1, d2;    /* INPUT spatial dims                                              */
  int32_t c_in, ld_in;   /* input channels / row pitch                                      */
  int32_t c_out, ld_out; /* GEMM N (for DIQT_CONV_UP: 8x the channels stored) / out pitch   */
  int32_t flags;         /* DIQT_CONV_FLAG_* (0 = defaults)                                 */
} diqt_conv_desc;

/* which kernel family DIQT_IMPL_AUTO resolves to for this shape (DIQT_IMPL_SIMT or DIQT_IMPL_TC) */
int diqt_conv_resolved_impl(const diqt_conv_desc* d, int* impl);
/* bytes needed for the packed weight buffer of this conv */
int diqt_conv_packed_bytes(const diqt_conv_desc* d, size_t* bytes);
/* w: device fp32 (C_out, C_in, k,k,k) exactly as in the reference state_dict; bias: device fp32 (C_out) or
 * NULL.  packed_w: device buffer of diqt_conv_packed_bytes; packed_bias: device fp32 [c_out] in GEMM column
 * order (DIQT_CONV_UP re-orders output channels to [sub-position][channel]). */
int diqt_conv_pack(const diqt_conv_desc* d, const float* w, const float* bias, void* packed_w, float* packed_bias,
                   void* stream);

/* A plan owns the TMA descriptors of one conv call site (input / output pointers are baked in). */
typedef struct diqt_conv_plan diqt_conv_plan;
int diqt_conv_plan_create(const diqt_conv_desc* d, const void* in, void* out, const void* packed_w,
                          const float* packed_bias, diqt_conv_plan** plan);
/* Ask the conv to also emit per-block channel (sum, sumsq) of the output it STORES (bf16-rounded), in the
 * layout of diqt_channel_stats: partial[n][*nblk][c_out][2].  *nblk = 0 means this plan cannot fuse them
 * (SIMT family, DIQT_CONV_UP, tiny volumes): run diqt_channel_stats on the output instead. */
int diqt_conv_plan_set_stats(diqt_conv_plan* plan, float* partial, int* nblk);
/* Small volumes (fewer 128-voxel output tiles than half the SMs) run the per-tap family with split-K: *bytes > 0 says the plan wants a
 * workspace of that size (256-byte aligned, its first bytes zero before the first run, shareable between plans that never run
 * concurrently); hand it over BEFORE diqt_conv_plan_set_stats*.  Without a workspace the plan runs unsplit (same results, slower). */
int diqt_conv_plan_workspace_bytes(const diqt_conv_plan* plan, size_t* bytes);
int diqt_conv_plan_set_workspace(diqt_conv_plan* plan, void* workspace, size_t bytes);
void diqt_conv_plan_destroy(diqt_conv_plan* plan);
int diqt_conv_run(const diqt_conv_plan* plan, void* stream);

/* ------------------------------------------------------------------------------------------
 * Channel statistics, GroupNorm + FiLM + Mish, SE gate, residual.
 * Replaces nn.GroupNorm :546, the FiLM line :559-561, nn.Mish :547, SE3D :617-632 and the
 * residual add :612 of Block / ResnetBlock.
 * ------------------------------------------------------------------------------------------ */
/* Sub-volume ("boundary") geometry, taken by the per-volume kernels as (sub_f, sub_h): sub_f <= 1 means n independent volumes of
 * `voxels` rows each.  sub_f = f > 1 is the reference's boundary mode (boundary_pad, imagen_pytorch3D.py:37-46, eval_config.yaml:26):
 * the n = f^3 sub-volumes of side sub_h are stored MERGED as one volume of side f*sub_h (so convolutions see their neighbours for
 * free), sub-volume b = zb + f*yb + f*f*xb being block (zb, yb, xb) along (d0, d1, d2) (utils_mine.py:25-67); statistics, FiLM and
 * gates stay per sub-volume. */
/* per-block partial channel sums: partial[n][nblk][c][2] (sum, sum of squares), fp32; block b of volume i
 * covers voxels [b*ceil(voxels/nblk), ...).  Reduced in a fixed order -> bitwise reproducible. */
int diqt_channel_stats(const void* x, int dtype, int n, int64_t voxels, int c, int ld, int nblk,
                       float* partial, int sub_f, int sub_h, void* stream);
/* Fold GroupNorm(groups, c, eps) (+ optional FiLM) into a per-(n,c) affine  y = a*x + b.
 * film: NULL or fp32 rows [..][film_ld] holding (scale[c], shift[c]) at film[row][0..2c);
 * row = (film_row ? *film_row : 0) + i * film_row_stride_n   (film_row is a DEVICE pointer so a
 * captured graph can follow the sampler step). */
int diqt_gn_finalize(const float* partial, int n, int nblk, int64_t voxels, int c, int groups, float eps,
                     const float* gamma, const float* beta, const float* film, int film_ld,
                     const int32_t* film_row, int film_row_stride_n, float* a, float* b, void* stream);
/* y = mish(a[n][c] * x + b[n][c]) */
int diqt_affine_mish(const void* x, int ld_x, void* y, int ld_y, int dtype, int n, int64_t voxels, int c,
                     const float* a, const float* b, int nblk, int sub_f, int sub_h, void* stream);
/* gate[n][c] = sigmoid(W2 . relu(W1 . mean[n][:])), mean from the channel partial sums;
 * W1: (hidden, c)  W2: (c, hidden), bias-free */
int diqt_se_gate(const float* partial, int n, int nblk, int64_t voxels, int c, int hidden, const float* w1,
                 const float* w2, float* gate, void* stream);
/* out = h * gate[n][c] + res  (gate may be NULL -> 1); if partial != NULL also writes per-block channel
 * stats of out (the input of the next GroupNorm) */
int diqt_scale_residual(const void* h, int ld_h, const void* res, int ld_res, void* out, int ld_out, int dtype,
                        int n, int64_t voxels, int c, const float* gate, int nblk, float* partial, int sub_f, int sub_h,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Grouped statistics: the same GroupNorm / FiLM / Mish / SE / residual arithmetic without the two
 * single-CTA finalisation kernels (diqt_gn_finalize, diqt_se_gate).  A producer given a "sink" also reduces its partial rows in
 * groups of consecutive rows (the last CTA of a group to finish sums that group, fixed order) into
 * group[n][ngroups][c][2], ngroups <= 16; the consumer folds the finalisation into its own prologue.
 * tickets: DEVICE uint32 [n * ngroups] (convs: [ngroups]), zero before first use, self-resetting.
 * Single-launch semantics are identical to the un-grouped calls; plain volumes only (no sub-volume geometry).
 * ------------------------------------------------------------------------------------------ */
/* ngroups a producer with `nblk` partial rows and `rows_per_cta` rows per CTA (conv z-march: 2, everything else: 1) will write */
int diqt_stats_groups(int nblk, int rows_per_cta, int* ngroups);
/* like diqt_conv_plan_set_stats, plus the grouped sink; *ngroups = 0 (and *nblk = 0) if the plan cannot fuse statistics */
int diqt_conv_plan_set_stats_g(diqt_conv_plan* plan, float* partial, float* group, uint32_t* tickets, int* nblk, int* ngroups);
int diqt_channel_stats_g(const void* x, int dtype, int n, int64_t voxels, int c, int ld, int nblk, float* partial,
                         float* group, uint32_t* tickets, void* stream);
/* Block.forward is GroupNorm -> FiLM -> Mish -> Conv3d (imagen_pytorch3D.py:555-565).  For plans of the z-march family
 * (diqt_conv_gn_fusable(desc) != 0: 3x3x3 bf16, c_in <= 256, n <= 2) the first three steps can ride on the conv's load path: the conv then
 * reads the RAW tensor x (not mish(GN(x))) and normalises every input plane in shared memory between the TMA and the tensor core, with
 * the statistics of x taken from `group` exactly as diqt_gn_mish_g does (bit-identical operands, one tensor read + write and one launch
 * less).  diqt_conv_plan_set_film adds / updates the FiLM rows (arguments as in diqt_gn_finalize); call it before diqt_conv_run
 * whenever the table pointer or the row selector changes. */
int diqt_conv_gn_fusable(const diqt_conv_desc* desc);
int diqt_conv_plan_set_gn(diqt_conv_plan* plan, const float* group, int ngroups, int64_t voxels, int groups, float eps,
                          const float* gamma, const float* beta);
int diqt_conv_plan_set_film(diqt_conv_plan* plan, const float* film, int film_ld, const int32_t* film_row, int film_row_stride_n);
/* Same fusion for any batch size / any z-march width: the conv applies y = mish(a * x + b) with the per-(volume, channel) affine
 * a, b ([n][c_in] fp32) that diqt_gn_finalize wrote in the launch before it (GroupNorm and FiLM already folded in). */
int diqt_conv_plan_set_gn_affine(diqt_conv_plan* plan, const float* a, const float* b);
/* y = mish(GroupNorm(groups, eps, gamma, beta)(x) [* (scale + 1) + shift]) with the statistics of x taken from `group`
 * (nn.GroupNorm :546, FiLM :559-561, nn.Mish :547); film arguments as in diqt_gn_finalize */
int diqt_gn_mish_g(const void* x, int ld_x, void* y, int ld_y, int dtype, int n, int64_t voxels, int c, const float* group,
                   int ngroups, int groups, float eps, const float* gamma, const float* beta, const float* film, int film_ld,
                   const int32_t* film_row, int film_row_stride_n, int nblk, void* stream);
/* out = h * SE(h) + res (SE3D :617-632 from the grouped statistics of h; se_group NULL = no gate), plus the statistics of out:
 * partial rows [n][nblk][c][2] and, if group_out != NULL, their grouped reduction */
int diqt_scale_residual_g(const void* h, int ld_h, const void* res, int ld_res, void* out, int ld_out, int dtype, int n,
                          int64_t voxels, int c, const float* se_group, int se_ngroups, int hidden, const float* w1,
                          const float* w2, int nblk, float* partial, float* group_out, uint32_t* tickets, void* stream);

/* dst[r][0..c) = src[r][0..c) * scale over `rows` pitched rows: the scaled skip connection
 * (scale_skip_connection, :1346, :1653) when it cannot be a pure view */
int diqt_scale_copy(const void* src, int ld_src, void* dst, int ld_dst, int dtype, int64_t rows, int c, float scale,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention blocks (attend_at_enc / attend_at_middle, imagen_pytorch3D.py:1392-1403, 1418-1430, 1610-1622, 1635-1646):
 * LinearAttention :926-1016, SoftMaxAttention :1018-1106, ChanFeedForward :1108-1116, ViT3D :871-910.  The 1x1x1 convolutions
 * inside them go through diqt_conv_* (DIQT_CONV_K1); these are the ops in between.  All take channels-last rows [row][ld].
 * The reference runs attention on the f^3 sub-volumes merged into one volume (utils_mine.py:44-67): (x_sub_f, x_sub_h) with
 * x_sub_f > 1 says "this operand is laid out as f^3 separate sub-volumes of side h" (or, for diqt_chan_layernorm, "x is merged
 * while out / res are separate") and the kernel translates rows; 0 = no translation.
 * act codes: 0 none, 1 Mish, 2 GELU (erf).
 * ------------------------------------------------------------------------------------------ */
/* out[r] = LayerNorm_c(act(x[r'])) * g (+ beta) (+ res1[r]) (+ res2[r]); biased variance, eps inside the sqrt (LayerNorm :361-382 with
 * dim=-4, nn.LayerNorm :725, :731, :898).  x_sub_f > 1: x is in merged order, out / res1 / res2 in sub-volume order. */
int diqt_chan_layernorm(const void* x, int ld_x, void* out, int ld_out, int dtype, int64_t rows, int c, const float* g,
                        const float* beta, float eps, int pre_act, const void* res1, int ld_res1, const void* res2, int ld_res2,
                        int x_sub_f, int x_sub_h, void* stream);
/* out = act(a) (+ b) (+ c2), row-wise with pitches: residual adds :1148-1149, :1622, :756-757 and stand-alone activations */
int diqt_rows_combine(const void* a, int ld_a, int act, const void* b, int ld_b, const void* c2, int ld_c, void* out, int ld_out,
                      int dtype, int64_t rows, int c, void* stream);
/* depthwise patch^3 / stride patch conv (Patchify :919, PatchEmbedding :847): tokens[(tz*g+ty)*g+tx][ch] over the merged volume of
 * side g*patch.  w: fp32 [patch^3][c] (tap = (dz*patch+dy)*patch+dx), bias fp32 [c] or NULL.  x_sub_f > 1: x is in sub-volume order. */
int diqt_dw_patchify(const void* x, int ld_x, void* tokens, int ld_t, int dtype, int grid_dim, int patch, int c, const float* w,
                     const float* bias, int x_sub_f, int x_sub_h, void* stream);
/* depthwise 3x3x3, padding 1, one volume (d0,d1,d2) (to_q/k/v.2 :963-975, reconstruct :955, depth_conv :787).  w: fp32 [27][c] */
int diqt_dw_conv3(const void* x, int ld_x, void* out, int ld_out, int dtype, int d0, int d1, int d2, int c, const float* w,
                  const float* bias, void* stream);
/* nn.Upsample(scale_factor=factor, mode='trilinear', align_corners=True) (:900, :954) of a token volume g^3 -> (g*factor)^3 */
int diqt_upsample_trilinear(const void* tokens, int ld_t, void* out, int ld_out, int dtype, int grid_dim, int factor, int c,
                            void* stream);
/* LinearAttention core (:1001-1011): out = act( (softmax_d(q) * scale) (softmax_n(k)^T v) ) per head; q, k, v: [tokens][ld_qkv] with
 * channel = head*dim_head + d.  col_stat: fp32 scratch [heads*dim_head][2]; partial: fp32 scratch
 * [chunks][heads][dim_head][dim_head] with chunks from diqt_linear_attention_chunks.  dim_head in {16, 32, 64}. */
int diqt_linear_attention_chunks(int tokens, int* chunks);
int diqt_linear_attention(const void* q, const void* k, const void* v, int ld_qkv, void* out, int ld_out, int dtype, int tokens,
                          int heads, int dim_head, float scale, int act, float* col_stat, float* partial, void* stream);
/* SoftMaxAttention / MultiHeadAttention core (:1087-1100, :826-836): out = act( softmax_k(q k^T * scale) v ) per head, online softmax */
int diqt_softmax_attention(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v, void* out, int ld_out,
                           int dtype, int tokens, int heads, int dim_head, float scale, int act, void* stream);

/* The same product on the tensor cores (csrc/attn_tc.cu): tcgen05.mma for Q K^T and P V with fp32 accumulators in TMEM, two passes over
 * the key tiles (row max / sum, then probabilities), nothing N x N in memory.  bf16, dim_head = 64, pitches multiples of 8, 16-byte aligned
 * pointers; q / k / v may be column blocks of one [tokens][3 * heads * 64] buffer.  `workspace`: diqt_attn_tc_workspace_bytes() bytes of
 * device memory, ZERO-initialised once by the caller (V^T padded to a multiple of 128 tokens: read by the two-pass kernel only; the
 * single-pass kernel takes V as an MN-major operand where it lies).  A plan bakes the pointers
 * (TMA descriptors); create it once outside stream capture. */
typedef struct diqt_attn_plan diqt_attn_plan;
int diqt_attn_tc_supported(int dtype, int dim_head, int ld_q, int ld_k, int ld_v, int ld_out);
int diqt_attn_tc_workspace_bytes(int tokens, int heads, size_t* bytes);
int diqt_attn_tc_plan_create(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v, void* out, int ld_out, int tokens,
                             int heads, float scale, int act, void* workspace, diqt_attn_plan** plan);
void diqt_attn_tc_plan_destroy(diqt_attn_plan* plan);
int diqt_attn_tc_run(const diqt_attn_plan* plan, void* stream);

/* LinearAttention core (:1001-1011) on the tensor cores (csrc/linattn_tc.cu): k^T v and q ctx as tcgen05.mma (the k / v boxes are used
 * as MN-major operands exactly as TMA lands them: no transposed copies), q, k, v read once, partial contexts of the token chunks summed
 * in a fixed order.  bf16, dim_head = 64, an even number of heads (<= 32), pitches multiples of 8, 16-byte aligned pointers; q / k / v are
 * column blocks of [tokens][ld_qkv] buffers.  `workspace`: diqt_linattn_tc_workspace_bytes() bytes, 256-byte aligned, no initialisation
 * needed.  A plan bakes the pointers (TMA descriptors); create it once outside stream capture.  act: 0 none, 1 Mish (:1011). */
typedef struct diqt_linattn_plan diqt_linattn_plan;
int diqt_linattn_tc_supported(int dtype, int dim_head, int heads, int ld_qkv, int ld_out);
int diqt_linattn_tc_workspace_bytes(int tokens, int heads, size_t* bytes);
int diqt_linattn_tc_plan_create(const void* q, const void* k, const void* v, int ld_qkv, void* out, int ld_out, int tokens, int heads,
                                float scale, int act, void* workspace, diqt_linattn_plan** plan);
void diqt_linattn_tc_plan_destroy(diqt_linattn_plan* plan);
int diqt_linattn_tc_run(const diqt_linattn_plan* plan, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step (csrc/backward.cu): the reverse pass of Unet.forward for loss.backward() in
 * Imagen.forward / p_losses (imagen_pytorch3D.py:2277-2387) and the optimizer update of
 * ImagenTrainer.update (trainer.py).  Data gradients of convolutions run through diqt_conv_*
 * with the flipped, transposed weights.
 * ------------------------------------------------------------------------------------------ */
/* partial[n][blk][c][2] = (sum_v t, sum_v t * x) with t = dz * mish'(a[n][c] x + b[n][c]) (mode 1: reverse of GroupNorm -> FiLM -> Mish,
 * :546-563) or t = dz (mode 0: reverse of the SE gate x * gate, :630).  a, b: the forward pass's folded affine (diqt_gn_finalize). */
int diqt_bwd_reduce(const void* x, int ld_x, const void* dz, int ld_dz, int dtype, int n, int64_t voxels, int c, const float* a,
                    const float* b, int mode, int nblk, float* partial, void* stream);
/* out = c1[n][c] * t + c2[n][c] * x + c3[n][c] (+ acc): the input gradient of GroupNorm -> FiLM -> Mish (mode 1) or of the SE join
 * (mode 0); acc adds the gradient arriving over the residual branch (:612).  c2, c3, acc, x (mode 0 without c2) may be NULL. */
int diqt_bwd_apply(const void* x, int ld_x, const void* dz, int ld_dz, const void* acc, int ld_acc, void* out, int ld_out, int dtype,
                   int n, int64_t voxels, int c, const float* a, const float* b, const float* c1, const float* c2, const float* c3,
                   int mode, int nblk, void* stream);
/* The (n, c)-sized step between the two: from the forward statistics (fwd_partial[n][nblk_f][c][2] of diqt_channel_stats) and
 * bwd_partial[n][nblk_b][c][2] of diqt_bwd_reduce (mode 1) to the coefficients c1, c2, c3 [n][c] of diqt_bwd_apply, per-volume rows
 * d gamma / d beta [n][c] (the caller adds them over n) and (film != NULL) d (scale | shift) [n][2c]; fp64 inside, one CTA per volume,
 * fixed summation order. */
int diqt_gn_bwd_finalize(const float* fwd_partial, int nblk_f, const float* bwd_partial, int nblk_b, int n, int64_t voxels, int c,
                         int groups, float eps, const float* gamma, const float* beta, const float* film, float* c1, float* c2, float* c3,
                         float* dgamma, float* dbeta, float* dfilm, void* stream);
/* Reverse of the SE gate MLP (:617-632) between the plain-mode reduce and apply: from the forward statistics of h, the reduce's
 * sum_v d_out * h and the forward gate to c3[n][c] = d mean / V and per-volume weight gradients dw1[n][hidden][c], dw2[n][c][hidden]
 * (the caller adds them over n); one CTA per volume. */
int diqt_se_bwd(const float* fwd_partial, int nblk_f, const float* bwd_partial, int nblk_b, int n, int64_t voxels, int c, int hidden,
                const float* w1, const float* w2, const float* gate, float* c3, float* dw1, float* dw2, void* stream);
/* dw[c_out][c_in][taps] (the layout of nn.Conv3d.weight, fp32) = sum over voxels of dy[v][c_out] * x[v + tap][c_in]; taps 27: 3x3x3 with
 * padding 1 (:550), taps 1: 1x1x1 (:597, :1388, :1477, and the pixel (un)shuffle convs on rearranged tensors).  Any channel counts,
 * fp32 accumulation, per-chunk partials summed in a fixed order.  workspace: diqt_conv_wgrad_workspace_bytes(). */
int diqt_conv_wgrad_workspace_bytes(int n, int d0, int d1, int d2, int c_in, int c_out, int taps, size_t* bytes);
/* impl: DIQT_IMPL_AUTO / _SIMT / _TC.  TC (csrc/wgrad_tc.cu): bf16, channel counts multiples of 64: tcgen05.mma over the voxels with the
 * dy box and the tap-shifted x boxes as MN-major operands exactly as TMA lands them (out-of-volume rows = zero padding), fp32
 * accumulators in TMEM.  AUTO picks it whenever the shape allows. */
int diqt_conv_wgrad_resolved_impl(int dtype, int c_in, int c_out, int ld_x, int ld_dy, int taps, int impl, int* resolved);
int diqt_conv_wgrad(const void* x, int ld_x, const void* dy, int ld_dy, int dtype, int n, int d0, int d1, int d2, int c_in, int c_out,
                    int taps, int impl, float* dw, float* workspace, void* stream);
/* losses = reduce(loss_fn(pred, target, 'none'), 'b ... -> b', 'mean') (:2355-2357) and d loss / d pred in one pass.  kind 0 l1, 1 l2,
 * 2 smooth-l1; clamp_lo: pred.clamp_(min = lo) first (x_start objective, :2353); sample_weight[n] = p2 weight / (batch * count);
 * loss_partial[n][nblk]: unweighted partial sums. */
int diqt_loss_grad(const float* pred, const float* target, int n, int64_t count, int kind, int clamp_lo, float lo,
                   const float* sample_weight, float* dpred, float* loss_partial, int nblk, void* stream);
/* torch.optim.Adam step (no amsgrad) on one fp32 parameter tensor, gradient pre-multiplied by grad_scale; ema != NULL also updates
 * the exponential moving average ema = ema * ema_decay + p * (1 - ema_decay) (ImagenTrainer.update). */
int diqt_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, float grad_scale, float* ema, float ema_decay, void* stream);

/* ------------------------------------------------------------------------------------------
 * Network ends.
 * ------------------------------------------------------------------------------------------ */
/* init_conv (:1291, :1576): 3x3x3 conv over up to 8 single-channel fp32 planes (the channel
 * concat of x_t, low-res patch, ...) -> channels-last activations.  planes[i] points at a
 * (n, d0, d1, d2) fp32 volume with batch stride plane_stride[i] elements.
 * w: fp32 packed [27][c_in][c_out]; */
int diqt_init_conv(const float* const* planes, const int64_t* plane_stride, int c_in, const float* w_packed,
                   const float* bias, void* out, int ld_out, int dtype, int n, int d0, int d1, int d2, int c_out,
                   int sub_f, int sub_h, void* stream);
int diqt_init_conv_pack(const float* w, int c_out, int c_in, float* packed, void* stream);
/* The same init_conv on the tensor cores (bf16 mode, 27 * c_in <= 64, plain volumes): col[row][k = tap * c_in + ci] = plane_ci at the
 * tap-shifted voxel (tap = (kz*3+ky)*3+kx, zero outside the volume and for k >= 27 * c_in), bf16 [n*d0*d1*d2][64]; followed by a
 * DIQT_CONV_K1 convolution 64 -> c_out whose weight is W[co][ci][tap] re-ordered to [co][tap * c_in + ci] and zero-padded to 64. */
/* CrossEmbedLayer as init conv (:661-686, init_cross_embed=True, the constructor default): one call per kernel size k (odd), writing
 * the channel slice [co_off, co_off + nco) of the channels-last output.  w: fp32 [k^3][c_in][nco] (tap = (kz*k+ky)*k+kx), zero padding
 * (k-1)/2.  Plain volumes (the reference's boundary mode cannot run with a cross-embed init conv: boundary_pad + padded convs). */
int diqt_init_conv_k(const float* const* planes, const int64_t* plane_stride, int c_in, int k, const float* w, const float* bias, void* out,
                     int ld_out, int co_off, int nco, int dtype, int n, int d0, int d1, int d2, void* stream);
int diqt_init_im2col(const float* const* planes, const int64_t* plane_stride, int c_in, void* col, int n, int d0, int d1, int d2,
                     void* stream);
/* init_conv (:1291) as ONE tcgen05 kernel (bf16 output): the im2col rows (K = 27 * c_in <= 64 columns) are built in shared memory, never
 * in global memory; w_packed: bf16 [c_out][64] with column k = tap * c_in + ci, zero padded, 16-byte chunks of every row XOR-swizzled by
 * (row & 7); c_out 64 or 128; d2 a multiple of 16; the input planes 16-byte aligned with strides that are multiples of 4 elements (the
 * kernel copies them with 16-byte cp.async).  partial (may be NULL): channel statistics of the output, one row per CTA,
 * partial[n][*nblk][c_out][2] with *nblk from diqt_init_conv_tc_blocks; group / tickets: optional grouped sink as in diqt_channel_stats_g. */
int diqt_init_conv_tc_supported(int c_in, int c_out, int d1, int d2);
int diqt_init_conv_tc_blocks(int n, int d0, int d1, int* nblk);
int diqt_init_conv_tc(const float* const* planes, const int64_t* plane_stride, int c_in, const void* w_packed, const float* bias, void* out,
                      int ld_out, int n, int d0, int d1, int d2, int c_out, float* partial, float* group, uint32_t* tickets, void* stream);

/* final_conv (:1477, 1x1x1, c -> c_out<=4) producing the fp32 NCDHW prediction, optionally fused
 * with the DDPM update of Imagen.p_mean_variance / p_sample / q_posterior (:1976-2056, :290-309):
 *   x0 = clamp(pred);  mean = an*(x_t*(1-c)/al + c*x0);  x_next = mean + ns*noise
 * sched: DEVICE fp32 table [steps][8] = (alpha, sigma, c, alpha_next, noise_scale, lo, hi, objective)
 * read at row *step.  step_mode 0: only write pred.  1: fused update (x_t, noise, x_next, x0). */
int diqt_final_conv(const void* x, int ld, int dtype, int n, int64_t voxels, int c, int c_out, const float* w,
                    const float* bias, float* pred, int step_mode, const float* sched, const int32_t* step,
                    const float* x_t, const float* noise, float* x_next, float* x0, int sub_f, int sub_h, void* stream);
/* the same elementwise update on an existing prediction (used with dynamic thresholding) */
int diqt_ddpm_update(const float* pred, const float* sched, const int32_t* step, const float* x_t,
                     const float* noise, float* x_next, float* x0, int64_t count, void* stream);

/* ------------------------------------------------------------------------------------------
 * Elucidated (Karras / Heun) sampler step: ElucidatedImagen.one_unet_sample and
 * preconditioned_network_forward, elucidated_imagen.py:329-358, :468-519.
 * table: DEVICE fp32 [forwards][16], one row per U-Net forward, read at row *step:
 *   0 S_noise, 1 sqrt(sigma_hat^2 - sigma^2), 2 c_in(sigma of this forward), 3 c_skip, 4 c_out,
 *   5 divisor sigma (sigma_hat in the Euler pass, sigma_next in the Heun pass),
 *   6 (sigma_next - sigma_hat), 7 0.5*(sigma_next - sigma_hat), 8 c_in(sigma_next), 9 clamp lo, 10 clamp hi
 * ------------------------------------------------------------------------------------------ */
/* x_hat = x + table[1] * (table[0] * eps);  x_in = table[2] * x_hat     (:476-481 and the c_in scaling of :349) */
int diqt_edm_prepare(const float* x, const float* eps, const float* table, const int32_t* step, float* x_hat, float* x_in,
                     int64_t count, void* stream);
/* final_conv (:1477) fused with one pass of the Heun step.  D(x) = clamp(c_skip*x + c_out*net, lo, hi).
 *   pass 0 (Euler, :488-498): slope = (x_hat - D(x_hat)) / sigma_hat;  state = x_hat + (sigma_next - sigma_hat) * slope;
 *                             next_input = c_in(sigma_next) * state;  x0 = D(x_hat)
 *   pass 1 (Heun,  :502-516): d' = (state - D(state)) / sigma_next;   state = x_hat + 0.5 (sigma_next - sigma_hat) (slope + d');
 *                             x0 = D(state_before) */
int diqt_final_conv_edm(const void* x, int ld, int dtype, int n, int64_t voxels, int c, int c_out, const float* w,
                        const float* bias, int pass, const float* table, const int32_t* step, const float* x_hat,
                        float* slope, float* state, float* x0, float* next_input, int sub_f, int sub_h, void* stream);
/* the same update given an already denoised tensor (dynamic thresholding takes its quantile in torch in between) */
int diqt_edm_update(const float* denoised, int pass, const float* table, const int32_t* step, const float* x_hat, float* slope,
                    float* state, float* next_input, int64_t count, void* stream);

/* x = min(max(x, lo), hi): the clamp after the loop (:2154-2157) */
int diqt_clamp(float* x, int64_t count, float lo, float hi, void* stream);

/* ------------------------------------------------------------------------------------------
 * Time conditioning (:518-533, :1305-1316, :586-589): tiny dense layers, run once per
 * sampler for all steps.
 * ------------------------------------------------------------------------------------------ */
/* out[r][0] = t[r]; out[r][1+j] = sin(2 pi t w_j); out[r][1+half+j] = cos(2 pi t w_j) */
int diqt_fourier_features(const float* t, int rows, const float* w, int half, float* out, void* stream);
/* y[r][o] = act_out( sum_k act_in(x[r][k]) * W[o][k] + b[o] );  act: 0 none, 1 mish */
int diqt_linear(const float* x, int ldx, int rows, int k, const float* w, const float* b, int out_features,
                float* y, int ldy, int act_in, int act_out, void* stream);
/* increments a device step counter (last node of the captured step graph) */
int diqt_advance_step(int32_t* step, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-volume inference around the sampler (data.py:139-202, utils_mine.py:25-67, test_all.py:239-300), fp32 volumes (d0, d1, d2).
 * ------------------------------------------------------------------------------------------ */
/* out[b] = volume[o0:o0+P, o1:o1+P, o2:o2+P] for the n_patches origins (DEVICE int32 [n][3]).  sub_f <= 1: out is (n, 1, P, P, P).
 * sub_f = f > 1: every patch is written as f^3 sub-volumes of side P/f in the order of convertVolume2subVolume (utils_mine.py:25-42):
 * out is (n * f^3, 1, P/f, P/f, P/f), which is what a batch_sample U-Net takes. */
int diqt_gather_patches(const float* volume, int d0, int d1, int d2, const int32_t* origins, int n_patches, int patch, int sub_f,
                        float* out, void* stream);
/* The stitch loop of test_all.py:239-298 for the whole volume at once.  patches: (kept, P^3) in the layout diqt_gather_patches writes;
 * slot_of_grid: DEVICE int32 [g0*g1*g2], the index of grid patch (i, j, k) in `patches` or -1 for a skipped patch (data.py:192-196).
 * Every voxel takes the value of the LAST patch in grid order whose centre crop covers it (margin overlap/2, none at volume faces;
 * `batch_sample` selects the :270-293 face rule; overlap >= patch: no crop), voxels no crop covers keep pred's value.  lowres != NULL:
 * voxels with lowres == min_val are set to min_val (:300). */
int diqt_stitch_patches(const float* patches, const int32_t* slot_of_grid, int g0, int g1, int g2, int stride, int patch, int overlap,
                        int batch_sample, int sub_f, float* pred, int d0, int d1, int d2, const float* lowres, float min_val, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIQT_H_ */
