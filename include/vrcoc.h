/*
 * vrcoc.h — C-ABI of libvrcoc.so: hand-written sm_100a CUDA kernels for the ASY-VRNet hot path
 * (context-cluster block + asymmetric vision<->radar fusion).
 *
 * The reference (GuanRunwei/ASY-VRNet) is pure PyTorch and has no FFI layer; its boundary for this path is the
 * nn.Module surface.  The Python host side (asy-vrnet_b200/vrcoc) mirrors those modules and binds
 * these entry points with ctypes.  Each entry point below cites the reference lines whose eager-op sequence it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator); the library never
 *    allocates, frees or retains memory, keeps no mutable global state and is re-entrant (DataParallel calls it
 *    from one host thread per GPU, autograd from its own thread);
 *  - tensors are dense NCHW; "dim2"/"dim3" are the two spatial axes; the reference names dim2 "w" and dim3 "h"
 *    (backbone/fusion/vr_coc.py:155), so fold_w/proposal_w act on dim2 and fold_h/proposal_h on dim3;
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); launches are
 *    asynchronous, never synchronise, and are CUDA-graph capturable;
 *  - return value: 0 on success, a negative VRCOC_E* code otherwise; vrcoc_last_error() gives the message of the
 *    last failure on the calling thread.  There is no CPU fallback anywhere.
 */
#ifndef VRCOC_H_
#define VRCOC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRCOC_OK 0
#define VRCOC_EINVAL (-1)  /* bad argument (null pointer, unsupported shape/dtype) */
#define VRCOC_ECUDA (-2)   /* CUDA runtime error at launch */
#define VRCOC_ENODEV (-3)  /* not an sm_100 device */

/* storage dtypes (arithmetic is always fp32 on chip; bf16 projections use tcgen05 with fp32 accumulation) */
#define VRCOC_F32 0
#define VRCOC_BF16 1

#define VRCOC_STAT_SLOTS 32

/* epilogue activations */
#define VRCOC_ACT_NONE 0
#define VRCOC_ACT_RELU 1
#define VRCOC_ACT_GELU 2 /* exact erf GELU, nn.GELU() default */
#define VRCOC_ACT_SILU 3
#define VRCOC_ACT_LRELU 4 /* LeakyReLU(0.1), reference normal_conv.py:17 */

const char* vrcoc_version(void);
const char* vrcoc_last_error(void);
/* 1 if the current device can run the kernels (compute capability 10.x), else 0 */
int vrcoc_device_ok(void);

/* ------------------------------------------------------------------------------------------------------------
 * Statistics currency.
 *   channel sums: float [B][C][2] = {sum_hw x, sum_hw x^2}   (ShuffleAttention GN / avg-pool: shuffle_attention.py:57-64,
 *                 eca.py:17, BatchNorm2d batch statistics: normal_conv.py:45)
 *   sample sums : double [B][VRCOC_STAT_SLOTS][2], slot-wise partial {sum_chw x, sum_chw x^2}; the statistic of sample b is
 *                 the sum over its slots (GroupNorm(1,C): vr_coc.py:105-111).  Producers add one partial per CTA into
 *                 slot (cta index mod VRCOC_STAT_SLOTS): a trace of the first version showed ~16 K fp64 atomics on 16
 *                 addresses serialising in L2 (15 us per CTA); spreading them removes the contention.
 * ---------------------------------------------------------------------------------------------------------- */
int vrcoc_channel_sums(const void* x, int dtype, int B, int C, int HW, float* chan_sums, double* sample_sums /*nullable, must be zeroed*/,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution-as-GEMM engine with table-driven prologue and fused epilogue.  One launch computes
 *
 *   z[b,k,p]   = T(src[b, chan_src[k], p])                               (prologue, per logical input channel k)
 *                T(x) = (x*scale + shift) * sigmoid(gate_a*x + gate_c)    [gate only if has_gate]
 *   acc[b,o,q] = sum_{k,ky,kx} W[o,k,ky,kx] * z[b,k, q*stride - pad + (ky,kx)]      (zero outside the map)
 *   y = act(acc*e_scale[o] + e_shift[o]) * post_scale[o] + res[b,o,q];  y = y*f_scale[o] + f_shift[o]
 *   out[b,o,q] = y        (+ optional side outputs: per-sample sums of y, global min/max of y)
 *
 * It replaces, in the reference: Cluster.fc1/fc_v/fc2 and Mlp.fc1/fc2 with the GroupNorm, GELU, layer-scale and
 * residual around them (vr_coc.py:156-157,191,217-223,264-271); PointRecuder convs (vr_coc.py:99-102, incl. the
 * cat([x,pos]) of :582-586 through the two-source input); BaseConv conv+BN+act (normal_conv.py:48-49); the
 * shuffle/attention/ECA/1x1/BN chain of RadarEnhanceByImage (vr_coc.py:331-357) and the radar projection of
 * ImageEnhanceByRadar (vr_coc.py:313).
 *
 * Prologue sources (exactly one of, or none):
 *   gn_sums != NULL  : GroupNorm(1,C) of src0: scale = rstd_b*gamma[k], shift = beta[k] - mean_b*rstd_b*gamma[k]
 *   table   != NULL  : float [B][Cin][4] = {scale, shift, gate_a, gate_c}
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct vrcoc_conv_desc {
  /* geometry */
  int32_t B, H_in, W_in, H_out, W_out;
  int32_t C0, C1;          /* channels of src0 / src1; logical input channels Cin = C0 + C1 */
  int32_t O;               /* output channels */
  int32_t kh, kw, stride, pad;
  /* inputs */
  const void* src0; int32_t src0_dtype; int64_t src0_bstride; /* batch stride in elements (0 broadcasts, e.g. fea_pos) */
  const void* src1; int32_t src1_dtype; int64_t src1_bstride; /* may be NULL when C1 == 0 */
  const int32_t* chan_src; /* nullable [Cin]: logical channel k reads channel chan_src[k] of the virtual concat [src0|src1] */
  /* prologue */
  const double* gn_sums; const float* gn_gamma; const float* gn_beta; float gn_eps; /* GroupNorm(1,C0) on src0 */
  const float* table; int32_t has_gate;
  /* weights: [O][Cin*kh*kw] row-major (== PyTorch conv weight), fp32 or bf16 */
  const void* weight; int32_t weight_dtype;
  /* epilogue (every pointer nullable = identity) */
  const float* e_scale; const float* e_shift; int32_t act; const float* post_scale;
  const void* res; int32_t res_dtype;
  const float* f_scale; const float* f_shift;
  /* outputs: channels [0,O_split) -> out, [O_split,O) -> out2 (out2 may be NULL when O_split == O) */
  void* out; int32_t out_dtype;
  void* out2; int32_t out2_dtype; int32_t O_split;
  double* out_sample_sums; /* nullable [B][VRCOC_STAT_SLOTS][2], accumulated (caller zeroes) */
  uint32_t* out_minmax;    /* nullable [2] = {max bits(y), max ~bits(y)}, y >= 0 required, caller zeroes */
  /* 0 = pick automatically, 1 = force CUDA-core fp32 path, 2 = force tcgen05 bf16 path */
  int32_t engine;
  int32_t dil;     /* dilation (0 or 1 = dense); ASPP branches use 6/12/18 (neck/coc_fpn_dual.py:55-67) */
  int32_t k_order; /* 0: weight[o][c][ky][kx] (PyTorch); 1: weight[o][ky][kx][c] (tap-major: the im2col rows of one k slab
                      are consecutive channels of one tap -> cheap gathers on the tcgen05 path);
                      2: ROW-TAP mode (bf16 tcgen05 TMA kernels only): src0 = vrcoc_im2col_rows(x) = [B][C0 = kw_orig*C][H][W],
                      kh = vertical taps, kw = 1, stride 1, pad = dil*(kh-1)/2 (rows only: H_out = H_in, W_out = W_in),
                      weight[o][ky][C0] (= the tap-major weight of the k x k convolution); C0 % 64 == 0, dil*W % 8 == 0,
                      H*W % 8 == 0, no prologue / second source.  The k slab (ky, c0) is the 64-point x 64-channel TMA box of
                      src0 moved by (ky - kh/2)*dil*W points (whole rows), zero-filled above / below the map. */
  /* GroupNorm(1,C0) FOLDED into a 1x1 projection (bf16 tcgen05 path; Cluster.fc1|fc_v after norm1, vr_coc.py:156-157,265):
   *   W.GN(x) + b  =  rstd_b * ((W diag(gamma)) . x)  -  rstd_b * mean_b * k1[o]  +  k0[o]
   * so the tensor core consumes the RAW activations and the per-sample statistics enter in the epilogue only.  When
   * gn_fold_k1 != NULL: weight is [O][2*C0] bf16 = [hi | lo], the two-term bf16 split of W diag(gamma) (rows >= O_split use
   * the hi half only); e_shift = k0[o] = b[o] + sum_c W[o,c] beta[c]; gn_fold_k1[o] = sum_c (hi + lo)[o,c]; gn_sums gives
   * the statistics; gn_gamma / gn_beta / e_scale must be NULL.  The split keeps the similarity operand `feat` exact to
   * ~2^-17: rounding GN(x) itself to bf16 flips ~0.05 % of the hard assignments and moves the Cluster output by ~2e-2. */
  const float* gn_fold_k1;
} vrcoc_conv_desc;

int vrcoc_conv_fwd(const vrcoc_conv_desc* d, void* stream);
/* 1 when the folded-GroupNorm form of a bf16 1x1 projection (gn_fold_k1 above) is available for [B][C][P] -> O channels
 * (first O_split of them fp32 `feat`); the host side asks before it builds the [hi | lo] weights. */
int vrcoc_gn_fold_supported(int B, int C, int O, int O_split, int P);

/* Backward helpers of the 1x1 projections (training path of vr_coc.py:156-157,191,217-223).
 *   wgrad: dW[o][k] (+)= sum_{b,p} dy[b,o,p] * z[b,k,p],  db[o] (+)= sum dy      (z = prologue(src), as in fwd)
 *   dgrad is vrcoc_conv_fwd itself with the transposed weight. */
int vrcoc_conv1x1_wgrad(const vrcoc_conv_desc* fwd_desc /* src*, tables, geometry of the forward call */,
                        const void* dy, int dy_dtype, float* dW /*[O][Cin] fp32*/, float* db /*nullable [O]*/,
                        float* workspace, int64_t workspace_floats, void* stream);
int64_t vrcoc_conv1x1_wgrad_workspace(const vrcoc_conv_desc* fwd_desc);

/* ------------------------------------------------------------------------------------------------------------
 * Cluster core: everything of Cluster.forward between the projections (vr_coc.py:158-190): head split, region
 * fold, adaptive-pool centre proposal, cosine similarity, sigmoid(beta + alpha*sim), arg-max hard assignment
 * (lowest index wins ties), similarity-weighted aggregation to the centres, dispatch back to the points, unfold.
 * feat/value/out are [B, E*D, H, W].  idx (uint8 [B,E,H,W]) and sim_max (float [B,E,H,W]) are optional outputs
 * (saved for backward / mask-parity checks).  alpha/beta are device scalars (Cluster.sim_alpha / sim_beta).
 * ---------------------------------------------------------------------------------------------------------- */
int vrcoc_cluster_core_fwd(const void* feat, int feat_dtype, const void* value, int value_dtype,
                           void* out, int out_dtype, uint8_t* idx, float* sim_max,
                           const float* alpha, const float* beta,
                           int B, int E, int D, int H, int W,
                           int fold_w, int fold_h, int proposal_w, int proposal_h,
                           int64_t feat_bstride, int64_t value_bstride, int64_t out_bstride /* elements; 0 = dense E*D*H*W */,
                           void* stream);

/* Backward of the above.  dalpha_beta: float [2], accumulated deterministically from
 * `partials` (float [2*n_region_heads] workspace, n_region_heads = B*E*fold_w*fold_h). */
int vrcoc_cluster_core_bwd(const void* feat, int feat_dtype, const void* value, int value_dtype,
                           const void* dout, int dout_dtype, const uint8_t* idx, const float* sim_max,
                           const float* alpha, const float* beta,
                           void* dfeat, int dfeat_dtype, void* dvalue, int dvalue_dtype,
                           float* dalpha_beta, float* partials,
                           int B, int E, int D, int H, int W,
                           int fold_w, int fold_h, int proposal_w, int proposal_h,
                           int64_t feat_bstride, int64_t value_bstride, int64_t dout_bstride, int64_t dfeat_bstride,
                           int64_t dvalue_bstride /* elements; 0 = dense */, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fusion helpers.
 * ---------------------------------------------------------------------------------------------------------- */
/* ShuffleAttention(G) + shuffle_channels(cat) + eca_block prologue table of RadarEnhanceByImage
 * (vr_coc.py:344-350, shuffle_attention.py:48-72, eca.py:16-22).  Step 1 computes the per-(b, image channel)
 * attention parameters and the spatial sums of the attended image; step 2 runs the ECA conv1d over the
 * interleaved channel means and emits the conv prologue table [B][Ci+Cr][4]. */
int vrcoc_sa_gate_sums(const void* image, int dtype, int B, int Ci, int HW, int G /*0 = no ShuffleAttention*/,
                       const float* chan_sums_img /*[B][Ci][2]; NULL: each block reduces its own plane first*/,
                       const float* cweight, const float* cbias, const float* sweight, const float* sbias,
                       const float* gn_weight, const float* gn_bias,
                       float* attn /*[B][Ci][4] = {scale, gate_a, gate_c, mean of attended channel}*/, void* stream);
/* Steps 0 and 1 of RadarEnhanceByImage's statistics in ONE launch (was: channel_sums(image), channel_sums(radar), sa_gate_sums —
 * three dependent launches of ~7 us each in front of every fusion stage): a block per image plane reduces its own plane (the
 * sums vrcoc_sa_gate_sums takes as an argument; chan_sums_img = NULL there does the same) and then the gated sum, a block per
 * radar plane writes chan_sums_radar [B][Cr][2].  Meant for planes a single block streams quickly (HW <= 65536). */
int vrcoc_fusion_stats(const void* image, const void* radar, int dtype, int B, int Ci, int Cr, int HW, int G,
                       const float* cweight, const float* cbias, const float* sweight, const float* sbias,
                       const float* gn_weight, const float* gn_bias, float* attn /*[B][Ci][4]*/,
                       float* chan_sums_radar /*[B][Cr][2]*/, void* stream);
int vrcoc_radar_enh_table(const float* attn /*[B][Ci][4]*/, const float* chan_sums_radar /*[B][Cr][2]*/,
                          const int32_t* chan_src /*[Ci+Cr] logical -> concat channel; image channels already
                                                    include ShuffleAttention's own channel_shuffle*/,
                          const float* eca_weight, int eca_k, int B, int Ci, int Cr, int HW,
                          float* table /*[B][Ci+Cr][4]*/, void* stream);
/* Same table with its rows in the channel order of the virtual concat [image | radar] (row chan_src[k] instead of k): for
 * running the projection on the two sources in memory order with permuted weight columns, W'[:, chan_src[k]] = W[:, k]. */
int vrcoc_radar_enh_table_concat_order(const float* attn, const float* chan_sums_radar, const int32_t* chan_src,
                                       const float* eca_weight, int eca_k, int B, int Ci, int Cr, int HW, float* table,
                                       void* stream);

/* Stand-alone application of a conv prologue (no contraction): out[b,k,p] = T(src[b, chan_src[k], p]) with T taken
 * from desc->gn_* or desc->table exactly as in vrcoc_conv_fwd (weight/epilogue fields are ignored, O must equal
 * C0+C1).  Used for GroupNorm, ShuffleAttention and eca_block when they are called outside a fused chain. */
int vrcoc_table_apply(const vrcoc_conv_desc* d, void* stream);

/* Per-channel affine / activation / residual / affine pass with optional side statistics:
 *   y = act(x*s1[c] + t1[c]) + res;  y = y*s2[c] + t2[c]
 * (BatchNorm2d application in train mode, normal_conv.py:49, vr_coc.py:315,341,356). */
int vrcoc_chan_affine(const void* x, int x_dtype, const void* res, int res_dtype, void* out, int out_dtype,
                      const float* s1, const float* t1, int act, const float* s2, const float* t2,
                      int B, int C, int HW, float* out_chan_sums /*nullable [B][C][2]*/,
                      uint32_t* out_minmax /*nullable*/, void* stream);

/* ImageEnhanceByRadar tail (vr_coc.py:314-315, data_normal :59-67 for a non-negative map):
 *   y = (1 + (k - mn)/(mx - mn)) * image;  out = y*s[c] + t[c]     mn/mx decoded from `minmax` (see out_minmax). */
int vrcoc_img_enh_finish(const void* k, int k_dtype, const void* image, int image_dtype, void* out, int out_dtype,
                         const uint32_t* minmax, const float* s, const float* t, int B, int C, int HW,
                         float* out_chan_sums /*nullable [B][C][2]*/, void* stream);


/* Debug aid: when `buf` is non-NULL the tcgen05 TMA kernel writes 8 u64 nanosecond phase timestamps per CTA into it
 * (start, setup done, first operands landed, last MMA committed, epilogue done, TMEM released); NULL switches it off. */
int vrcoc_debug_set_trace(unsigned long long* buf);

/* Fused channel MLP of a ClusterBlock (reference vr_coc.py:208-228 Mlp inside :264-275), inference, bf16:
 *   out = x + layer_scale * ( W2 . gelu( W1 . GroupNorm1(x) + b1 ) + b2 ),   hidden never leaves the SM (csrc/mlp_fused.cuh).
 * x, out: [B][C][P] bf16; gn_sums: per-sample slot sums of x (vrcoc_channel_sums / a producer's out_sample_sums); w1 [hidden][C],
 * w2 [C][hidden] bf16; b1, b2, layer_scale (may be NULL = 1), gamma, beta fp32; out_sample_sums (may be NULL, else zeroed by the
 * caller) receives the slot sums of `out`.  vrcoc_mlp_fused_supported tells whether a shape is covered (C in {64, 128},
 * hidden % 128 == 0, P % 8 == 0); other shapes use two vrcoc_conv_fwd calls. */
int vrcoc_mlp_fused_supported(int dtype, int C, int hidden, int P);
int vrcoc_mlp_fused_fwd(const void* x, const double* gn_sums, const float* gamma, const float* beta, float eps, const void* w1,
                        const float* b1, const void* w2, const float* b2, const float* layer_scale, void* out,
                        double* out_sample_sums, int B, int C, int hidden, int P, void* stream);

/* Fused token-mixer half of a ClusterBlock (reference vr_coc.py:155-192 Cluster.forward inside :264-267), inference, bf16:
 *   out = x + layer_scale * ( W2 . cluster_core( fc1(GN(x)), fc_v(GN(x)) ) + b2 )
 * in ONE persistent kernel (csrc/token_mixer_fused.cu): one CTA per 16x16 region, feat / value / core output never leave the
 * SM.  GroupNorm is folded into the projection weights exactly as for vrcoc_conv_fwd with gn_fold_k1: w_fold [2*E*D][2*C]
 * bf16 = [hi | lo], k0 / k1 [2*E*D] fp32.  x, out [B][C][H][W] bf16; w2 [C][E*D] bf16; b2, layer_scale (nullable = 1) fp32;
 * gn_sums = slot sums of x; out_sample_sums (nullable, caller zeroes) receives the slot sums of out; idx (uint8 [B][E][H][W])
 * and sim_max (float [B][E][H][W]) are optional outputs as in vrcoc_cluster_core_fwd.  Covered: C in {64, 128}, 4 heads x 32,
 * 16x16 regions (H / fold_w = W / fold_h = 16), 2x2 proposals = stages 1 and 2 of the backbone; vrcoc_token_mixer_supported
 * tells; everything else runs the three-launch path. */
int vrcoc_token_mixer_supported(int dtype, int C, int H, int W, int heads, int head_dim, int fold_w, int fold_h, int proposal_w,
                                int proposal_h);
/* The same fusion without fc2 for the wide stage (C = 320, heads a multiple of 4, head_dim 32, 16x16 regions: stage 3 of the
 * backbone): GN-folded fc1|fc_v -> cluster core in one launch, one CTA per (region, group of four heads); o [B][E*D][H][W] bf16 is
 * the input of the separate fc2 projection (fc2 contracts over all heads). */
int vrcoc_token_mixer_core_supported(int dtype, int C, int H, int W, int heads, int head_dim, int fold_w, int fold_h, int proposal_w,
                                     int proposal_h);
int vrcoc_token_mixer_core_fwd(const void* x, const double* gn_sums, float gn_eps, const void* w_fold, const float* k0, const float* k1,
                               const float* alpha, const float* beta, void* o, uint8_t* idx, float* sim_max, int B, int C, int H, int W,
                               int heads, int head_dim, int fold_w, int fold_h, void* stream);
/* Debug aid: non-NULL `buf` (u64 [grid][4 iterations][16]) makes the fused token-mixer kernel record %globaltimer phase stamps. */
int vrcoc_debug_set_tm_trace(unsigned long long* buf);
int vrcoc_token_mixer_fwd(const void* x, const double* gn_sums, float gn_eps, const void* w_fold, const float* k0, const float* k1,
                          const float* alpha, const float* beta, const void* w2, const float* b2, const float* layer_scale, void* out,
                          double* out_sample_sums, uint8_t* idx, float* sim_max, int B, int C, int H, int W, int heads, int head_dim,
                          int fold_w, int fold_h, void* stream);

/* Debug / A-B switch: on = 0 routes 1x1 projections to the point-major tcgen05 kernels instead of the channel-major one
 * (conv_tc_cm.cuh); results are identical up to fp32 summation order.  Default 1. */
int vrcoc_debug_set_cm(int on);

/* Explicit tap-major im2col: col[b][(ky*kw+kx)*C + c][oy][ox] = x[b][c][oy*s-p+ky*d][ox*s-p+kx*d] (zero outside).  Feeds the
 * TMA-only tcgen05 kernel for k x k convolutions with many channels on small maps (vr_coc.py:99-102,313; coc_fpn_dual.py:55-67). */
int vrcoc_im2col(const void* x, void* col, int dtype, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int dil,
                 void* stream);
/* Horizontal taps only: cols[b][kx*C + c][y][x] = x[b][c][y][x + (kx - kw/2)*dil] (zero outside), kw odd.  Source of the
 * convolution engine's row-tap mode (vrcoc_conv_desc.k_order = 2): the vertical taps are TMA boxes of `cols` shifted by rows, so a
 * 3x3 convolution materialises 3x its input instead of the 9x of vrcoc_im2col (vr_coc.py:99-102,313; coc_fpn_dual.py:55-67). */
int vrcoc_im2col_rows(const void* x, void* cols, int dtype, int B, int C, int H, int W, int kw, int dil, void* stream);

/* Patch embedding (vr_coc.py:83-102 PointRecuder with patch = stride = 4, pad 0, called from :575-587 on cat([x, pos])): a 4x4 / stride-4
 * convolution of C0 (+ C1 from `extra`, [C1][H][W] with extra_bstride 0 = the batch-broadcast position grid, or [B][C1][H][W]) channels
 * into O = 64 channels, bias added; weight [O][C0+C1][4][4] bf16 (PyTorch layout), out [B][O][H/4][W/4] bf16.  out_sample_sums
 * (nullable, [B][VRCOC_STAT_SLOTS][2], accumulated) receives the per-sample sum / sum of squares of the output, as the convolution
 * engine's epilogue does.  HBM-bound gather + a K = 16*(C0+C1) contraction on mma.sync (csrc/patch_embed.cu).
 * _supported: bf16, patch 4, O = 64, C0 + C1 <= 8, W % 64 == 0, (H/4)*(W/4) % 128 == 0. */
int vrcoc_patch_embed_supported(int dtype, int C0, int C1, int H, int W, int O, int patch);
int vrcoc_patch_embed(const void* x, const void* extra, int64_t extra_bstride, const void* weight, const float* bias, void* out,
                      double* out_sample_sums, int dtype, int B, int C0, int C1, int H, int W, int O, int patch, void* stream);

/* Depthwise k x k convolution (k = 3 or 5; weight [C][k][k] in the activation dtype, optional fp32 bias): DWConv.dconv of the
 * decoupled head (backbone/conv_utils/normal_conv.py:26-27, head/decouplehead.py:24-37). */
int vrcoc_dwconv(const void* x, const void* weight, const float* bias, void* out, int dtype, int B, int C, int H, int W, int k,
                 int stride, int pad, void* stream);

/* Bilinear upsample with align_corners=True over `planes` = B*C maps (nn.Upsample inside CoCUpsample,
 * reference neck/coc_fpn_dual.py:19-22). */
int vrcoc_upsample_bilinear(const void* x, void* out, int dtype, int planes, int H, int W, int Ho, int Wo, void* stream);
/* Serving tail of the segmentation half (reference deeplab.py:149-167: F.interpolate(logits, frame size, bilinear,
 * align_corners=True) then arg-max over the classes) as ONE launch: out[b][oy][ox] (uint8) = argmax_c of the upsampled logits
 * x [B][C][H][W], each value rounded to `dtype` first and the lowest class index winning ties — bit-identical to
 * vrcoc_upsample_bilinear followed by an arg-max, without the [B][C][Ho][Wo] logits reaching memory.
 * vrcoc_upsample_argmax_supported returns non-zero (the strip height used) when a shape fits: all classes of a 16-, 8- or 4-row
 * strip of source rows resident in shared memory. */
int vrcoc_upsample_argmax_supported(int C, int H, int W, int Ho, int Wo);
int vrcoc_upsample_argmax(const void* x, uint8_t* out, int dtype, int B, int C, int H, int W, int Ho, int Wo, void* stream);

/* Backward elementwise passes of the projections (training path of ClusterBlock, vr_coc.py:264-271).
 *   gelu_bwd    : out = dy * gelu'(u)                                   (all three tensors share `dtype`)
 *   gn_bwd_sums : out[b][c] = {sum_p dz, sum_p dz*x}                    GroupNorm(1,C) backward statistics
 *   gn_bwd_apply: out = dz*a[b,c] + x*bb[b] + cc[b] (+ extra)           GroupNorm(1,C) input gradient (+ residual grad) */
int vrcoc_gelu_bwd(const void* dy, const void* u, void* out, int dtype, int64_t n, void* stream);
int vrcoc_gn_bwd_sums(const void* dz, const void* x, int dtype, int B, int C, int HW, float* out, void* stream);
int vrcoc_gn_bwd_apply(const void* dz, const void* x, const void* extra /*nullable*/, void* out, int dtype, const float* a,
                       const float* bb, const float* cc, int B, int C, int HW, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training path, small backward kernels (csrc/bwd.cu).
 * ---------------------------------------------------------------------------------------------------------- */
/* GroupNorm(1,C) backward coefficients (vr_coc.py:105-111 inside :265,:270) from s = vrcoc_gn_bwd_sums output [B][C][2] and the
 * forward statistics: a[b][c] = rstd_b*gamma_c, bb[b], cc[b] for vrcoc_gn_bwd_apply (dx = dz*a + x*bb + cc), dgamma[c], dbeta[c].
 * workspace: float [B][C].  Deterministic (fixed reduction order). */
int vrcoc_gn_bwd_coef(const float* s, const double* gn_sums, const float* gamma, float eps, int B, int C, int HW, float* a, float* bb,
                      float* cc, float* dgamma, float* dbeta, float* workspace, void* stream);
/* Gradients of out = res + ls * (W h + b) w.r.t. W, b, ls from G[o][k] = sum dout*h and sum_dy[o] = sum dout (vrcoc_conv1x1_wgrad):
 * dW = ls*G, db = ls*sum_dy, dls = sum_k W*G + b*sum_dy, and wt[k][o] = W[o][k]*ls[o] (the dgrad weight, in W's dtype).
 * bias / ls / db / dls nullable (vr_coc.py:191,222 with the layer scale of :266-271). */
int vrcoc_proj_res_bwd_coef(const float* G, const float* sum_dy, const void* W, int w_dtype, const float* bias, const float* ls, int O,
                            int K, float* dW, float* db, float* dls, void* wt, void* stream);
/* Adjoint of vrcoc_im2col (same tap-major layout): dx[b][c][y][x] = sum of the dcol entries that were copies of x[b][c][y][x];
 * with dcol = W^T . dy this is the input gradient of a k x k convolution (PointRecuder / BaseConv, vr_coc.py:99-102,313). */
int vrcoc_col2im(const void* dcol, void* dx, int dtype, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int dil,
                 void* stream);
/* Per-channel affine (+ activation) backward = BatchNorm2d backward (normal_conv.py:45-49; vr_coc.py:315,341,356):
 *   g = dy * act'(.)   relu / lrelu: from y_act = forward output of the activation (NULL = no activation);
 *                      silu: from z = u*z_scale[c] + z_shift[c], the recomputed pre-activation
 *   sums : out[b][c] = { sum_p g, sum_p g*u }
 *   apply: out = g*ca[c] + u*cb[c] + cd[c] (+ extra)          (cb / cd / extra / z_* nullable) */
int vrcoc_chan_bwd_sums(const void* dy, const void* y_act, const void* u, int dtype, int act, const float* z_scale, const float* z_shift,
                        int B, int C, int HW, float* out, void* stream);
int vrcoc_chan_bwd_apply(const void* dy, const void* y_act, const void* u, const void* extra, void* out, int dtype, int act,
                         const float* ca, const float* cb, const float* cd, const float* z_scale, const float* z_shift, int B, int C,
                         int HW, void* stream);

/* ImageEnhanceByRadar tail backward (vr_coc.py:314-315, data_normal :59-67): given dyv = dL/d((1 + kn) * image),
 *   dimage = dyv*(1 + kn),  dk = dyv*image/(mx - mn),  part[b*C + c] = { sum dkn, sum dkn*k, #(k == mn), #(k == mx) }, dkn = dyv*image;
 * vrcoc_minmax_scatter then adds coef[0] at k == mn and coef[1] at k == mx (the gradient through the global min / max, spread
 * evenly over ties like torch's min() / max() backward). */
int vrcoc_img_enh_bwd(const void* dyv, const void* image, const void* k, int dtype, const uint32_t* minmax, int B, int C, int HW,
                      void* dimage, void* dk, float* part, void* stream);
int vrcoc_minmax_scatter(const void* k, void* dk, int dtype, const uint32_t* minmax, const float* coef /*device [2]*/, int64_t n,
                         void* stream);

/* Weight (and bias) gradient of the depthwise 3x3 / stride 1 / pad 1 convolution of the head's DWConv towers (normal_conv.py:23-33):
 * dW [C][3][3] fp32, db [C] fp32 (may be NULL); W % 8 == 0.  The input gradient is vrcoc_dwconv on dy with the flipped kernel. */
int vrcoc_dwconv3_wgrad(const void* x, const void* dy, int dtype, int B, int C, int H, int W, float* dW, float* db, void* stream);
/* BatchNorm2d bookkeeping of the training path (normal_conv.py:45-49; vr_coc.py:315,341,356), one launch each instead of 20-30
 * [C]-sized library launches.
 *   bn_stats   : per-(b,c) sums {sum x, sum x^2} [B][C][2] -> batch mean / biased variance (fp64, optional), the running-stat update
 *                of nn.BatchNorm2d (momentum >= 0 and non-NULL running_*: unbiased variance; num_batches_tracked += 1 when given) and
 *                the folded affine BN(x) = x*scale + shift (optional; gamma / beta may be NULL).
 *   bn_bwd_coef: backward of y = act(BN(u)): z = u*zs + zt recomputes the pre-activation (optional outputs); with sums [B][C][2] =
 *                {sum g, sum g*u}: du = g*ca + u*cb + cd (cb, cd only in training mode), dgamma, dbeta. */
int vrcoc_bn_stats(const float* chan_sums, int B, int C, double count, const float* gamma, const float* beta, float eps, float momentum,
                   float* running_mean, float* running_var, long long* num_batches_tracked, float* scale, float* shift, double* mean_out,
                   double* var_out, void* stream);
int vrcoc_bn_bwd_coef(const float* sums, int B, int C, const double* mean, const double* var, const float* gamma, const float* beta, float eps,
                      double N, int training, float* zs, float* zt, float* ca, float* cb, float* cd, float* dgamma, float* dbeta, void* stream);
/* YOLOX decode of the three detection maps [B][channels][h][w] (channels = 4 box + 1 objectness + classes) into
 * out [B][h3*w3 + h4*w4 + h5*w5][channels] fp32, normalised box centre / size + sigmoid scores: utils/utils_bbox.py:32-84
 * (decode_outputs) in one launch, no intermediate cat / permute / grid tensors. */
int vrcoc_decode_outputs(const void* p3, const void* p4, const void* p5, int dtype, int B, int channels, int h3, int w3, int h4, int w4,
                         int h5, int w5, int input_h, int input_w, float* out, void* stream);
/* Adjoint of vrcoc_upsample_bilinear (align_corners=True): dx[planes][H][W] from dy[planes][Ho][Wo]. */
int vrcoc_upsample_bilinear_bwd(const void* dy, void* dx, int dtype, int planes, int H, int W, int Ho, int Wo, void* stream);
/* Backward of the table-driven prologue z = x*s*h(x)*e, h = sigmoid(ga*x + gc) (the ShuffleAttention gates and the ECA scale in front
 * of RadarEnhanceByImage's projection, vr_coc.py:344-350; also ShuffleAttention / eca_block on their own), source-channel order:
 *   sums : out[b][c] = { sum dz*x*h, sum dz*x^2*h(1-h), sum dz*x*h(1-h), sum x*h, sum x^2*h(1-h), sum x*h(1-h) }
 *   apply: dx = (dz*cd + cj)*(h + x*ga*h(1-h)) + c1 + 2*x*c2 (+ extra),   coef[b][c] = {cd, cj, c1, c2}
 * dz is [B][K][HW] in LOGICAL channel order and is read at kidx[c] (NULL = identity); gate [B][C][2] = {ga, gc} (NULL = no gate).
 * The O(B*C) chain from the sums to the coefficients (means / variances -> gates -> ECA conv1d) lives on the host side. */
int vrcoc_table_bwd_sums(const void* dz, const void* x, int dtype, const int32_t* kidx, const float* gate, int B, int C, int K, int HW,
                         float* out, void* stream);
int vrcoc_table_bwd_apply(const void* dz, const void* x, const void* extra, void* out, int dtype, const int32_t* kidx, const float* gate,
                          const float* coef, int B, int C, int K, int HW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VRCOC_H_ */
