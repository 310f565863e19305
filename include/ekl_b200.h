/* ekl_b200 -- C ABI of the B200-native (sm_100a) kernels behind the StackGAN++-style G+D training step of
 * Multimodal-Group/Text2img_EKL.
 *
 * The reference has NO native / FFI layer: its hot path is Python nn.Modules calling torch (cuDNN/cuBLAS).
 * This header is therefore the interface a maintainer binds *below* the reference's Python modules; each entry
 * point names the reference operation it replaces (file:line under /root/reference).  INTEGRATION.md shows the
 * ctypes stub.
 *
 * Conventions
 *  - plain C: device pointers + sizes, no torch types.  The library never allocates or frees caller tensors;
 *    scratch is passed in (sizes from the *_rows / *_elems queries).
 *  - activations: NHWC bf16 (row = pixel, channels contiguous).  Parameters / gradients / statistics: fp32.
 *    Conv filters: fp32, either [Cout][KH][KW][Cin] (torch channels_last storage of the reference's
 *    [Cout,Cin,KH,KW] parameters; coalesced for the kernels) or torch-default [Cout][Cin][KH][KW] (ekl_conv.w_layout).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and returns
 *    0 = ok, <0 = invalid argument, >0 = cudaError_t / 1000+CUresult.  ekl_last_error() = thread-local text.
 *  - no CPU fallback: ekl_require_sm100() fails on anything but compute capability 10.x.
 */
#ifndef EKL_B200_H_
#define EKL_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EKL_B200_VERSION 100

const char* ekl_last_error(void);
int ekl_version(void);
int ekl_require_sm100(void);

/* ---------------------------------------------------------------- convolutions ----------------------------
 * mode EKL_S1    conv3x3 s1 p1 no bias                      model.py:79-82  (conv3x3; ResBlock :107-123,
 *                                                           Block3x3_relu :98-104, Block3x3_leakRelu :812-818)
 * mode EKL_UP2   nn.Upsample(x2, nearest) + conv3x3         model.py:87-94  (upBlock; the 4x tensor is never
 *                                                           materialised: 4 sub-pixel 2x2 convs, 2.25x fewer MACs)
 * mode EKL_DOWN2 conv4x4 s2 p1 no bias                      model.py:822-850 (downBlock, encode_image_by_16times)
 * forward / backward-data / backward-weight replace the cuDNN fprop / dgrad / wgrad calls autograd makes. */
enum { EKL_S1 = 0, EKL_UP2 = 1, EKL_DOWN2 = 2 };
enum { EKL_IMPL_TC = 0, EKL_IMPL_SIMT = 1 };              /* tcgen05 product path | SIMT small-channel / cross-check */
enum { EKL_FMT_NHWC_BF16 = 0, EKL_FMT_NCHW_F32 = 1 };     /* NCHW fp32 = the loader's images (datasets.py:346) */
enum { EKL_W_KRSC = 0, EKL_W_KCRS = 1 };                  /* KRSC = torch channels_last storage, KCRS = torch default */
enum { EKL_ACT_NONE = 0, EKL_ACT_GLU = 1, EKL_ACT_LRELU = 2, EKL_ACT_RELU = 3, EKL_ACT_TANH = 4 };

typedef struct ekl_conv {
  int mode;          /* EKL_S1 | EKL_UP2 | EKL_DOWN2 */
  int B, H, W;       /* extents of the conv INPUT x */
  int Cin, Cout;
  int group_b;       /* batch extent of one BatchNorm-statistics group (0 = B) */
  int impl;          /* EKL_IMPL_TC | EKL_IMPL_SIMT */
  int x_fmt, y_fmt;  /* SIMT only: EKL_FMT_* of x and y */
  int act;           /* fused epilogue on the conv output: EKL_ACT_NONE | EKL_ACT_LRELU | EKL_ACT_TANH (no BN partials then) */
  int w_layout;      /* master filter / gradient memory: EKL_W_KRSC [Cout][KH][KW][Cin] | EKL_W_KCRS [Cout][Cin][KH][KW] */
  /* Window of the master filter (all 0 = the whole filter).  The master (and its gradient) has w_cin_total input
   * channels of which this conv contracts [w_cin_off, w_cin_off + Cin) -- NEXT_STAGE_G.jointConv (model.py:403) whose
   * first ef channels, the tiled condition code, are folded into ekl_conv_fwd_bias9's bias -- and w_cout_valid real
   * output channels of the Cout-wide tile -- GET_IMAGE_G (model.py:432: 3 filters in a 16-wide tile; the packed rows
   * beyond are zero and their weight gradients are not written). */
  int w_cin_total, w_cin_off, w_cout_valid;
} ekl_conv;

/* bf16 elements of the packed forward (dgrad=0) / data-gradient (dgrad=1) filter operand */
int64_t ekl_conv_packed_elems(const ekl_conv* c, int dgrad);
/* fp32 master filter -> packed bf16 operands (either output may be NULL) */
int ekl_conv_pack(const ekl_conv* c, const float* w_master, void* w_fwd, void* w_dgrad, void* stream);
/* y = conv(x); stats (may be NULL; TC impl): BatchNorm batch statistics of y, double [groups][2][Cout] = per-group,
 * per-channel sum and sum-of-squares of the bf16 values stored, ACCUMULATED with fp64 atomics by the epilogue (zero the
 * buffer first); groups = B / group_b.  Consumed directly by ekl_bn_act_fwd: there is no finalize pass. */
int ekl_conv_fwd(const ekl_conv* c, const void* x, const void* w_fwd, void* y, double* stats, void* stream);
/* y = conv3x3(x) + bias9[b][border class][Cout]: the spatially constant input channels of a jointConv -- the tiled
 * condition code of NEXT_STAGE_G (model.py:411-414, jointConv :403) -- folded into a per-sample fp32 bias with 9
 * border variants (class = 3*rc + cc; rc / cc = 0 first row / column, 1 interior, 2 last).  EKL_S1, EKL_IMPL_TC. */
int ekl_conv_fwd_bias9(const ekl_conv* c, const void* x, const void* w_fwd, const float* bias9, void* y, double* stats,
                       void* stream);
/* The folded code channels themselves (model.py:403, 411-414).  W: the FULL fp32 master filter [N][3][3][Ctot]
 * (channels_last storage of [N, Ctot, 3, 3]; the first ef input channels are the tiled code), code [B][ef] fp32:
 *   ekl_code_bias9_fwd   bias9[b][q][n] = sum_{t valid for class q} sum_{c < ef} code[b][c] * W[n][t][c]
 *   ekl_border_sums9     S[b][q][n] += sum of dy[b,h,w,n] (bf16 NHWC) over the pixels of border class q -- the gradient of
 *                        ekl_conv_fwd_bias9's bias9 argument, one pass over dy; S fp32 [B][9][N], ZERO on entry
 *   ekl_code_bias9_bwd   dcode[b][c] (fp32, ZERO on entry, may be NULL) and dW[n][t][c] += (c < ef; may be NULL) from S */
int ekl_code_bias9_fwd(const float* code, const float* w, int B, int ef, int Ctot, int N, float* bias9, void* stream);
int ekl_border_sums9(const void* dy, int B, int H, int W, int N, float* S, void* stream);
int ekl_code_bias9_bwd(const float* S, const float* code, const float* w, int B, int ef, int Ctot, int N, float* dcode,
                       float* dw, void* stream);
/* dx = conv^T(dy) */
int ekl_conv_bwd_data(const ekl_conv* c, const void* dy, const void* w_dgrad, void* dx, void* stream);
/* Workspace variants.  Plans with few output tiles and a long contraction (the 4x4 / 8x8 discriminator tails) run
 * split-K: up to 4 work items per tile red-add fp32 partial tiles into `ws`, a finishing pass rounds to bf16 and takes
 * the BatchNorm statistics.  ws: ekl_conv_workspace_elems(c, dgrad) floats owned by the caller, ZERO before the
 * first call (every call leaves it zero).  ws == NULL or workspace_elems == 0: identical to the plain calls. */
int64_t ekl_conv_workspace_elems(const ekl_conv* c, int dgrad);
int ekl_conv_fwd_ws(const ekl_conv* c, const void* x, const void* w_fwd, void* y, double* stats, float* ws, void* stream);
int ekl_conv_bwd_data_ws(const ekl_conv* c, const void* dy, const void* w_dgrad, void* dx, float* ws, void* stream);
/* Data-gradient straight from the FORWARD-packed filter (MN-major tensor-core operand; no transposed copy).  Usable
 * when ekl_conv_dgrad_from_fwd(c) != 0: stride-1 / stride-2 convs with Cin, Cout multiples of 64 on the generic kernel.
 * ws: split-K workspace as for ekl_conv_bwd_data_ws (may be NULL). */
int ekl_conv_dgrad_from_fwd(const ekl_conv* c);
int ekl_conv_bwd_data_fw(const ekl_conv* c, const void* dy, const void* w_fwd, void* dx, float* ws, void* stream);
/* Split-K convolution + train-mode BatchNorm + activation (model.py:816-830 downBlock / Block3x3_leakRelu on the 4x4 / 8x8
 * discriminator tails) as two launches: the split conv into ws, then ONE kernel that finishes y (bf16, kept for the
 * backward pass), takes the per-group batch statistics (mean / rstd [groups][Cout] written), applies one running-statistics
 * update per group in group order, and writes out = act(BN(y)) with bn_act in {EKL_ACT_NONE, EKL_ACT_LRELU, EKL_ACT_RELU}.
 * Available when ekl_conv_split_bn_fusable(c, bn_act) != 0 (the plan splits, Cout % 32 == 0, <= 768 pixels per group).
 * aux: ekl_conv_split_bn_aux_floats(c) ZEROED floats of caller scratch (left zero). */
int ekl_conv_split_bn_fusable(const ekl_conv* c, int bn_act);
int64_t ekl_conv_split_bn_aux_floats(const ekl_conv* c);
int ekl_conv_fwd_split_bn_act(const ekl_conv* c, const void* x, const void* w_fwd, float* ws, void* y, float eps, float momentum,
                              float* mean, float* rstd, float* running_mean, float* running_var, const float* gamma,
                              const float* beta, int bn_act, void* out, void* aux, void* stream);
/* accounting only: the kernel family a call runs on -- 0 generic gather-GEMM kernel, 1 resident-filter 3x3 kernel,
 * 2 generic kernel with split-K + finishing pass (-1: not a tcgen05 plan) */
int ekl_conv_route(const ekl_conv* c, int dgrad);
/* dw[Cout][KH][KW][Cin] += x (*) dy   (fp32, accumulated: zero it first for a fresh gradient) */
int ekl_conv_bwd_weight(const ekl_conv* c, const void* x, const void* dy, float* dw, void* stream);

/* test hook: integer structure of the gather plan (taps, parity views, master-filter taps summed per packed tap) */
int ekl_conv_plan_dump(const ekl_conv* c, int dgrad, int* out, int cap);

/* ---------------------------------------------------------------- BatchNorm (train mode) + activation -----
 * replaces nn.BatchNorm2d/1d + GLU (model.py:68-76,91-93) | LeakyReLU(0.2) (:816,826) | ReLU (:177-178) and the
 * ResBlock skip add (:119-123).  y: conv output bf16 [M][C]; rows form `groups` contiguous equal groups with
 * independent batch statistics (the reference's separate real / wrong / fake forwards, cub_trainer...:418-420). */
/* batch statistics of a STORED tensor (layers whose producer is not a conv of this library: the BatchNorm1d of the
 * generator stem): sums [groups][2][C] doubles, accumulated (zero first) */
int ekl_col_stats(const void* y, int64_t M, int C, int groups, double* sums, void* stream);
/* out = act(gamma * (y - mean) * rstd + beta) (+ residual).  Training (sums != NULL): mean / rstd come from the fp64
 * sums [groups][2][Cy] (count = M / groups rows per group), are WRITTEN to mean / rstd [groups][Cy] for the backward pass,
 * and running_mean / running_var (may be NULL) get one momentum update per group, in group order, with the unbiased
 * variance -- exactly one update per reference forward call.  Inference (sums == NULL): mean / rstd are inputs. */
int ekl_bn_act_fwd(const void* y, int64_t M, int Cy, int groups, const double* sums, float eps, float momentum, float* mean,
                   float* rstd, float* running_mean, float* running_var, const float* gamma, const float* beta, int act,
                   const void* residual, void* out, void* stream);
/* dy from dout; dgamma/dbeta are accumulated (+=).  sums: ekl_bn_bwd_scratch_doubles(M, Cy, groups, act) doubles of
 * caller scratch, ZERO on entry: the two backward reductions sum(dz), sum(dz * xhat) are accumulated there between the two
 * passes, one sum per 128-byte line ([groups][2][Cy] entries EKL_BN_BWD_SPREAD doubles apart) so that the fp64 reds of a
 * block spread over the L2 slices.  Layers with few rows per group run as one launch and need no scratch (0 doubles). */
#define EKL_BN_BWD_SPREAD 16
int64_t ekl_bn_bwd_scratch_doubles(int64_t M, int Cy, int groups, int act);
int ekl_bn_act_bwd(const void* y, const void* dout, int64_t M, int Cy, int groups, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, int act, double* sums, float* dgamma, float* dbeta, void* dy,
                   void* stream);
/* LeakyReLU(0.2) backward from the OUTPUT (first discriminator conv has no BN, model.py:835-836) */
int ekl_lrelu_bwd(const void* out, const void* dout, void* dx, int64_t n, void* stream);
/* cat(tile(c_code), h) along channels (model.py:411-414, 956-959) and its backward (dcode accumulated) */
int ekl_cat_code(const float* code, int Cc, const void* x, int Cx, int B, int HW, void* out, void* stream);
int ekl_cat_code_bwd(const void* dcat, int Cc, int Cx, int B, int HW, float* dcode, void* dx, void* stream);

/* ---------------------------------------------------------------- discriminator stem: image layout ----------
 * First discriminator conv = conv4x4 s2 p1 on the loader's 3-channel NCHW fp32 images (model.py:832-836;
 * datasets.py:346).  Space-to-depth by 2 makes it a 3x3 s1 conv over 12 (padded to 16) channels for the tcgen05
 * kernel:  out[b,i,j,(c*2+ph)*2+pw] = x[b,c,2i+ph,2j+pw]  (bf16 NHWC, channels 12..15 zero).  `groups` (1..3)
 * source batches of B images each (real / wrong / fake, cub_trainer_splitz_cap_ca.py:418-420) are gathered into
 * one [groups*B, H/2, W/2, 16] tensor.  ekl_img_s2d_bwd is the inverse map for the image gradient (fp32 NCHW). */
int ekl_img_s2d(const float* x0, const float* x1, const float* x2, int groups, int B, int H, int W, void* out, void* stream);
int ekl_img_s2d_bwd(const void* dxs, int B, int H, int W, float* dx, void* stream);

/* ---------------------------------------------------------------- loader image pyramid on the device ----------
 * datasets.py:43-68 get_imgs: every stage below the last gets a PIL BILINEAR resize (transforms.Scale) of the final-size
 * uint8 crop; each level is ToTensor()'d and Normalize((0.5,)*3, (0.5,)*3)'d.  One level per call, bit-exact with
 * PIL's 8-bit resampler (integer weights with 22 fractional bits, horizontal then vertical pass, uint8 intermediate):
 * src uint8 [B][S][S][3] (HWC), out fp32 [B][3][s][s].  bounds [s][2] = (first source index, count) and kk [s][ksize] are
 * PIL's coefficient tables of the resize S -> s (device int32, computed on the host: text2img_ekl_b200/datasets.py);
 * tmp: B*S*s*3 bytes of scratch.  s == S converts without resampling. */
int ekl_img_pyramid_level(const void* src_u8, int B, int S, int s, const int* bounds, const int* kk, int ksize, void* tmp_u8,
                          float* out, void* stream);

/* ---------------------------------------------------------------- generator image head ---------------------
 * GET_IMAGE_G (model.py:426-437) = conv3x3(ngf -> 3) + tanh.  The conv runs on ekl_conv_fwd with the 3 filters
 * zero-padded to C (>= 8, multiple of 8) output channels; these passes apply tanh to channels 0..2 of the padded
 * NHWC bf16 conv output and write the NCHW fp32 image, and map the image gradient back (padding channels zero). */
int ekl_head_tanh_fwd(const void* y, int B, int HW, int C, float* img, void* stream);
int ekl_head_tanh_bwd(const void* y, const float* dimg, int B, int HW, int C, void* dy, void* stream);

/* ---------------------------------------------------------------- colour-consistency statistics -----------
 * compute_mean_covariance (cub_trainer_splitz_cap_ca.py:33-52): img fp32 NCHW [B,3,HW] -> mean [B,3] and channel
 * covariance [B,3,3] (divided by HW).  `scratch`: 9*B doubles of caller workspace (zeroed here).  HW % 4 == 0.
 * The backward pass takes the gradients of both outputs (either may be NULL) and writes dimg [B,3,HW]. */
int ekl_color_stats_fwd(const float* img, int B, int HW, double* scratch, float* mean, float* cov, void* stream);
int ekl_color_stats_bwd(const float* img, const float* mean, const float* dmean, const float* dcov, int B, int HW,
                        float* dimg, void* stream);

/* ---------------------------------------------------------------- discriminator heads + GAN losses -----------
 * Heads: `logits` / `uncond_logits` = Conv2d(8ndf, 1, 4, stride 4) + Sigmoid on the 4x4 code map (model.py:886-888,
 * 935-952) == one dot of length K = 16*8ndf per sample.  x_code / h_c: bf16 [GB][K] (NHWC-flattened trunk output /
 * jointConv output, h_c may be NULL); w_*: fp32 [K] in the same element order (channels_last storage of the
 * [1,8ndf,4,4] filter); outputs are RAW logits (the sigmoid is applied by the loss kernel / the module forward).
 * Backward: dx_code / dh_c are written (bf16, either may be NULL), parameter gradients accumulated (+=, may be NULL). */
int ekl_dhead_dots(const void* x_code, const void* h_c, const float* w_u, const float* b_u, const float* w_m,
                   const float* b_m, int GB, int K, float* logit_u, float* logit_m, void* stream);
int ekl_dhead_dots_bwd(const void* x_code, const void* h_c, const float* w_u, const float* w_m, const float* g_u,
                       const float* g_m, int GB, int K, void* dx_code, void* dh_c, float* dw_u, float* db_u,
                       float* dw_m, float* db_m, void* stream);
/* Losses of one discriminator pass over `groups` (1..3) stacked groups of B samples
 * (cub_trainer_splitz_cap_ca.py:423-448 train_joint_Dnet: groups real/wrong/fake; :470-487 loss_joint_Gnet: one
 * group): per group a constant 0/1 BCE label for the match and the uncond probability (t_match / t_uncond, host
 * int[groups]) and a soft class target set (cls_tgt[g]: -1 none, 0 -> cp0, 1 -> cp1; each [B][E1]).
 * losses[4] = {total, match, uncond_coeff * uncond, class}; p_m / p_u [groups*B] sigmoid probabilities;
 * logp [groups*B][E1] = log_softmax(cls_logits).  torch semantics: BCE log clamped at -100, means over B,
 * ce_loss = -sum(p * logq) / B (cub:60-65).  ekl_dloss_bwd: gradients of losses[0] w.r.t. the raw logits,
 * scaled by the device scalar go[0]. */
int ekl_dloss_fwd(int groups, int B, int E1, const int* t_match, const int* t_uncond, const int* cls_tgt,
                  float uncond_coeff, const float* logit_m, const float* logit_u, const float* cls_logits,
                  const float* cp0, const float* cp1, float* losses, float* p_m, float* p_u, float* logp, void* stream);
int ekl_dloss_bwd(int groups, int B, int E1, const int* t_match, const int* t_uncond, const int* cls_tgt,
                  float uncond_coeff, const float* go, const float* p_m, const float* p_u, const float* logp,
                  const float* cp0, const float* cp1, float* g_m, float* g_u, float* g_cls, void* stream);

/* ---------------------------------------------------------------- conditioning augmentation + KL --------------
 * CA_NET.reparametrize / VC_NET.reparameterize (model.py:145-152, 182-184, 198) fused with KL_loss
 * (cub_trainer_splitz_cap_ca.py:54-58):  std = exp(0.5*logvar), c = eps*std + mu,
 * kl[0] = -0.5 * mean(1 + logvar - mu^2 - exp(logvar)).  mu / logvar: fp32 [B][D] with row strides in elements (they
 * are column halves of one Linear / GLU output); eps, c, std: dense [B][D].  Backward: dc / dstd / dkl may be NULL. */
int ekl_reparam_kl_fwd(const float* mu, int64_t mu_row_stride, const float* logvar, int64_t lv_row_stride,
                       const float* eps, int B, int D, float* c, float* stdv, float* kl, void* stream);
int ekl_reparam_kl_bwd(const float* mu, int64_t mu_row_stride, const float* logvar, int64_t lv_row_stride,
                       const float* eps, int B, int D, const float* dc, const float* dstd, const float* dkl, float* dmu,
                       float* dlogvar, void* stream);

/* VC_NET's hidden layers (model.py:169-181): h = ReLU(BatchNorm1d(x W^T + b)) on a batch of B <= 64 rows as one kernel
 * per direction (a warp owns an output column for all rows: the batch statistics stay in registers).  All fp32.
 * forward: training != 0 -> batch statistics (running_mean / running_var updated, may be NULL), xhat [B][N] and rstd [N]
 * saved for the backward; training == 0 -> running statistics (xhat / rstd may be NULL).
 * backward: dW [N][K], dbias, dgamma, dbeta accumulated (+=, each may be NULL); dy [B][N] = gradient of the Linear's
 * output (the caller forms dx = dy W with one library GEMM where the input needs a gradient). */
int ekl_linear_bn_relu_fwd(const float* x, const float* W, const float* bias, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int B, int K, int N, float eps, float momentum,
                           int training, float* h, float* xhat, float* rstd, void* stream);
int ekl_linear_bn_relu_bwd(const float* dh, const float* h, const float* xhat, const float* rstd, const float* gamma,
                           const float* x, int B, int K, int N, float* dW, float* dbias, float* dgamma, float* dbeta,
                           float* dy, void* stream);

/* ---------------------------------------------------------------- capsule routing (generator stem) -----------
 * CapsuleLinear of COND_INIT_STAGE_G_withCap (model.py:245-267; third-party `capsule_layer`, un-vendored: parity is
 * against oracle/capsule_ref.py).  Shared weights W [O][L][K]; x [B][I][K].  Dynamic routing in K space -- the
 * [B,O,I,L] prior tensor is never formed:
 *   proj_u   u[b,o,:] = W[o]^T v[b,o,:]                    proj_s   s[b,o,:] = W[o] y[b,o,:] (+ v = squash(s))
 *   agree    y[b,o,:] = sum_i softmax_o(<x[b,i], u[b,o]>) x[b,i]   (logits / couplings live in registers; M, Z [B][I]
 *            are the softmax max / normaliser saved for the backward)
 *   squash_bwd gs = d squash(s)^T gv, gy = W^T gs           outer    gW[o,l,k] += sum_b A[b,o,l] Bm[b,o,k]
 * Small-K regime only (ekl_caps_supported): K in {4,8}, L <= 64, I % 16 == 0. All tensors fp32, dense. */
int ekl_caps_supported(int I, int K, int O, int L);
int ekl_caps_proj_u(const float* W, const float* v, int B, int O, int L, int K, float* u, void* stream);
int ekl_caps_proj_s(const float* W, const float* y, int B, int O, int L, int K, float* s, float* v_squashed, void* stream);
int ekl_caps_squash_bwd(const float* W, const float* s, const float* gv, int B, int O, int L, int K, float* gs, float* gy,
                        void* stream);
int ekl_caps_outer(const float* A, const float* Bm, int B, int O, int L, int K, float* gW, void* stream);
int ekl_caps_agree_fwd(const float* x, const float* u, int B, int I, int O, int K, float* y, float* M, float* Z, void* stream);
int ekl_caps_agree_bwd(const float* x, const float* u, const float* M, const float* Z, const float* gy, int B, int I, int O,
                       int K, float* gu, float* gx, void* stream);

/* ---------------------------------------------------------------- capsule routing (discriminator class head) --
 * CapsuleLinear of JOINT_D_NET64/128.fc_ac_cap (model.py:943,967-971,1082: 201 out-capsules of length 16 on the 16
 * positions x 512 channels of the trunk output; parity against oracle/capsule_ref.py like the generator stem).
 * prior [B][I=16][O][L=16] fp32 = x[B*16, 512] W^T (one library GEMM; it IS small here) -> three routing iterations in
 * one kernel: v [B][O][16] (nullable) and norms [B][O] = |v| (nullable; model.py:969-970 takes .norm(dim=-1)).
 * The backward recomputes the routing on chip: g_v [B][O][16] and / or g_norm [B][O] -> g_prior (same layout as prior).
 * Supported: I == 16, L == 16, iters == 3, O <= 224 (ekl_caps_route_supported). */
int ekl_caps_route_supported(int I, int O, int L, int iters);
int ekl_caps_route_fwd(const float* prior, int B, int I, int O, int L, int iters, float* v, float* norms, void* stream);
int ekl_caps_route_bwd(const float* prior, const float* g_v, const float* g_norm, int B, int I, int O, int L, int iters,
                       float* g_prior, void* stream);

/* ---------------------------------------------------------------- optimiser ---------------------------------
 * torch.optim.Adam as define_optimizers configures it (cub_trainer_splitz_cap_ca.py:199-215: lr 2e-4, betas (0.5, 0.999),
 * eps 1e-8, no weight decay) over a network's flat fp32 parameter / gradient / moment buffers (n % 4 == 0, 16-byte
 * aligned).  state: 3 device floats {step, 1-beta1^t, sqrt(1-beta2^t)}, zero at start; the step advances on the device
 * (graph-capturable).  shadow_bf16 (may be NULL): bf16 copy of the updated parameters = the packed forward filter operand
 * of channels_last stride-1 / stride-2 convolutions. */
int ekl_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float* state, float lr,
                  float beta1, float beta2, float eps, void* stream);
/* Data-parallel gradient exchange (replaces nn.DataParallel's fp32 reduce-to-GPU0, cub_trainer_splitz_cap_ca.py:139,163):
 * ekl_cast_bf16 rounds a slice of the flat fp32 gradient buffer to bf16 (the NCCL all-reduce payload: half the NVLink
 * bytes), ekl_adam_step_g16 is ekl_adam_step reading the averaged gradient from that bf16 buffer.  n % 4 == 0. */
int ekl_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
int ekl_adam_step_g16(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, int64_t n, float* state, float lr,
                      float beta1, float beta2, float eps, void* stream);
/* The same update in pieces, so that the optimiser step overlaps the backward pass (the deepest layers' gradients are
 * final first and hold most of a discriminator's parameters): ekl_adam_tick advances the step count and the bias
 * corrections once, ekl_adam_apply then updates any 16-byte aligned slice of the flat buffers, n % 4 == 0, without touching
 * the count.  g_is_bf16: the gradient slice is bf16 (8-byte aligned). */
int ekl_adam_tick(float* state, float beta1, float beta2, void* stream);
int ekl_adam_apply(float* p, const void* g, int g_is_bf16, float* m, float* v, void* shadow_bf16, int64_t n, const float* state,
                   float lr, float beta1, float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EKL_B200_H_ */
