// C-ABI entry points of the convolution family: ekl_conv descriptor -> gather-GEMM plans -> kernels.
#include "../../include/ekl_b200.h"
#include "conv_plan.h"
#include "ekl_common.cuh"

int ekl_tc_supported(const EklGather* g);
void ekl_tc_geometry(const EklGather* g, int group_b, int* tb, int* th, int* tw);
int ekl_gather_gemm_tc(const EklGather* g, const void* w_packed, double* stats, int group_b, int act, const float* bias9,
                       float* scratch, int* mtiles_out, cudaStream_t st, int w_is_fwd_packed = 0);
int ekl_tc_dgrad_from_fwd_ok(const EklGather* g);
int64_t ekl_tc_split_elems(const EklGather* g, int group_b);
int ekl_splitk_finish(float* scratch, int64_t M, int C, int groups, void* y, double* sums, cudaStream_t st);
int ekl_splitk_bn_fusable(int64_t M, int C, int groups, int act);
int ekl_splitk_bn_act_fwd(float* scratch, int64_t M, int C, int groups, float eps, float momentum, float* mean, float* rstd,
                          float* running_mean, float* running_var, const float* gamma, const float* beta, int act, void* y,
                          void* out, void* aux, cudaStream_t st);
int ekl_rw_supported(const EklGather* g, int group_b);
int ekl_conv3x3_rw(const EklGather* g, const void* w_packed, double* stats, int act, const float* bias9, cudaStream_t st);
int ekl_gather_simt(const EklGather* g, const void* w_packed, int act, cudaStream_t st);
int ekl_wgrad_simt(const EklGather* fwd_plan, float* dw_master, cudaStream_t st);
int ekl_wgrad_tc_supported(const EklGather* g);
int ekl_wgrad_tc(const EklGather* g, float* dw, cudaStream_t st);
int ekl_pack_weights(const EklGather* g, const float* w_master, void* out, int Cout, int Cin, cudaStream_t st);

#include <stdlib.h>

namespace {

// EKL_DISABLE_RW=1 routes the small-channel 3x3 layers through the generic gather-GEMM kernel (A/B measurements)
bool rw_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EKL_DISABLE_RW"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// stride-1 3x3 plans with one K block per tap run on the resident-filter kernel (conv_rw.cu), everything else on the
// generic gather-GEMM kernel (conv_tc.cu)
int run_tc(const EklGather* g, const void* w, double* stats, int group_b, int act, const float* bias9, cudaStream_t st) {
  if (rw_enabled() && ekl_rw_supported(g, group_b)) return ekl_conv3x3_rw(g, w, stats, act, bias9, st);
  return ekl_gather_gemm_tc(g, w, stats, group_b, act, bias9, nullptr, nullptr, st);
}

// fp32 workspace elements of the split-K path for this plan (0 = the plan does not split)
int64_t split_elems(const ekl_conv* c, const EklGather* g, int group_b) {
  if (c->impl != EKL_IMPL_TC || c->act != EKL_ACT_NONE || c->x_fmt != 0 || c->y_fmt != 0) return 0;
  if (rw_enabled() && ekl_rw_supported(g, group_b)) return 0;
  static int on = -1;
  if (on < 0) { const char* e = getenv("EKL_DISABLE_SPLITK"); on = (e && e[0] == '1') ? 0 : 1; }
  return on ? ekl_tc_split_elems(g, group_b) : 0;
}

// split-K conv into `ws` (zero on entry, zero on exit) + finish into the bf16 output `out` (+ statistics)
int run_split(const EklGather* g, const void* w, int group_b, float* ws, void* out, double* stats, cudaStream_t st,
              int w_is_fwd = 0) {
  if (int rc = ekl_gather_gemm_tc(g, w, nullptr, group_b, 0, nullptr, ws, nullptr, st, w_is_fwd)) return rc;
  const int groups = (group_b > 0 && g->mB % group_b == 0) ? g->mB / group_b : 1;
  return ekl_splitk_finish(ws, (int64_t)g->mB * g->mH * g->mW, g->N, stats ? groups : 1, out, stats, st);
}

void out_extent(const ekl_conv* c, int* Ho, int* Wo) {
  if (c->mode == EKL_UP2) { *Ho = 2 * c->H; *Wo = 2 * c->W; }
  else if (c->mode == EKL_DOWN2) { *Ho = c->H / 2; *Wo = c->W / 2; }
  else { *Ho = c->H; *Wo = c->W; }
}

EklView make_view(const void* p, int B, int H, int W, int C, int fmt) {
  if (fmt == EKL_FMT_NCHW_F32) {
    EklView v;
    v.base = const_cast<void*>(p); v.sC = (int64_t)H * W; v.sW = 1; v.sH = W; v.sB = (int64_t)C * H * W;
    v.dB = B; v.dH = H; v.dW = W; v.C = C; v.f32 = 1;
    return v;
  }
  return ekl_view_nhwc(const_cast<void*>(p), B, H, W, C, 0);
}

int check(const ekl_conv* c) {
  EKL_REQUIRE(c != nullptr, "null ekl_conv");
  EKL_REQUIRE(c->mode >= EKL_S1 && c->mode <= EKL_DOWN2, "bad conv mode %d", c->mode);
  EKL_REQUIRE(c->B > 0 && c->H > 0 && c->W > 0 && c->Cin > 0 && c->Cout > 0, "bad conv extents");
  EKL_REQUIRE(c->mode != EKL_DOWN2 || (c->H % 2 == 0 && c->W % 2 == 0), "DOWN2 needs even H, W");
  EKL_REQUIRE(c->w_cin_total == 0 || (c->w_cin_off >= 0 && c->w_cin_off + c->Cin <= c->w_cin_total && c->w_layout == EKL_W_KRSC &&
                                      c->w_cin_off % 4 == 0 && c->w_cin_total % 4 == 0),
              "bad master-filter channel window (KRSC layout, offsets %% 4)");
  EKL_REQUIRE(c->w_cout_valid >= 0 && c->w_cout_valid <= c->Cout, "bad w_cout_valid");
  return 0;
}

int plan(const ekl_conv* c, int dgrad, const void* x, const void* y, EklGather* g) {
  int Ho, Wo;
  out_extent(c, &Ho, &Wo);
  EklView xv = make_view(x, c->B, c->H, c->W, c->Cin, c->x_fmt);
  EklView yv = make_view(y, c->B, Ho, Wo, c->Cout, c->y_fmt);
  int rc = ekl_build_gather(g, c->mode, dgrad, xv, yv, c->Cin, c->Cout);
  g->w_kcrs = c->w_layout == EKL_W_KCRS;
  g->w_ld = c->w_cin_total > 0 ? c->w_cin_total : c->Cin;
  g->w_off = c->w_cin_off;
  g->w_cout = c->w_cout_valid > 0 ? c->w_cout_valid : c->Cout;
  return rc;
}

}  // namespace

extern "C" int64_t ekl_conv_packed_elems(const ekl_conv* c, int dgrad) {
  if (check(c)) return -1;
  EklGather g;
  plan(c, dgrad, nullptr, nullptr, &g);
  return ekl_packed_elems(&g);
}

extern "C" int ekl_conv_pack(const ekl_conv* c, const float* w_master, void* w_fwd, void* w_dgrad, void* stream) {
  if (int rc = check(c)) return rc;
  EklGather g;
  if (w_fwd) {
    plan(c, 0, nullptr, nullptr, &g);
    if (int rc = ekl_pack_weights(&g, w_master, w_fwd, c->Cout, c->Cin, (cudaStream_t)stream)) return rc;
  }
  if (w_dgrad) {
    plan(c, 1, nullptr, nullptr, &g);
    if (int rc = ekl_pack_weights(&g, w_master, w_dgrad, c->Cout, c->Cin, (cudaStream_t)stream)) return rc;
  }
  return 0;
}

extern "C" int ekl_conv_fwd(const ekl_conv* c, const void* x, const void* w_fwd, void* y, double* stats, void* stream) {
  if (int rc = check(c)) return rc;
  EKL_REQUIRE(x != nullptr && w_fwd != nullptr && y != nullptr, "conv_fwd: null pointer argument");
  EklGather g;
  plan(c, 0, x, y, &g);
  if (c->impl == EKL_IMPL_SIMT) {
    EKL_REQUIRE(stats == nullptr, "SIMT conv does not produce BatchNorm statistics (use ekl_col_stats)");
    const int act = c->act == EKL_ACT_LRELU ? 1 : (c->act == EKL_ACT_TANH ? 2 : 0);
    return ekl_gather_simt(&g, w_fwd, act, (cudaStream_t)stream);
  }
  EKL_REQUIRE(c->x_fmt == 0 && c->y_fmt == 0, "TC conv: NHWC bf16 only");
  EKL_REQUIRE(c->act == EKL_ACT_NONE || c->act == EKL_ACT_LRELU || c->act == EKL_ACT_TANH,
              "TC conv: fused epilogue activation must be none, LeakyReLU or tanh");
  EKL_REQUIRE(c->act == EKL_ACT_NONE || stats == nullptr, "TC conv: BatchNorm partials are of the raw conv output");
  return run_tc(&g, w_fwd, stats, c->group_b, c->act, nullptr, (cudaStream_t)stream);
}

// y = conv3x3(x) + bias9[b][border class of (h,w)][:]  -- the spatially constant (tiled condition code) input channels
// of a jointConv (model.py:411-414, 403) folded into a per-sample bias with 9 border variants: class = 3*rc + cc,
// rc / cc = 0 first row / column, 2 last, 1 interior.  EKL_S1, tcgen05 implementation only.
extern "C" int ekl_conv_fwd_bias9(const ekl_conv* c, const void* x, const void* w_fwd, const float* bias9, void* y,
                                  double* stats, void* stream) {
  if (int rc = check(c)) return rc;
  EKL_REQUIRE(c->mode == EKL_S1 && c->impl == EKL_IMPL_TC && c->x_fmt == 0 && c->y_fmt == 0 && c->act == EKL_ACT_NONE,
              "conv_fwd_bias9: stride-1 3x3, tcgen05, NHWC bf16, no activation");
  EKL_REQUIRE(c->H >= 2 && c->W >= 2, "conv_fwd_bias9: H, W >= 2");
  EklGather g;
  plan(c, 0, x, y, &g);
  return run_tc(&g, w_fwd, stats, c->group_b, 0, bias9, (cudaStream_t)stream);
}

extern "C" int ekl_conv_bwd_data(const ekl_conv* c, const void* dy, const void* w_dgrad, void* dx, void* stream) {
  if (int rc = check(c)) return rc;
  EKL_REQUIRE(dy != nullptr && w_dgrad != nullptr && dx != nullptr, "conv_bwd_data: null pointer argument");
  EklGather g;
  plan(c, 1, dx, dy, &g);
  if (c->impl == EKL_IMPL_SIMT) return ekl_gather_simt(&g, w_dgrad, 0, (cudaStream_t)stream);
  EKL_REQUIRE(c->x_fmt == 0 && c->y_fmt == 0, "TC conv: NHWC bf16 only");
  return run_tc(&g, w_dgrad, nullptr, 0, 0, nullptr, (cudaStream_t)stream);
}

// ---- workspace variants: few-tile / long-contraction plans run split-K through an fp32 workspace the caller owns
// (ekl_conv_workspace_elems floats, ZERO before the first call; every call leaves it zero again).  With ws == NULL or a
// plan that does not split they are the plain calls.
extern "C" int64_t ekl_conv_workspace_elems(const ekl_conv* c, int dgrad) {
  if (check(c)) return -1;
  EklGather g;
  plan(c, dgrad, nullptr, nullptr, &g);
  return split_elems(c, &g, dgrad ? 0 : c->group_b);
}

extern "C" int ekl_conv_fwd_ws(const ekl_conv* c, const void* x, const void* w_fwd, void* y, double* stats, float* ws,
                               void* stream) {
  if (int rc = check(c)) return rc;
  EklGather g;
  plan(c, 0, x, y, &g);
  if (ws == nullptr || split_elems(c, &g, c->group_b) == 0) return ekl_conv_fwd(c, x, w_fwd, y, stats, stream);
  return run_split(&g, w_fwd, c->group_b, ws, y, stats, (cudaStream_t)stream);
}

// Split-K convolution followed by train-mode BatchNorm + activation as TWO launches instead of three: the conv's work
// items red-add into ws, then one kernel finishes y (bf16, kept for the backward pass), takes the batch statistics,
// updates the running statistics and writes out = act(BN(y)).  ekl_conv_split_bn_fusable(c, act) != 0 says when: the
// plan splits (ekl_conv_workspace_elems > 0), act is none / LeakyReLU / ReLU, Cout % 32 == 0 and a statistics group
// has <= 768 pixels.  aux: ekl_conv_split_bn_aux_floats(c) zeroed floats of caller scratch (left zero).
extern "C" int ekl_conv_split_bn_fusable(const ekl_conv* c, int bn_act) {
  if (check(c) || c->mode == EKL_UP2) return 0;
  EklGather g;
  plan(c, 0, nullptr, nullptr, &g);
  if (split_elems(c, &g, c->group_b) == 0) return 0;
  const int groups = (c->group_b > 0 && c->B % c->group_b == 0) ? c->B / c->group_b : 1;
  return ekl_splitk_bn_fusable((int64_t)g.mB * g.mH * g.mW, g.N, groups, bn_act);
}

extern "C" int64_t ekl_conv_split_bn_aux_floats(const ekl_conv* c) {
  if (check(c)) return -1;
  const int groups = (c->group_b > 0 && c->B % c->group_b == 0) ? c->B / c->group_b : 1;
  return (int64_t)c->Cout / 32 + (int64_t)groups * c->Cout;
}

extern "C" int ekl_conv_fwd_split_bn_act(const ekl_conv* c, const void* x, const void* w_fwd, float* ws, void* y, float eps,
                                         float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                                         const float* gamma, const float* beta, int bn_act, void* out, void* aux, void* stream) {
  if (int rc = check(c)) return rc;
  EKL_REQUIRE(x != nullptr && w_fwd != nullptr && ws != nullptr && y != nullptr, "conv_fwd_split_bn_act: null pointer argument");
  EKL_REQUIRE(ekl_conv_split_bn_fusable(c, bn_act), "conv_fwd_split_bn_act: this layer / shape does not take the fused path");
  EklGather g;
  plan(c, 0, x, y, &g);
  if (int rc = ekl_gather_gemm_tc(&g, w_fwd, nullptr, c->group_b, 0, nullptr, ws, nullptr, (cudaStream_t)stream)) return rc;
  const int groups = (c->group_b > 0 && c->B % c->group_b == 0) ? c->B / c->group_b : 1;
  return ekl_splitk_bn_act_fwd(ws, (int64_t)g.mB * g.mH * g.mW, g.N, groups, eps, momentum, mean, rstd, running_mean, running_var,
                               gamma, beta, bn_act, y, out, aux, (cudaStream_t)stream);
}

extern "C" int ekl_conv_bwd_data_ws(const ekl_conv* c, const void* dy, const void* w_dgrad, void* dx, float* ws, void* stream) {
  if (int rc = check(c)) return rc;
  EklGather g;
  plan(c, 1, dx, dy, &g);
  if (ws == nullptr || split_elems(c, &g, 0) == 0) return ekl_conv_bwd_data(c, dy, w_dgrad, dx, stream);
  return run_split(&g, w_dgrad, 0, ws, dx, nullptr, (cudaStream_t)stream);
}

// Data-gradient reading the FORWARD-packed filter (ekl_conv_pack's w_fwd, or the optimiser's bf16 shadow of a
// channels_last master) as an MN-major tensor-core operand: no transposed copy of the filter is needed.  Available when
// ekl_conv_dgrad_from_fwd(c) != 0 (stride-1 / stride-2 convs, Cin and Cout multiples of 64, generic tcgen05 kernel).
extern "C" int ekl_conv_dgrad_from_fwd(const ekl_conv* c) {
  if (check(c) || c->impl != EKL_IMPL_TC || c->x_fmt != 0 || c->y_fmt != 0 || c->mode == EKL_UP2) return 0;
  EklGather g;
  plan(c, 1, nullptr, nullptr, &g);
  if (rw_enabled() && ekl_rw_supported(&g, 0)) return 0;
  return ekl_tc_dgrad_from_fwd_ok(&g);
}

// Which kernel family a conv call of this descriptor runs on (accounting / profiling only):
// 0 generic gather-GEMM kernel, 1 resident-filter 3x3 kernel, 2 generic kernel split-K (+ finishing pass).
extern "C" int ekl_conv_route(const ekl_conv* c, int dgrad) {
  if (check(c) || c->impl != EKL_IMPL_TC) return -1;
  EklGather g;
  plan(c, dgrad, nullptr, nullptr, &g);
  const int group_b = dgrad ? 0 : c->group_b;
  if (rw_enabled() && ekl_rw_supported(&g, group_b)) return 1;
  return split_elems(c, &g, group_b) > 0 ? 2 : 0;
}

extern "C" int ekl_conv_bwd_data_fw(const ekl_conv* c, const void* dy, const void* w_fwd, void* dx, float* ws, void* stream) {
  if (int rc = check(c)) return rc;
  EKL_REQUIRE(ekl_conv_dgrad_from_fwd(c), "conv_bwd_data_fw: this layer needs the transposed operand (ekl_conv_bwd_data)");
  EKL_REQUIRE(dy != nullptr && w_fwd != nullptr && dx != nullptr, "conv_bwd_data_fw: null pointer argument");
  EklGather g;
  plan(c, 1, dx, dy, &g);
  if (ws != nullptr && split_elems(c, &g, 0) > 0) return run_split(&g, w_fwd, 0, ws, dx, nullptr, (cudaStream_t)stream, 1);
  return ekl_gather_gemm_tc(&g, w_fwd, nullptr, 0, 0, nullptr, nullptr, nullptr, (cudaStream_t)stream, 1);
}

extern "C" int ekl_conv_bwd_weight(const ekl_conv* c, const void* x, const void* dy, float* dw, void* stream) {
  if (int rc = check(c)) return rc;
  EklGather g;
  plan(c, 0, x, dy, &g);
  if (c->impl == EKL_IMPL_SIMT) return ekl_wgrad_simt(&g, dw, (cudaStream_t)stream);
  EKL_REQUIRE(c->x_fmt == 0 && c->y_fmt == 0, "TC conv: NHWC bf16 only");
  return ekl_wgrad_tc(&g, dw, (cudaStream_t)stream);
}

// Debug / test hook: dump a plan's integer structure so the host logic can be verified without a GPU.
// out: [nvar, ntaps, n_a, mB, mH, mW, Cin, N, transposed, KH, KW] then per (v,t): map, dh, dw, nsrc, src[0..3]
extern "C" int ekl_conv_plan_dump(const ekl_conv* c, int dgrad, int* out, int cap) {
  if (int rc = check(c)) return rc;
  EklGather g;
  plan(c, dgrad, nullptr, nullptr, &g);
  const int need = 11 + g.nvar * g.ntaps * 8;
  EKL_REQUIRE(cap >= need, "plan_dump: need %d ints", need);
  int* o = out;
  *o++ = g.nvar; *o++ = g.ntaps; *o++ = g.n_a; *o++ = g.mB; *o++ = g.mH; *o++ = g.mW; *o++ = g.Cin; *o++ = g.N;
  *o++ = g.transposed; *o++ = g.KH; *o++ = g.KW;
  for (int v = 0; v < g.nvar; ++v)
    for (int t = 0; t < g.ntaps; ++t) {
      const EklTap& tp = g.taps[v][t];
      *o++ = tp.map; *o++ = tp.dh; *o++ = tp.dw; *o++ = tp.nsrc;
      for (int i = 0; i < 4; ++i) *o++ = tp.src[i];
    }
  return 0;
}
