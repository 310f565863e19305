// Gather-GEMM plans: every convolution of the step (forward, data-gradient, weight-gradient; stride-1 3x3,
// nearest-up-x2 + 3x3, 4x4 stride-2) is expressed as
//     out_v[p, n] = sum_{t < ntaps} sum_{k < Cin}  A_{map(v,t)}[p + (dh,dw)(v,t), k] * Wp[v][n][t][k]
// over an "M grid" of pixels p = (b,h,w), for nvar in {1,4} output variants (parity classes).
// Out-of-range gathers read zero (== conv zero padding); parity classes turn stride-2 access into plain
// strided 4-D views, so the tcgen05 kernel can fetch every operand tile with one tiled TMA box.
#pragma once
#include <stdint.h>
#include <string.h>

#define EKL_MAX_TAPS 16
#define EKL_MAX_VAR 4

struct EklTap {
  int8_t map, dh, dw, nsrc;   // A view index, pixel offset, number of master-filter taps summed into this tap
  int8_t src[4];              // master-filter tap indices (kh*KW+kw) this packed tap is the sum of
};

// NHWC-like strided 4-D view. Strides in ELEMENTS. dB/dH/dW are the valid extents (reads outside -> 0, writes dropped).
struct EklView {
  void* base;
  int64_t sB, sH, sW, sC;
  int dB, dH, dW, C;
  int f32;                    // element type: 0 bf16, 1 fp32 (SIMT path only)
};

struct EklGather {
  EklView a[4];
  int n_a;
  EklView o[EKL_MAX_VAR];
  int mB, mH, mW;             // M grid
  int Cin, N, ntaps, nvar;
  EklTap taps[EKL_MAX_VAR][EKL_MAX_TAPS];
  int transposed;             // packed weights are [.][Cin_master][t][Cout_master] (data-gradient) instead of [Cout][t][Cin]
  int KH, KW;                 // master filter taps
  int w_kcrs;                 // master filter / gradient memory is [Cout][Cin][KH][KW] instead of [Cout][KH][KW][Cin]
  // window of the master filter this conv uses: the master has w_ld input channels per (co, tap) of which the conv's
  // channels are [w_off, w_off + Cin_master) (folded jointConv: the code channels are handled as a bias), and w_cout
  // real output channels (image head: 3 of a 16-wide tile; packed rows beyond are zero, their gradients are dropped)
  int w_ld, w_off, w_cout;
};

// element offset of master-filter entry (co, tap, ci); Cin = channels per (co, tap) row of the MASTER (EklGather.w_ld),
// ci already includes the window offset
#define EKL_WIDX(kcrs, co, tap, ci, KK, Cin) \
  ((kcrs) ? (((int64_t)(co) * (Cin) + (ci)) * (KK) + (tap)) : (((int64_t)(co) * (KK) + (tap)) * (Cin) + (ci)))

enum { EKL_CONV_S1 = 0, EKL_CONV_UP2 = 1, EKL_CONV_DOWN2 = 2 };

static inline EklView ekl_view_nhwc(void* base, int B, int H, int W, int C, int f32 = 0) {
  EklView v;
  v.base = base; v.sC = 1; v.sW = C; v.sH = (int64_t)W * C; v.sB = (int64_t)H * W * C;
  v.dB = B; v.dH = H; v.dW = W; v.C = C; v.f32 = f32;
  return v;
}
// parity class (ph,pw) of a full-resolution view: pixel (i,j) of the result is pixel (2i+ph, 2j+pw) of v
static inline EklView ekl_view_parity(const EklView& v, int ph, int pw) {
  EklView r = v;
  int64_t esz = v.f32 ? 4 : 2;
  r.base = (char*)v.base + (ph * v.sH + pw * v.sW) * esz;
  r.sH = 2 * v.sH; r.sW = 2 * v.sW; r.dH = v.dH / 2; r.dW = v.dW / 2;
  return r;
}

// rows/cols of the 3x3 master filter that collapse onto tap a in {0,1} of the 2x2 sub-pixel filter of output parity p
static inline int ekl_up_rows(int p, int a, int8_t* out) {
  if (p == 0) { if (a == 0) { out[0] = 0; return 1; } out[0] = 1; out[1] = 2; return 2; }
  if (a == 0) { out[0] = 0; out[1] = 1; return 2; }
  out[0] = 2; return 1;
}

// mode, direction (0 forward, 1 data-gradient). x: conv input view [B,H,W,Cin]; y: conv output view.
// For forward: A = x, out = y. For data-gradient: A = dy (shape of y), out = dx (shape of x).
static inline int ekl_build_gather(EklGather* g, int mode, int dgrad, EklView x, EklView y, int Cin, int Cout) {
  memset(g, 0, sizeof(*g));
  g->transposed = dgrad;
  g->KH = g->KW = (mode == EKL_CONV_DOWN2) ? 4 : 3;
  const int KW = g->KW;
  g->Cin = dgrad ? Cout : Cin;
  g->N = dgrad ? Cin : Cout;
  // which tap structure: "direct" (stride 1), "sub" (4 variants x 2x2 taps, writes parity views), "strided" (16 taps on parity views)
  int structure;
  if (mode == EKL_CONV_S1) structure = 0;
  else if ((mode == EKL_CONV_UP2 && !dgrad) || (mode == EKL_CONV_DOWN2 && dgrad)) structure = 1;
  else structure = 2;
  EklView A = dgrad ? y : x, O = dgrad ? x : y;
  if (structure == 0) {
    g->n_a = 1; g->a[0] = A; g->nvar = 1; g->o[0] = O; g->ntaps = 9;
    g->mB = O.dB; g->mH = O.dH; g->mW = O.dW;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        EklTap& t = g->taps[0][kh * 3 + kw];
        t.map = 0; t.nsrc = 1; t.src[0] = (int8_t)(kh * 3 + kw);
        t.dh = (int8_t)(dgrad ? 1 - kh : kh - 1); t.dw = (int8_t)(dgrad ? 1 - kw : kw - 1);
      }
  } else if (structure == 1) {
    // out parity (ph,pw) <- 2x2 taps on the low-resolution A
    g->n_a = 1; g->a[0] = A; g->nvar = 4; g->ntaps = 4;
    g->mB = A.dB; g->mH = A.dH; g->mW = A.dW;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        int v = ph * 2 + pw;
        g->o[v] = ekl_view_parity(O, ph, pw);
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            EklTap& t = g->taps[v][a * 2 + b];
            t.map = 0;
            if (mode == EKL_CONV_UP2) {   // forward of up+conv3x3: source offset (ph-1+a, pw-1+b), summed master taps
              t.dh = (int8_t)(ph - 1 + a); t.dw = (int8_t)(pw - 1 + b);
              int8_t r[2], c[2];
              int nr = ekl_up_rows(ph, a, r), nc = ekl_up_rows(pw, b, c);
              t.nsrc = 0;
              for (int i = 0; i < nr; ++i)
                for (int j = 0; j < nc; ++j) t.src[t.nsrc++] = (int8_t)(r[i] * 3 + c[j]);
            } else {                      // data-gradient of conv4x4/s2/p1: dx[2i+ph] <- kh with kh = ph+1 (mod 2)
              int kh = (ph == 0) ? (a == 0 ? 1 : 3) : (a == 0 ? 0 : 2);
              int kw = (pw == 0) ? (b == 0 ? 1 : 3) : (b == 0 ? 0 : 2);
              t.dh = (int8_t)((2 * 0 + ph + 1 - kh) / 2); t.dw = (int8_t)((pw + 1 - kw) / 2);
              if (ph + 1 - kh < 0) t.dh = -1;     // (ph+1-kh) in {0, -2, +2, 0}: ho = i + (ph+1-kh)/2
              if (pw + 1 - kw < 0) t.dw = -1;
              t.nsrc = 1; t.src[0] = (int8_t)(kh * KW + kw);
            }
          }
      }
  } else {
    // stride-2 gather: 16 taps over the 4 parity views of the high-resolution A
    g->n_a = 4; g->nvar = 1; g->o[0] = O; g->ntaps = 16;
    g->mB = O.dB; g->mH = O.dH; g->mW = O.dW;
    for (int rh = 0; rh < 2; ++rh)
      for (int rw = 0; rw < 2; ++rw) g->a[rh * 2 + rw] = ekl_view_parity(A, rh, rw);
    // tap u in 0..3 reads high-res row 2i-1+u: u=0 -> parity 1, di -1; u=1 -> parity 0, di 0; u=2 -> parity 1, di 0; u=3 -> parity 0, di +1
    static const int8_t par[4] = {1, 0, 1, 0}, off[4] = {-1, 0, 0, 1};
    for (int u = 0; u < 4; ++u)
      for (int w = 0; w < 4; ++w) {
        EklTap& t = g->taps[0][u * 4 + w];
        t.map = (int8_t)(par[u] * 2 + par[w]); t.dh = off[u]; t.dw = off[w];
        if (mode == EKL_CONV_DOWN2) {     // forward conv4x4/s2/p1: master tap (u,w)
          t.nsrc = 1; t.src[0] = (int8_t)(u * 4 + w);
        } else {                          // data-gradient of up+conv3x3: rows u=0->{2}, 1->{1,2}, 2->{0,1}, 3->{0}
          static const int8_t nr[4] = {1, 2, 2, 1};
          static const int8_t rr[4][2] = {{2, 0}, {1, 2}, {0, 1}, {0, 0}};
          t.nsrc = 0;
          for (int i = 0; i < nr[u]; ++i)
            for (int j = 0; j < nr[w]; ++j) t.src[t.nsrc++] = (int8_t)(rr[u][i] * 3 + rr[w][j]);
        }
      }
  }
  return 0;
}

// number of bf16 elements of the packed weights for a plan
static inline int64_t ekl_packed_elems(const EklGather* g) { return (int64_t)g->nvar * g->N * g->ntaps * g->Cin; }
