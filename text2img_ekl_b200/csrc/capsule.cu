// Capsule routing-by-agreement for the generator stem (COND_INIT_STAGE_G_withCap.fc_cap, model.py:245-267:
// CapsuleLinear(out_capsules = 1024, in_length = 8, out_length = 32), shared weights W[O][L][K]).
//
// *Parity unpinned*: the reference imports the un-vendored `capsule_layer` package; the arithmetic restated in
// oracle/capsule_ref.py (dynamic routing: softmax over OUT capsules, squash) is what these kernels implement.
//
// The prior tensor prior[b,o,i,:] = W[o] x[b,i] ([B,O,I,L]: 201 MB fp32 at B = 32) is never formed: every routing
// quantity lives in in_length (K) space,
//     u[b,o]   = W[o]^T vsum[b,o]                      (caps_proj_u)
//     a[b,o,i] = <x[b,i], u[b,o]>,  c = softmax_o(a),  y[b,o] = sum_i c[b,o,i] x[b,i]        (caps_agree)
//     s[b,o]   = W[o] y[b,o],  v = squash(s)           (caps_s_squash)
// and the logits / coupling coefficients a, c exist only in registers: one block per sample keeps x[b] in shared
// memory, a thread owns out-capsules, and the softmax statistics over the O axis (max, sum of exponentials, and the
// backward's sum_o c*gc) are block reductions built from warp shuffles.  Small-K regime: K in {4, 8}, L <= 64, I in {16, 32, 48, 64}.
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

constexpr float CAPS_EPS = 1e-8f;      // oracle/capsule_ref.py squash epsilon

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Thread mapping of the per-(b,o) kernels: consecutive threads walk the BATCH index for a fixed out-capsule, so a warp
// reads one W[o] row as broadcast loads (8 KB of W per block instead of 256 KB) and each lane streams its own v / y row.
#define CAPS_BO_INDEX()                                   \
  const int tid_ = blockIdx.x * 256 + threadIdx.x;        \
  if (tid_ >= BO) return;                                 \
  const int o = tid_ / Bn, idx = (tid_ % Bn) * O + o;

// ---- u[b,o,k] = sum_l W[o,l,k] * v[b,o,l]        one thread per (b,o)
template <int K>
__global__ void __launch_bounds__(256) caps_proj_u_kernel(const float* __restrict__ W, const float* __restrict__ v, int BO, int O, int Lh,
                                                          int Bn, float* __restrict__ u) {
  CAPS_BO_INDEX()
  const float* w = W + (int64_t)o * Lh * K;
  const float* vv = v + (int64_t)idx * Lh;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (int l = 0; l < Lh; ++l) {
    const float x = vv[l];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += w[l * K + k] * x;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) u[(int64_t)idx * K + k] = acc[k];
}

// ---- s[b,o,l] = sum_k W[o,l,k] * y[b,o,k];  optionally v = squash(s)  (v may be null)
template <int K>
__global__ void __launch_bounds__(256) caps_proj_s_kernel(const float* __restrict__ W, const float* __restrict__ y, int BO, int O, int Lh,
                                                          int Bn, float* __restrict__ s, float* __restrict__ v) {
  CAPS_BO_INDEX()
  const float* w = W + (int64_t)o * Lh * K;
  float yy[K];
#pragma unroll
  for (int k = 0; k < K; ++k) yy[k] = y[(int64_t)idx * K + k];
  float n2 = 0.f;
  for (int l = 0; l < Lh; ++l) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) a += w[l * K + k] * yy[k];
    s[(int64_t)idx * Lh + l] = a;
    n2 += a * a;
  }
  if (v != nullptr) {
    const float f = n2 / (1.f + n2) * rsqrtf(n2 + CAPS_EPS);
    for (int l = 0; l < Lh; ++l) {       // recomputed (W row is a broadcast L1 hit) rather than re-read from global
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) a += w[l * K + k] * yy[k];
      v[(int64_t)idx * Lh + l] = a * f;
    }
  }
}

// ---- backward of v = squash(s), fused with gy = W^T gs:  gs = gv*f + s * (2 f'(n2) <gv, s>)
template <int K>
__global__ void __launch_bounds__(256) caps_squash_bwd_kernel(const float* __restrict__ W, const float* __restrict__ s,
                                                              const float* __restrict__ gv, int BO, int O, int Lh, int Bn,
                                                              float* __restrict__ gs, float* __restrict__ gy) {
  CAPS_BO_INDEX()
  const float* w = W + (int64_t)o * Lh * K;
  const float* ss = s + (int64_t)idx * Lh;
  const float* gg = gv + (int64_t)idx * Lh;
  float n2 = 0.f, dot = 0.f;
  for (int l = 0; l < Lh; ++l) { n2 += ss[l] * ss[l]; dot += gg[l] * ss[l]; }
  const float r = rsqrtf(n2 + CAPS_EPS), q = 1.f / (1.f + n2);
  const float f = n2 * q * r;
  const float fp = q * r - n2 * q * q * r - 0.5f * n2 * q * r * r * r;      // f'(n2)
  const float coef = 2.f * fp * dot;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (int l = 0; l < Lh; ++l) {
    const float g = gg[l] * f + ss[l] * coef;
    gs[(int64_t)idx * Lh + l] = g;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += w[l * K + k] * g;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) gy[(int64_t)idx * K + k] = acc[k];
}

// ---- gW[o,l,k] += sum_b A[b,o,l] * Bm[b,o,k]        one thread per (o,l): exclusive owner of its K outputs
template <int K>
__global__ void __launch_bounds__(256) caps_outer_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int B, int O, int Lh,
                                                         float* __restrict__ gW) {
  const int idx = blockIdx.x * 256 + threadIdx.x;       // o*Lh + l
  if (idx >= O * Lh) return;
  const int o = idx / Lh;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float a = A[(int64_t)b * O * Lh + idx];
    const float* bm = Bm + ((int64_t)b * O + o) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += a * bm[k];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) gW[(int64_t)idx * K + k] += acc[k];
}

// 16-lane (half-warp) butterfly sum
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int NCH = 4;                 // in-capsule chunks of 16 (I <= 64)
constexpr int AGT = 512;               // threads per block of the agreement kernels
constexpr int AGH = AGT / 16;          // half-warps per block

// Thread mapping of the agreement kernels: one block (512 threads = 32 half-warps) per sample.  Lane j of a half-warp
// owns in-capsules i = 16*ch + j (their x rows live in registers), half-warp h walks out-capsules o = h, h+32, ...
// (u_o / gy_o are broadcast loads).  Softmax statistics over the O axis are therefore lane-local running max / sums,
// combined across the half-warps through shared memory; sums over i are 16-lane shuffles.

// ---- agreement step: y[b,o] = sum_i softmax_o(<x[b,i], u[b,o]>) x[b,i];  saves the softmax statistics M, Z [B][I]
template <int K>
__global__ void __launch_bounds__(AGT) caps_agree_fwd_kernel(const float* __restrict__ x, const float* __restrict__ u, int I, int O,
                                                             float* __restrict__ y, float* __restrict__ Ms, float* __restrict__ Zs) {
  __shared__ float red[AGH][NCH * 16];
  __shared__ float Mi[NCH * 16], Zi[NCH * 16];
  const int b = blockIdx.x;
  const int h = threadIdx.x >> 4, j = threadIdx.x & 15;
  const int nch = I / 16;
  float xr[NCH][K];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int k = 0; k < K; ++k) xr[ch][k] = ch < nch ? x[((int64_t)b * I + ch * 16 + j) * K + k] : 0.f;
  const float* ub = u + (int64_t)b * O * K;
  // pass 1: max over o
  float m[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) m[ch] = -INFINITY;
  for (int o = h; o < O; o += AGH) {
    float uo[K];
#pragma unroll
    for (int k = 0; k < K; ++k) uo[k] = ub[(int64_t)o * K + k];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) a += xr[ch][k] * uo[k];
      m[ch] = fmaxf(m[ch], a);
    }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) red[h][ch * 16 + j] = m[ch];
  __syncthreads();
  if (threadIdx.x < NCH * 16) {
    float r = red[0][threadIdx.x];
    for (int q = 1; q < AGH; ++q) r = fmaxf(r, red[q][threadIdx.x]);
    Mi[threadIdx.x] = r;
  }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) m[ch] = Mi[ch * 16 + j];
  // pass 2: sum of exponentials over o
  float z[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) z[ch] = 0.f;
  for (int o = h; o < O; o += AGH) {
    float uo[K];
#pragma unroll
    for (int k = 0; k < K; ++k) uo[k] = ub[(int64_t)o * K + k];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) a += xr[ch][k] * uo[k];
      z[ch] += __expf(a - m[ch]);
    }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) red[h][ch * 16 + j] = z[ch];
  __syncthreads();
  if (threadIdx.x < NCH * 16) {
    float r = 0.f;
    for (int q = 0; q < AGH; ++q) r += red[q][threadIdx.x];
    Zi[threadIdx.x] = r;
    if ((int)threadIdx.x < I) { Ms[(int64_t)b * I + threadIdx.x] = Mi[threadIdx.x]; Zs[(int64_t)b * I + threadIdx.x] = r; }
  }
  __syncthreads();
  float rz[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) rz[ch] = ch < nch ? 1.f / Zi[ch * 16 + j] : 0.f;
  // pass 3: y_o = sum_i c[o,i] x_i
  const int n_o = (O + AGH - 1) / AGH;
  for (int it = 0; it < n_o; ++it) {
    const int o = h + it * AGH;
    const bool live = o < O;
    float uo[K], yo[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { uo[k] = live ? ub[(int64_t)o * K + k] : 0.f; yo[k] = 0.f; }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) a += xr[ch][k] * uo[k];
      const float c = __expf(a - m[ch]) * rz[ch];
#pragma unroll
      for (int k = 0; k < K; ++k) yo[k] += c * xr[ch][k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) yo[k] = half_sum(yo[k]);
    if (live && j < K) y[((int64_t)b * O + o) * K + j] = yo[j < K ? j : 0];
  }
}

// ---- backward of the agreement step.  gc[o,i] = <gy_o, x_i>;  D_i = sum_o c gc;  ga = c (gc - D_i);
//      gu_o = sum_i ga x_i;  gx_i = sum_o (c gy_o + ga u_o)
template <int K>
__global__ void __launch_bounds__(AGT) caps_agree_bwd_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                             const float* __restrict__ Ms, const float* __restrict__ Zs,
                                                             const float* __restrict__ gy, int I, int O, float* __restrict__ gu,
                                                             float* __restrict__ gx) {
  __shared__ float red[AGH][NCH * 16];
  __shared__ float Di[NCH * 16];
  __shared__ float gxs[NCH * 16 * K];
  const int b = blockIdx.x;
  const int h = threadIdx.x >> 4, j = threadIdx.x & 15;
  const int nch = I / 16;
  float xr[NCH][K], m[NCH], rz[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const bool on = ch < nch;
#pragma unroll
    for (int k = 0; k < K; ++k) xr[ch][k] = on ? x[((int64_t)b * I + ch * 16 + j) * K + k] : 0.f;
    m[ch] = on ? Ms[(int64_t)b * I + ch * 16 + j] : 0.f;
    rz[ch] = on ? 1.f / Zs[(int64_t)b * I + ch * 16 + j] : 0.f;
  }
  for (int i = threadIdx.x; i < NCH * 16 * K; i += AGT) gxs[i] = 0.f;
  const float* ub = u + (int64_t)b * O * K;
  const float* gyb = gy + (int64_t)b * O * K;
  // pass 1: D_i = sum_o c[o,i] * gc[o,i]
  float d[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) d[ch] = 0.f;
  for (int o = h; o < O; o += AGH) {
    float uo[K], go[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { uo[k] = ub[(int64_t)o * K + k]; go[k] = gyb[(int64_t)o * K + k]; }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float a = 0.f, gc = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) { a += xr[ch][k] * uo[k]; gc += xr[ch][k] * go[k]; }
      d[ch] += __expf(a - m[ch]) * rz[ch] * gc;
    }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) red[h][ch * 16 + j] = d[ch];
  __syncthreads();
  if (threadIdx.x < NCH * 16) {
    float r = 0.f;
    for (int q = 0; q < AGH; ++q) r += red[q][threadIdx.x];
    Di[threadIdx.x] = r;
  }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) d[ch] = Di[ch * 16 + j];
  // pass 2: gu_o (16-lane sums) and lane-local gx_i
  float gxi[NCH][K];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int k = 0; k < K; ++k) gxi[ch][k] = 0.f;
  const int n_o = (O + AGH - 1) / AGH;
  for (int it = 0; it < n_o; ++it) {
    const int o = h + it * AGH;
    const bool live = o < O;
    float uo[K], go[K], guo[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      uo[k] = live ? ub[(int64_t)o * K + k] : 0.f;
      go[k] = live ? gyb[(int64_t)o * K + k] : 0.f;
      guo[k] = 0.f;
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float a = 0.f, gc = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) { a += xr[ch][k] * uo[k]; gc += xr[ch][k] * go[k]; }
      const float c = live ? __expf(a - m[ch]) * rz[ch] : 0.f;
      const float ga = c * (gc - d[ch]);
#pragma unroll
      for (int k = 0; k < K; ++k) { guo[k] += ga * xr[ch][k]; gxi[ch][k] += c * go[k] + ga * uo[k]; }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) guo[k] = half_sum(guo[k]);
    if (live && j < K) gu[((int64_t)b * O + o) * K + j] = guo[j < K ? j : 0];
  }
  // combine the half-warps' gx partials
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int k = 0; k < K; ++k) atomicAdd(&gxs[(ch * 16 + j) * K + k], gxi[ch][k]);
  __syncthreads();
  for (int i = threadIdx.x; i < I * K; i += AGT) gx[(int64_t)b * I * K + i] = gxs[i];
}

}  // namespace

#define EKL_CAPS_K(K, CALL)                         \
  switch (K) {                                      \
    case 4: { constexpr int KK = 4; CALL; } break;  \
    case 8: { constexpr int KK = 8; CALL; } break;  \
    default: return ekl_fail(-1, "capsule kernels: in_length must be 4 or 8 (got %d)", K); \
  }

extern "C" int ekl_caps_supported(int I, int K, int O, int Lh) {
  return (K == 4 || K == 8) && Lh >= 1 && Lh <= 64 && I % 16 == 0 && I >= 16 && I <= 64 && O >= 1;
}

extern "C" int ekl_caps_proj_u(const float* W, const float* v, int B, int O, int Lh, int K, float* u, void* stream) {
  const int BO = B * O;
  EKL_CAPS_K(K, (caps_proj_u_kernel<KK><<<ekl_cdiv(BO, 256), 256, 0, (cudaStream_t)stream>>>(W, v, BO, O, Lh, B, u)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_caps_proj_s(const float* W, const float* y, int B, int O, int Lh, int K, float* s, float* v_squashed,
                               void* stream) {
  const int BO = B * O;
  EKL_CAPS_K(K, (caps_proj_s_kernel<KK><<<ekl_cdiv(BO, 256), 256, 0, (cudaStream_t)stream>>>(W, y, BO, O, Lh, B, s, v_squashed)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_caps_squash_bwd(const float* W, const float* s, const float* gv, int B, int O, int Lh, int K, float* gs,
                                   float* gy, void* stream) {
  const int BO = B * O;
  EKL_CAPS_K(K, (caps_squash_bwd_kernel<KK><<<ekl_cdiv(BO, 256), 256, 0, (cudaStream_t)stream>>>(W, s, gv, BO, O, Lh, B, gs, gy)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_caps_outer(const float* A, const float* Bm, int B, int O, int Lh, int K, float* gW, void* stream) {
  EKL_CAPS_K(K, (caps_outer_kernel<KK><<<ekl_cdiv(O * Lh, 256), 256, 0, (cudaStream_t)stream>>>(A, Bm, B, O, Lh, gW)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_caps_agree_fwd(const float* x, const float* u, int B, int I, int O, int K, float* y, float* M, float* Z,
                                  void* stream) {
  EKL_REQUIRE(ekl_caps_supported(I, K, O, 1), "caps_agree: unsupported shape I=%d K=%d", I, K);
  EKL_CAPS_K(K, (caps_agree_fwd_kernel<KK><<<B, AGT, 0, (cudaStream_t)stream>>>(x, u, I, O, y, M, Z)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_caps_agree_bwd(const float* x, const float* u, const float* M, const float* Z, const float* gy, int B,
                                  int I, int O, int K, float* gu, float* gx, void* stream) {
  EKL_REQUIRE(ekl_caps_supported(I, K, O, 1), "caps_agree: unsupported shape I=%d K=%d", I, K);
  EKL_CAPS_K(K, (caps_agree_bwd_kernel<KK><<<B, AGT, 0, (cudaStream_t)stream>>>(x, u, M, Z, gy, I, O, gu, gx)));
  EKL_LAUNCH_CHECK();
  return 0;
}
