// Folded condition-code channels of the generator's jointConv (model.py:403, 411-414): the reference tiles c_code over the
// H x W map and concatenates it in front of h_code; a 3x3 / pad-1 convolution over spatially constant channels is a
// per-sample bias that only depends on which taps fall inside the map, i.e. on the pixel's border class
// q = 3*rc + cc (rc / cc = 0 first row / column, 1 interior, 2 last):
//     bias9[b,q,n] = sum_t VALID(q,t) * T[b,t,n],   T[b,t,n] = sum_{c < ef} code[b,c] * W[n][t][c]
// (W: the fp32 master filter [N][3][3][Ctot] in channels_last storage, code channels first).  Three small kernels
// replace ~50 launches of einsum / slicing / reduction glue per generator stage:
//   code_bias9_fwd      code, W -> bias9                               (input of ekl_conv_fwd_bias9)
//   border_sums9        dy [B,H,W,N] bf16 -> S[b,q,n] = sum of dy over the pixels of class q (one pass over dy)
//   code_bias9_bwd      S, code, W -> dcode[b,c], dW[n][t][c] (+=)     (dT[b,t,n] = sum_q VALID(q,t) S[b,q,n])
// HBM-bound (border_sums9 reads dy once) / latency-bound (the two code kernels).
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

// tap t = 3*kh + kw of a 3x3 / pad-1 conv reads inside the map for a pixel of border class q = 3*rc + cc
__device__ __forceinline__ bool tap_valid(int q, int t) {
  const int rc = q / 3, cc = q - 3 * rc, kh = t / 3, kw = t - 3 * kh;
  return !((rc == 0 && kh == 0) || (rc == 2 && kh == 2) || (cc == 0 && kw == 0) || (cc == 2 && kw == 2));
}

// grid = N blocks (one output channel each), 256 threads.  smem: W[n] code slice [9][ef], T [B][9]
__global__ void __launch_bounds__(256) code_bias9_fwd_kernel(const float* __restrict__ code, const float* __restrict__ w, int B, int ef,
                                                             int Ctot, int N, float* __restrict__ bias9) {
  extern __shared__ float sm[];
  float* wn = sm;                 // [9][ef]
  float* T = sm + 9 * ef;         // [B][9]
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < 9 * ef; i += 256) {
    const int t = i / ef, c = i - t * ef;
    wn[i] = w[((size_t)n * 9 + t) * Ctot + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * 9; i += 256) {
    const int b = i / 9, t = i - 9 * b;
    const float* cb = code + (size_t)b * ef;
    const float* wt = wn + t * ef;
    float acc = 0.f;
    for (int c = 0; c < ef; ++c) acc += cb[c] * wt[c];
    T[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * 9; i += 256) {
    const int b = i / 9, q = i - 9 * b;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc += tap_valid(q, t) ? T[b * 9 + t] : 0.f;
    bias9[((size_t)b * 9 + q) * N + n] = acc;
  }
}

// S[b][q][n] += sum over the pixels of border class q of dy[b,h,w,n].  grid = (row chunks, B): chunk 0 = image row 0,
// the last chunk = row H-1, the chunks between cover the interior rows R at a time, so a block's row class is uniform.
// block = 256 threads = (N/8 channel octets) x (column lanes); S is fp32, ZERO on entry.
__global__ void __launch_bounds__(256) border_sums9_kernel(const bf16* __restrict__ dy, int H, int W, int N, int R,
                                                           float* __restrict__ S) {
  __shared__ float red[256 * 24];
  const int b = blockIdx.y;
  const int nchunks = gridDim.x;
  int r0, r1, rc;
  if (blockIdx.x == 0) { r0 = 0; r1 = 1; rc = 0; }
  else if ((int)blockIdx.x == nchunks - 1) { r0 = H - 1; r1 = H; rc = 2; }
  else { r0 = 1 + ((int)blockIdx.x - 1) * R; r1 = r0 + R < H - 1 ? r0 + R : H - 1; rc = 1; }
  const int noct = N / 8;
  const int lanes = 256 / noct;
  const int oct = threadIdx.x % noct, lane = threadIdx.x / noct;
  float acc[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  if (lane < lanes) {
    for (int r = r0; r < r1; ++r) {
      const bf16* row = dy + (((size_t)b * H + r) * W) * N + oct * 8;
      for (int x = lane; x < W; x += lanes) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + (size_t)x * N);
        const int cc = x == 0 ? 0 : (x == W - 1 ? 2 : 1);
        const float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k == cc) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[k][i] += f[i];
          }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 24 + k * 8 + i] = acc[k][i];
  __syncthreads();
  // thread (oct, lane 0..2) sums column class `lane` over the column lanes and adds it to S
  if (lane < 3) {
    const int k = lane;
    float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int i = 0; i < 8; ++i) tot[i] += red[(l * noct + oct) * 24 + k * 8 + i];
    float* dst = S + ((size_t)b * 9 + rc * 3 + k) * N + oct * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dst + i, tot[i]);
  }
}

// grid = N blocks (one output channel each), 256 threads.  dcode: fp32 [B][ef], ZERO on entry (atomics over n);
// dw: master gradient [N][9][Ctot], code channels [0, ef) accumulated (+=), each element by exactly one thread.
__global__ void __launch_bounds__(256) code_bias9_bwd_kernel(const float* __restrict__ S, const float* __restrict__ code,
                                                             const float* __restrict__ w, int B, int ef, int Ctot, int N,
                                                             float* __restrict__ dcode, float* dw) {
  extern __shared__ float sm[];
  float* dT = sm;                 // [B][9]
  float* wn = sm + B * 9;         // [9][ef]
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < B * 9; i += 256) {
    const int b = i / 9, t = i - 9 * b;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) acc += tap_valid(q, t) ? S[((size_t)b * 9 + q) * N + n] : 0.f;
    dT[i] = acc;
  }
  if (dcode != nullptr)
    for (int i = threadIdx.x; i < 9 * ef; i += 256) {
      const int t = i / ef, c = i - t * ef;
      wn[i] = w[((size_t)n * 9 + t) * Ctot + c];
    }
  __syncthreads();
  if (dw != nullptr)
    for (int i = threadIdx.x; i < 9 * ef; i += 256) {
      const int t = i / ef, c = i - t * ef;
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc += dT[b * 9 + t] * code[(size_t)b * ef + c];
      dw[((size_t)n * 9 + t) * Ctot + c] += acc;
    }
  if (dcode != nullptr)
    for (int i = threadIdx.x; i < B * ef; i += 256) {
      const int b = i / ef, c = i - b * ef;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc += dT[b * 9 + t] * wn[t * ef + c];
      atomicAdd(dcode + i, acc);
    }
}

}  // namespace

extern "C" int ekl_code_bias9_fwd(const float* code, const float* w, int B, int ef, int Ctot, int N, float* bias9, void* stream) {
  EKL_REQUIRE(code != nullptr && w != nullptr && bias9 != nullptr, "code_bias9_fwd: null pointer argument");
  EKL_REQUIRE(B > 0 && ef > 0 && ef <= Ctot && N > 0, "code_bias9_fwd: bad shape");
  const size_t smem = (size_t)(9 * ef + 9 * B) * sizeof(float);
  EKL_REQUIRE(smem <= 48 * 1024, "code_bias9_fwd: ef / batch too large for one block (%d, %d)", ef, B);
  code_bias9_fwd_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(code, w, B, ef, Ctot, N, bias9);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_border_sums9(const void* dy, int B, int H, int W, int N, float* S, void* stream) {
  EKL_REQUIRE(dy != nullptr && S != nullptr, "border_sums9: null pointer argument");
  EKL_REQUIRE(B > 0 && H >= 2 && W >= 2 && N % 8 == 0 && N / 8 <= 64, "border_sums9: H, W >= 2, N %% 8 == 0, N <= 512 (N=%d)", N);
  const int R = 8;
  const int chunks = 2 + (H > 2 ? ekl_cdiv(H - 2, R) : 0);
  border_sums9_kernel<<<dim3(chunks, B), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, H, W, N, R, S);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_code_bias9_bwd(const float* S, const float* code, const float* w, int B, int ef, int Ctot, int N, float* dcode,
                                  float* dw, void* stream) {
  EKL_REQUIRE(S != nullptr && code != nullptr && w != nullptr, "code_bias9_bwd: null pointer argument");
  EKL_REQUIRE(B > 0 && ef > 0 && ef <= Ctot && N > 0, "code_bias9_bwd: bad shape");
  const size_t smem = (size_t)(9 * ef + 9 * B) * sizeof(float);
  EKL_REQUIRE(smem <= 48 * 1024, "code_bias9_bwd: ef / batch too large for one block (%d, %d)", ef, B);
  code_bias9_bwd_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(S, code, w, B, ef, Ctot, N, dcode, dw);
  EKL_LAUNCH_CHECK();
  return 0;
}
