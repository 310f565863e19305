// VC_NET's hidden layers (model.py:160-201): h = ReLU(BatchNorm1d(x W^T + b)) over a batch of 24..64 rows, forward and
// backward as ONE kernel each (the reference's F.linear + batch_norm + relu is ~6 launches forward and ~10 backward on
// [B, <= 1325] tensors: pure launch latency at the head of the generator's critical path).
//
// The batch is tiny, so a WARP owns an output column n for ALL rows: lanes split the K dimension (coalesced reads of
// W[n][:]), x is staged through shared memory in K chunks, the B dot products are combined with a butterfly so that
// every lane ends up with all B pre-activations, and the BatchNorm statistics over the batch (per column = per warp)
// never leave registers.  fp32 throughout (the reference runs these layers in fp32).
//   forward  : y = x W^T + b ; xhat = (y - mean_B) * rstd_B ; h = max(gamma * xhat + beta, 0)      (+ running stats)
//   backward : dz = dh * [h > 0] ; dbeta = sum dz ; dgamma = sum dz * xhat ;
//              dy = gamma * rstd * (dz - dbeta/B - xhat * dgamma/B) ; dW[n][:] += dy^T x ; db[n] += sum dy
//              (dx = dy W is left to one library GEMM: only the second layer needs it)
// Latency-bound; bytes moved = W once per direction.
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

constexpr int KC = 128;      // K chunk staged in shared memory
constexpr int WARPS = 8;     // columns per block
constexpr int KSPLIT = 256;  // backward: K range of one block row (multiple of KC)

// Forward K chunk: KCF columns of x for all B rows live in (dynamic) shared memory at a time.  The chunk loop is a latency
// chain (stage x -> barrier -> read W -> multiply), so chunks are large (3 for K = 1225 instead of 10 of 128) and the
// warp's W values of a chunk are requested BEFORE the staging barrier: both global round trips of a chunk overlap.
constexpr int KCF = 512;

template <int BMAX>
__global__ void __launch_bounds__(32 * WARPS) linear_bn_relu_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* running_mean, float* running_var, int B, int K, int N, float eps, float momentum,
    int training, float* __restrict__ h, float* __restrict__ xhat, float* __restrict__ rstd_out) {
  extern __shared__ float xs[];          // [B][KCF]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * WARPS + warp;
  float acc[BMAX];
#pragma unroll
  for (int b = 0; b < BMAX; ++b) acc[b] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KCF) {
    float w[KCF / 32];
#pragma unroll
    for (int j = 0; j < KCF / 32; ++j) {
      const int k = k0 + j * 32 + lane;
      w[j] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
    }
    __syncthreads();                      // the previous chunk has been consumed
    for (int i = threadIdx.x; i < B * KCF; i += 32 * WARPS) {
      const int b = i / KCF, k = i - b * KCF;
      xs[i] = (k0 + k) < K ? x[(size_t)b * K + k0 + k] : 0.f;
    }
    __syncthreads();
    if (n < N) {
#pragma unroll
      for (int j = 0; j < KCF / 32; ++j) {
        const int k = j * 32 + lane;
#pragma unroll
        for (int b = 0; b < BMAX; ++b)
          if (b < B) acc[b] = fmaf(xs[b * KCF + k], w[j], acc[b]);
      }
    }
  }
  if (n >= N) return;
  // butterfly: every lane ends with the full dot products
#pragma unroll
  for (int b = 0; b < BMAX; ++b)
    if (b < B) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
      acc[b] += bias != nullptr ? bias[n] : 0.f;
    }
  float mean, var;
  if (training) {
    float s = 0.f;
#pragma unroll
    for (int b = 0; b < BMAX; ++b) if (b < B) s += acc[b];
    mean = s / (float)B;
    float q = 0.f;
#pragma unroll
    for (int b = 0; b < BMAX; ++b) if (b < B) { const float d = acc[b] - mean; q += d * d; }
    var = q / (float)B;
    if (lane == 0 && running_mean != nullptr) {
      const float unb = B > 1 ? var * (float)B / (float)(B - 1) : var;
      running_mean[n] = (1.f - momentum) * running_mean[n] + momentum * mean;
      running_var[n] = (1.f - momentum) * running_var[n] + momentum * unb;
    }
  } else {
    mean = running_mean[n]; var = running_var[n];
  }
  const float r = 1.f / sqrtf(var + eps);
  const float g = gamma[n], be = beta[n];
  if (lane == 0 && rstd_out != nullptr) rstd_out[n] = r;
#pragma unroll
  for (int b = 0; b < BMAX; ++b)
    if (b < B && (b & 31) == lane) {
      const float xh = (acc[b] - mean) * r;
      if (xhat != nullptr) xhat[(size_t)b * N + n] = xh;
      h[(size_t)b * N + n] = fmaxf(g * xh + be, 0.f);
    }
}

template <int BMAX>
__global__ void __launch_bounds__(32 * WARPS) linear_bn_relu_bwd_kernel(
    const float* __restrict__ dh, const float* __restrict__ h, const float* __restrict__ xhat, const float* __restrict__ rstd,
    const float* __restrict__ gamma, const float* __restrict__ x, int B, int K, int N, float* dW, float* dbias, float* dgamma,
    float* dbeta, float* __restrict__ dy_out) {
  __shared__ float xs[BMAX * KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * WARPS + warp;
  float dy[BMAX];
  if (n < N) {
    float sb = 0.f, sg = 0.f;
    float dz[BMAX], xh[BMAX];
#pragma unroll
    for (int b = 0; b < BMAX; ++b)
      if (b < B) {
        const float hv = h[(size_t)b * N + n];
        dz[b] = hv > 0.f ? dh[(size_t)b * N + n] : 0.f;
        xh[b] = xhat[(size_t)b * N + n];
        sb += dz[b]; sg += dz[b] * xh[b];
      }
    const float k = gamma[n] * rstd[n], inv = 1.f / (float)B;
    float sdy = 0.f;
#pragma unroll
    for (int b = 0; b < BMAX; ++b)
      if (b < B) {
        dy[b] = k * (dz[b] - sb * inv - xh[b] * sg * inv);
        sdy += dy[b];
        if ((b & 31) == lane && blockIdx.y == 0) dy_out[(size_t)b * N + n] = dy[b];
      }
    if (lane == 0 && blockIdx.y == 0) {
      if (dgamma != nullptr) dgamma[n] += sg;
      if (dbeta != nullptr) dbeta[n] += sb;
      if (dbias != nullptr) dbias[n] += sdy;
    }
  }
  if (dW == nullptr) return;
  // the K range of this block row: the weight gradient is split along K over gridDim.y block rows (each dW element has
  // exactly one owner, so no atomics); the short prologue above is repeated per row, its outputs written by row 0
  const int kbeg = blockIdx.y * KSPLIT, kend = kbeg + KSPLIT < K ? kbeg + KSPLIT : K;
  for (int k0 = kbeg; k0 < kend; k0 += KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < B * KC; i += 32 * WARPS) {
      const int b = i / KC, kk = i - b * KC;
      xs[b * KC + kk] = (k0 + kk) < K ? x[(size_t)b * K + k0 + kk] : 0.f;
    }
    __syncthreads();
    if (n < N) {
#pragma unroll
      for (int j = 0; j < KC / 32; ++j) {
        const int kk = j * 32 + lane;
        if (k0 + kk < K) {
          float a = 0.f;
#pragma unroll
          for (int b = 0; b < BMAX; ++b)
            if (b < B) a += dy[b] * xs[b * KC + kk];
          dW[(size_t)n * K + k0 + kk] += a;
        }
      }
    }
  }
}

}  // namespace

extern "C" int ekl_linear_bn_relu_fwd(const float* x, const float* W, const float* bias, const float* gamma, const float* beta,
                                      float* running_mean, float* running_var, int B, int K, int N, float eps, float momentum,
                                      int training, float* h, float* xhat, float* rstd, void* stream) {
  EKL_REQUIRE(x != nullptr && W != nullptr && gamma != nullptr && beta != nullptr && h != nullptr, "linear_bn_relu_fwd: null pointer argument");
  EKL_REQUIRE(B >= 1 && B <= 64 && K > 0 && N > 0, "linear_bn_relu_fwd: batch 1..64 (got %d)", B);
  EKL_REQUIRE(training || (running_mean != nullptr && running_var != nullptr), "linear_bn_relu_fwd: inference needs running statistics");
  const int grid = ekl_cdiv(N, WARPS);
  const size_t smem = (size_t)B * KCF * sizeof(float);           // <= 128 KB at B = 64
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(linear_bn_relu_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * KCF * 4));
    EKL_CHECK_CUDA(cudaFuncSetAttribute(linear_bn_relu_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * KCF * 4));
    attr_done = true;
  }
  if (B <= 32)
    linear_bn_relu_fwd_kernel<32><<<grid, 32 * WARPS, smem, (cudaStream_t)stream>>>(x, W, bias, gamma, beta, running_mean, running_var, B,
                                                                                   K, N, eps, momentum, training, h, xhat, rstd);
  else
    linear_bn_relu_fwd_kernel<64><<<grid, 32 * WARPS, smem, (cudaStream_t)stream>>>(x, W, bias, gamma, beta, running_mean, running_var, B,
                                                                                   K, N, eps, momentum, training, h, xhat, rstd);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_linear_bn_relu_bwd(const float* dh, const float* h, const float* xhat, const float* rstd, const float* gamma,
                                      const float* x, int B, int K, int N, float* dW, float* dbias, float* dgamma, float* dbeta,
                                      float* dy, void* stream) {
  EKL_REQUIRE(dh != nullptr && h != nullptr && xhat != nullptr && rstd != nullptr && gamma != nullptr && x != nullptr && dy != nullptr,
              "linear_bn_relu_bwd: null pointer argument");
  EKL_REQUIRE(B >= 1 && B <= 64 && K > 0 && N > 0, "linear_bn_relu_bwd: batch 1..64 (got %d)", B);
  const dim3 grid(ekl_cdiv(N, WARPS), dW != nullptr ? ekl_cdiv(K, KSPLIT) : 1);
  if (B <= 32)
    linear_bn_relu_bwd_kernel<32><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(dh, h, xhat, rstd, gamma, x, B, K, N, dW, dbias, dgamma, dbeta, dy);
  else
    linear_bn_relu_bwd_kernel<64><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(dh, h, xhat, rstd, gamma, x, B, K, N, dW, dbias, dgamma, dbeta, dy);
  EKL_LAUNCH_CHECK();
  return 0;
}
