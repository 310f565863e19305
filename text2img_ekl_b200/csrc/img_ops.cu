// Image-side layout kernels of the discriminator stem (HBM-bound, fully coalesced).
//
// The first discriminator convolution is conv4x4 / stride 2 / pad 1 on a 3-channel fp32 NCHW image
// (model.py:832-836 encode_image_by_16times).  Space-to-depth by 2 turns it into a 3x3 / stride 1 convolution over
// 12 channels:   x'[b, i, j, (c*2 + ph)*2 + pw] = x[b, c, 2i+ph, 2j+pw]      (channels 12..15 are zero padding)
// so the layer runs on the tcgen05 implicit-GEMM kernel with one 16-channel K block per tap, reading a tensor
// 4x smaller than a channel-padded full-resolution image.  Up to three source images (the real / wrong / fake
// batches of a discriminator update, cub_trainer_splitz_cap_ca.py:418-420) are gathered in one pass: no torch.cat.
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

struct S2dSrc { const float* p[3]; };

// one thread per output pixel (b, i, j): 6 coalesced float2 loads (3 channels x 2 rows), 32 contiguous bytes stored
__global__ void __launch_bounds__(256) img_s2d_kernel(S2dSrc src, int Bg, int H, int W, int64_t npix, bf16* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo);
    int64_t r = idx / Wo;
    const int i = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const float* x = src.p[b / Bg] + (int64_t)(b % Bg) * 3 * H * W;
    float v[12];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const float2 t = *reinterpret_cast<const float2*>(x + ((int64_t)c * H + 2 * i + ph) * W + 2 * j);
        v[(c * 2 + ph) * 2] = t.x; v[(c * 2 + ph) * 2 + 1] = t.y;
      }
    uint4 lo = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    uint4 hi = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), 0u, 0u);
    uint4* dst = reinterpret_cast<uint4*>(out + idx * 16);
    dst[0] = lo; dst[1] = hi;
  }
}

// inverse (gradient of the fake image in the generator step): dx[b,c,2i+ph,2j+pw] = dx'[b,i,j,(c*2+ph)*2+pw]
__global__ void __launch_bounds__(256) img_s2d_bwd_kernel(const bf16* __restrict__ dxs, int H, int W, int64_t npix,
                                                          float* __restrict__ dx) {
  const int Ho = H / 2, Wo = W / 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo);
    int64_t r = idx / Wo;
    const int i = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const uint4* s = reinterpret_cast<const uint4*>(dxs + idx * 16);
    const uint4 lo = s[0], hi = s[1];
    const uint32_t w[6] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y};
    float* x = dx + (int64_t)b * 3 * H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const uint32_t u = w[c * 2 + ph];
        *reinterpret_cast<float2*>(x + ((int64_t)c * H + 2 * i + ph) * W + 2 * j) = make_float2(bf16_lo(u), bf16_hi(u));
      }
  }
}

// Image head (model.py:426-437 GET_IMAGE_G: conv3x3(ngf -> 3) + tanh): the conv kernel leaves the 3 real channels in
// a C-channel-padded NHWC bf16 tensor; this pass applies tanh in fp32 and writes the reference's NCHW fp32 image.
// One thread per pixel: 8-byte load, three coalesced 4-byte stores (one per colour plane).
__global__ void __launch_bounds__(256) head_tanh_fwd_kernel(const bf16* __restrict__ y, int C, int64_t HW, int64_t npix,
                                                            float* __restrict__ img) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += (int64_t)gridDim.x * blockDim.x) {
    const uint2 u = *reinterpret_cast<const uint2*>(y + idx * C);
    const int64_t b = idx / HW, r = idx - b * HW;
    float* o = img + b * 3 * HW + r;
    o[0] = tanhf(bf16_lo(u.x)); o[HW] = tanhf(bf16_hi(u.x)); o[2 * HW] = tanhf(bf16_lo(u.y));
  }
}

// dy[pixel, 0..2] = dimg * (1 - tanh(y)^2) (tanh recomputed from the pre-activation), padding channels zero
__global__ void __launch_bounds__(256) head_tanh_bwd_kernel(const bf16* __restrict__ y, const float* __restrict__ dimg, int C,
                                                            int64_t HW, int64_t npix, bf16* __restrict__ dy) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += (int64_t)gridDim.x * blockDim.x) {
    const uint2 u = *reinterpret_cast<const uint2*>(y + idx * C);
    const int64_t b = idx / HW, r = idx - b * HW;
    const float* d = dimg + b * 3 * HW + r;
    const float t0 = tanhf(bf16_lo(u.x)), t1 = tanhf(bf16_hi(u.x)), t2 = tanhf(bf16_lo(u.y));
    const float g0 = d[0] * (1.f - t0 * t0), g1 = d[HW] * (1.f - t1 * t1), g2 = d[2 * HW] * (1.f - t2 * t2);
    uint4* dst = reinterpret_cast<uint4*>(dy + idx * C);
    dst[0] = make_uint4(pack_bf16x2(g0, g1), pack_bf16x2(g2, 0.f), 0u, 0u);
    for (int k = 1; k < C / 8; ++k) dst[k] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Colour statistics of an image batch (cub_trainer_splitz_cap_ca.py:33-52 compute_mean_covariance): per image the
// channel mean mu[3] and the channel covariance cov[3][3] = (1/HW) sum_p (x_p - mu)(x_p - mu)^T.  HBM-bound: every
// pixel is read exactly once.  Pass 1 accumulates, per image, the first and second moments of y = x - s with the shift
// s = the image's first pixel (kills the cancellation of raw moments on nearly flat images): fp32 per thread (<= 64
// pixels), warp shuffles, then one fp64 atomic per block and moment.  acc[b] = {sum y_c (3), sum y_i y_j (i <= j: 6)}.
__global__ void __launch_bounds__(256) color_moments_kernel(const float* __restrict__ img, int HW, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const float* x = img + (int64_t)b * 3 * HW;
  const float s0 = x[0], s1 = x[HW], s2 = x[2 * (int64_t)HW];
  float m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = 0.f;
  const int n4 = HW / 4;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 c = reinterpret_cast<const float4*>(x + HW)[i];
    const float4 d = reinterpret_cast<const float4*>(x + 2 * (int64_t)HW)[i];
    const float r[4] = {a.x - s0, a.y - s0, a.z - s0, a.w - s0};
    const float g[4] = {c.x - s1, c.y - s1, c.z - s1, c.w - s1};
    const float u[4] = {d.x - s2, d.y - s2, d.z - s2, d.w - s2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m[0] += r[k]; m[1] += g[k]; m[2] += u[k];
      m[3] += r[k] * r[k]; m[4] += r[k] * g[k]; m[5] += r[k] * u[k];
      m[6] += g[k] * g[k]; m[7] += g[k] * u[k]; m[8] += u[k] * u[k];
    }
  }
  __shared__ float part[8][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v = warp_sum(m[k]);
    if (lane == 0) part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)part[w][threadIdx.x];
    atomicAdd(acc + (int64_t)b * 9 + threadIdx.x, t);
  }
}

// Pass 2 (one thread per image): mu = s + E[y], cov_ij = E[y_i y_j] - E[y_i] E[y_j], in fp64.
__global__ void color_finalize_kernel(const float* __restrict__ img, const double* __restrict__ acc, int B, int HW,
                                      float* __restrict__ mean, float* __restrict__ cov) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* x = img + (int64_t)b * 3 * HW;
  const double* a = acc + (int64_t)b * 9;
  const double inv = 1.0 / (double)HW;
  const double e[3] = {a[0] * inv, a[1] * inv, a[2] * inv};
  const double sft[3] = {(double)x[0], (double)x[HW], (double)x[2 * (int64_t)HW]};
  const int idx[3][3] = {{3, 4, 5}, {4, 6, 7}, {5, 7, 8}};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mean[b * 3 + i] = (float)(sft[i] + e[i]);
#pragma unroll
    for (int j = 0; j < 3; ++j) cov[b * 9 + i * 3 + j] = (float)(a[idx[i][j]] * inv - e[i] * e[j]);
  }
}

// Gradient: d mean_k / d x_kp = 1/HW;  d cov_ij / d x_kp = (delta_ik (x_jp - mu_j) + delta_jk (x_ip - mu_i)) / HW
// (the terms through mu vanish because sum_p (x_p - mu) = 0), so
//   dx[k][p] = ( dmean[k] + sum_j (dcov[k][j] + dcov[j][k]) (x[j][p] - mu[j]) ) / HW .      One read, one write per element.
__global__ void __launch_bounds__(256) color_stats_bwd_kernel(const float* __restrict__ img, const float* __restrict__ mean,
                                                              const float* __restrict__ dmean, const float* __restrict__ dcov,
                                                              int HW, float* __restrict__ dimg) {
  const int b = blockIdx.y;
  const float* x = img + (int64_t)b * 3 * HW;
  float* dx = dimg + (int64_t)b * 3 * HW;
  const float inv = 1.f / (float)HW;
  float mu[3], dm[3], sym[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mu[i] = mean[b * 3 + i];
    dm[i] = dmean ? dmean[b * 3 + i] * inv : 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) sym[i][j] = dcov ? (dcov[b * 9 + i * 3 + j] + dcov[b * 9 + j * 3 + i]) * inv : 0.f;
  }
  const int n4 = HW / 4;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 c = reinterpret_cast<const float4*>(x + HW)[i];
    const float4 d = reinterpret_cast<const float4*>(x + 2 * (int64_t)HW)[i];
    const float r[4] = {a.x - mu[0], a.y - mu[0], a.z - mu[0], a.w - mu[0]};
    const float g[4] = {c.x - mu[1], c.y - mu[1], c.z - mu[1], c.w - mu[1]};
    const float u[4] = {d.x - mu[2], d.y - mu[2], d.z - mu[2], d.w - mu[2]};
    float o[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) o[k][q] = dm[k] + sym[k][0] * r[q] + sym[k][1] * g[q] + sym[k][2] * u[q];
    reinterpret_cast<float4*>(dx)[i] = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
    reinterpret_cast<float4*>(dx + HW)[i] = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
    reinterpret_cast<float4*>(dx + 2 * (int64_t)HW)[i] = make_float4(o[2][0], o[2][1], o[2][2], o[2][3]);
  }
}

int color_slices(int HW) {
  int s = HW / 4096;          // >= 4 float4 loads per thread and channel
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}

}  // namespace

extern "C" int ekl_color_stats_fwd(const float* img, int B, int HW, double* scratch, float* mean, float* cov, void* stream) {
  EKL_REQUIRE(img && scratch && mean && cov, "color_stats_fwd: null pointer");
  EKL_REQUIRE(B > 0 && B <= 65535 && HW >= 4 && HW % 4 == 0, "color_stats_fwd: HW must be a positive multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 9 * (size_t)B, st);
  if (e != cudaSuccess) return (int)e;
  color_moments_kernel<<<dim3(color_slices(HW), B), 256, 0, st>>>(img, HW, scratch);
  EKL_LAUNCH_CHECK();
  color_finalize_kernel<<<ekl_cdiv(B, 64), 64, 0, st>>>(img, scratch, B, HW, mean, cov);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_color_stats_bwd(const float* img, const float* mean, const float* dmean, const float* dcov, int B, int HW,
                                   float* dimg, void* stream) {
  EKL_REQUIRE(img && mean && dimg, "color_stats_bwd: null pointer");
  EKL_REQUIRE(B > 0 && B <= 65535 && HW >= 4 && HW % 4 == 0, "color_stats_bwd: HW must be a positive multiple of 4");
  color_stats_bwd_kernel<<<dim3(color_slices(HW), B), 256, 0, (cudaStream_t)stream>>>(img, mean, dmean, dcov, HW, dimg);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_head_tanh_fwd(const void* y, int B, int HW, int C, float* img, void* stream) {
  EKL_REQUIRE(C % 8 == 0 && C >= 8 && B > 0 && HW > 0, "head_tanh_fwd: C %% 8");
  const int64_t npix = (int64_t)B * HW;
  int blocks = (int)((npix + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  head_tanh_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, C, HW, npix, img);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_head_tanh_bwd(const void* y, const float* dimg, int B, int HW, int C, void* dy, void* stream) {
  EKL_REQUIRE(C % 8 == 0 && C >= 8 && B > 0 && HW > 0, "head_tanh_bwd: C %% 8");
  const int64_t npix = (int64_t)B * HW;
  int blocks = (int)((npix + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  head_tanh_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, dimg, C, HW, npix, (bf16*)dy);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_img_s2d(const float* x0, const float* x1, const float* x2, int groups, int B, int H, int W, void* out,
                           void* stream) {
  EKL_REQUIRE(groups >= 1 && groups <= 3 && H % 2 == 0 && W % 2 == 0 && B > 0, "img_s2d: bad arguments");
  S2dSrc s;
  s.p[0] = x0; s.p[1] = x1; s.p[2] = x2;
  const int64_t npix = (int64_t)groups * B * (H / 2) * (W / 2);
  int blocks = (int)((npix + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  img_s2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(s, B, H, W, npix, (bf16*)out);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_img_s2d_bwd(const void* dxs, int B, int H, int W, float* dx, void* stream) {
  EKL_REQUIRE(H % 2 == 0 && W % 2 == 0 && B > 0, "img_s2d_bwd: bad arguments");
  const int64_t npix = (int64_t)B * (H / 2) * (W / 2);
  int blocks = (int)((npix + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  img_s2d_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dxs, H, W, npix, dx);
  EKL_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- loader image pyramid (datasets.py:43-68 get_imgs)
// The reference resizes the final-size uint8 crop to every lower stage size with PIL's BILINEAR resize
// (transforms.Scale) and normalises each level (ToTensor + Normalize(0.5, 0.5)).  PIL's 8-bit resampler
// (libImaging/Resample.c) is integer arithmetic: per output pixel a window [xmin, xmin+n) of fixed-point weights with
// 22 fractional bits, accumulator started at 1 << 21, result clip8(acc >> 22); horizontal pass, then vertical pass on the
// uint8 intermediate.  The weight tables are computed on the host exactly as PIL does (double arithmetic) and passed in,
// so the kernels are bit-exact.  Layout: uint8 [B][S][S][3] (PIL's HWC) in, fp32 NCHW [-1, 1] out.
namespace {

constexpr int PIL_PRECISION_BITS = 22;

// horizontal: src u8 [rows][S][3] -> dst u8 [rows][s][3]; one thread per (row, xx), 3 channels
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ src, int64_t rows, int S, int s,
                                                         const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                         uint8_t* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < rows * s; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / s;
    const int xx = (int)(i - r * s);
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int* k = kk + (size_t)xx * ksize;
    int a0 = 1 << (PIL_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    const uint8_t* p = src + (r * S + xmin) * 3;
    for (int x = 0; x < n; ++x) {
      const int w = k[x];
      a0 += p[3 * x] * w; a1 += p[3 * x + 1] * w; a2 += p[3 * x + 2] * w;
    }
    uint8_t* o = dst + i * 3;
    o[0] = (uint8_t)min(max(a0 >> PIL_PRECISION_BITS, 0), 255);
    o[1] = (uint8_t)min(max(a1 >> PIL_PRECISION_BITS, 0), 255);
    o[2] = (uint8_t)min(max(a2 >> PIL_PRECISION_BITS, 0), 255);
  }
}

__device__ __forceinline__ float to_norm(int u) {          // ToTensor (u / 255) then Normalize ((x - 0.5) / 0.5), fp32 like torch
  const float t = __fdiv_rn((float)u, 255.0f);
  return __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
}

// vertical + normalise: src u8 [B][S][s][3] -> dst fp32 [B][3][s][s]; one thread per (b, yy, xx)
__global__ void __launch_bounds__(256) resample_v_norm_kernel(const uint8_t* __restrict__ src, int B, int S, int s,
                                                              const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                              float* __restrict__ dst) {
  const int64_t total = (int64_t)B * s * s;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xx = (int)(i % s);
    const int yy = (int)((i / s) % s);
    const int b = (int)(i / ((int64_t)s * s));
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = kk + (size_t)yy * ksize;
    int a0 = 1 << (PIL_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    const uint8_t* p = src + (((int64_t)b * S + ymin) * s + xx) * 3;
    for (int y = 0; y < n; ++y) {
      const int w = k[y];
      a0 += p[(int64_t)y * s * 3] * w; a1 += p[(int64_t)y * s * 3 + 1] * w; a2 += p[(int64_t)y * s * 3 + 2] * w;
    }
    const int64_t plane = (int64_t)s * s;
    float* o = dst + (int64_t)b * 3 * plane + (int64_t)yy * s + xx;
    o[0] = to_norm(min(max(a0 >> PIL_PRECISION_BITS, 0), 255));
    o[plane] = to_norm(min(max(a1 >> PIL_PRECISION_BITS, 0), 255));
    o[2 * plane] = to_norm(min(max(a2 >> PIL_PRECISION_BITS, 0), 255));
  }
}

// the final-size level: uint8 HWC -> fp32 NCHW, normalised
__global__ void __launch_bounds__(256) u8_norm_kernel(const uint8_t* __restrict__ src, int B, int S, float* __restrict__ dst) {
  const int64_t plane = (int64_t)S * S, total = (int64_t)B * plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / plane, px = i - b * plane;
    const uint8_t* p = src + i * 3;
    float* o = dst + b * 3 * plane + px;
    o[0] = to_norm(p[0]); o[plane] = to_norm(p[1]); o[2 * plane] = to_norm(p[2]);
  }
}

}  // namespace

// One pyramid level.  s == S: plain conversion (tmp / tables unused).  Otherwise bounds [s][2] (xmin, n) and kk [s][ksize]
// are the PIL coefficient tables of the resize S -> s (device int32; the same tables serve both passes of a square
// resize), tmp is caller scratch of B*S*s*3 bytes.
extern "C" int ekl_img_pyramid_level(const void* src_u8, int B, int S, int s, const int* bounds, const int* kk, int ksize,
                                     void* tmp_u8, float* out, void* stream) {
  EKL_REQUIRE(src_u8 != nullptr && out != nullptr && B > 0 && S > 0 && s > 0, "img_pyramid_level: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (s == S) {
    int64_t total = (int64_t)B * S * S;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    u8_norm_kernel<<<blocks, 256, 0, st>>>((const uint8_t*)src_u8, B, S, out);
    EKL_LAUNCH_CHECK();
    return 0;
  }
  EKL_REQUIRE(bounds != nullptr && kk != nullptr && tmp_u8 != nullptr && ksize > 0, "img_pyramid_level: coefficient tables / scratch missing");
  {
    const int64_t rows = (int64_t)B * S, total = rows * s;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    resample_h_kernel<<<blocks, 256, 0, st>>>((const uint8_t*)src_u8, rows, S, s, bounds, kk, ksize, (uint8_t*)tmp_u8);
    EKL_LAUNCH_CHECK();
  }
  {
    const int64_t total = (int64_t)B * s * s;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    resample_v_norm_kernel<<<blocks, 256, 0, st>>>((const uint8_t*)tmp_u8, B, S, s, bounds, kk, ksize, out);
    EKL_LAUNCH_CHECK();
  }
  return 0;
}
