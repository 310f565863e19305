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

}  // namespace

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
