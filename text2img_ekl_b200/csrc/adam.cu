// Fused Adam over a network's FLAT parameter / gradient / moment buffers (torch.optim.Adam semantics as the reference
// configures it: cub_trainer_splitz_cap_ca.py:199-215, lr 2e-4, betas (0.5, 0.999), eps 1e-8, no weight decay), one
// elementwise pass per network instead of a multi-tensor launch per parameter group, which also emits the bf16 shadow
// of the updated parameters: for stride-1 / stride-2 convolutions stored channels_last that shadow IS the packed forward
// filter operand ([Cout][tap][Cin]), so no separate pack pass reads the fp32 master again.
// HBM-bound: 16 B read + 14 B written per parameter.
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

// state[0] = step count, state[1] = 1 - beta1^t, state[2] = sqrt(1 - beta2^t)
__global__ void adam_tick_kernel(float* state, float beta1, float beta2) {
  const double t = (double)state[0] + 1.0;
  state[0] = (float)t;
  state[1] = (float)(1.0 - pow((double)beta1, t));
  state[2] = (float)sqrt(1.0 - pow((double)beta2, t));
}

__device__ __forceinline__ float4 load_grad4(const float4* g, int64_t i) { return g[i]; }
__device__ __forceinline__ float4 load_grad4(const uint2* g, int64_t i) {      // 4 bf16 gradients (the all-reduced staging buffer)
  const uint2 r = g[i];
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}

template <typename G4>
__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ p, const G4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, uint2* __restrict__ shadow, int64_t n4,
                                                        const float* __restrict__ state, float lr, float beta1, float beta2,
                                                        float eps) {
  const float step_size = lr / state[1];
  const float inv_bc2 = 1.f / state[2];
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = load_grad4(g, i);
#define EKL_ADAM1(c)                                                   \
    mm.c = beta1 * mm.c + (1.f - beta1) * gg.c;                        \
    vv.c = beta2 * vv.c + (1.f - beta2) * gg.c * gg.c;                 \
    pp.c -= step_size * mm.c / (sqrtf(vv.c) * inv_bc2 + eps);
    EKL_ADAM1(x) EKL_ADAM1(y) EKL_ADAM1(z) EKL_ADAM1(w)
#undef EKL_ADAM1
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (shadow != nullptr) shadow[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
  }
}

// fp32 -> bf16 (round to nearest even) of a gradient slice: the NVLink payload of the data-parallel gradient average
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 a = src[i];
    dst[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
  }
}

int adam_launch(float* p, const void* g, int g_bf16, float* m, float* v, void* shadow_bf16, int64_t n, float* state, float lr,
                float beta1, float beta2, float eps, cudaStream_t st, bool tick) {
  EKL_REQUIRE(n % 4 == 0 && n > 0, "adam_step: n %% 4");
  EKL_REQUIRE((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)g & (g_bf16 ? 7 : 15)) == 0 &&
                  ((uintptr_t)shadow_bf16 & 7) == 0, "adam_step: buffers must be 16-byte aligned (bf16 buffers: 8)");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, st>>>(state, beta1, beta2);
    EKL_LAUNCH_CHECK();
  }
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (g_bf16)
    adam_step_kernel<uint2><<<(int)blocks, 256, 0, st>>>((float4*)p, (const uint2*)g, (float4*)m, (float4*)v, (uint2*)shadow_bf16,
                                                         n / 4, state, lr, beta1, beta2, eps);
  else
    adam_step_kernel<float4><<<(int)blocks, 256, 0, st>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v,
                                                          (uint2*)shadow_bf16, n / 4, state, lr, beta1, beta2, eps);
  EKL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// n must be a multiple of 4 and all buffers 16-byte aligned.  state: 3 device floats {step, 1-b1^t, sqrt(1-b2^t)},
// zero-initialised by the caller; every call advances the step (device side, so the call is CUDA-graph capturable).
extern "C" int ekl_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float* state,
                             float lr, float beta1, float beta2, float eps, void* stream) {
  return adam_launch(p, g, 0, m, v, shadow_bf16, n, state, lr, beta1, beta2, eps, (cudaStream_t)stream, true);
}

// Same update with the gradient read from a bf16 buffer: the data-parallel path all-reduces gradients as bf16 (half the
// NVLink bytes of the reference's fp32 DataParallel reduce) and the optimiser consumes that staging buffer directly.
extern "C" int ekl_adam_step_g16(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, int64_t n, float* state,
                                 float lr, float beta1, float beta2, float eps, void* stream) {
  return adam_launch(p, g_bf16, 1, m, v, shadow_bf16, n, state, lr, beta1, beta2, eps, (cudaStream_t)stream, true);
}

// The same update in pieces, for an optimiser step that overlaps the backward pass: ekl_adam_tick advances the step count
// (and both bias corrections) ONCE, then ekl_adam_apply updates any slice [p, p + n) of the flat buffers -- as soon as that
// slice's gradients are final -- without touching the count.  g_is_bf16: the gradient slice is bf16 (all-reduce staging).
extern "C" int ekl_adam_tick(float* state, float beta1, float beta2, void* stream) {
  EKL_REQUIRE(state != nullptr, "adam_tick: null state");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, beta1, beta2);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_adam_apply(float* p, const void* g, int g_is_bf16, float* m, float* v, void* shadow_bf16, int64_t n,
                              const float* state, float lr, float beta1, float beta2, float eps, void* stream) {
  return adam_launch(p, g, g_is_bf16, m, v, shadow_bf16, n, const_cast<float*>(state), lr, beta1, beta2, eps, (cudaStream_t)stream, false);
}

// dst[i] = bf16(src[i]), n % 4 == 0, both 16 / 8-byte aligned: gradient slice -> NVLink payload
extern "C" int ekl_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  EKL_REQUIRE(n % 4 == 0 && n > 0, "cast_bf16: n %% 4");
  EKL_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_bf16 & 7) == 0, "cast_bf16: alignment");
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint2*)dst_bf16, n / 4);
  EKL_LAUNCH_CHECK();
  return 0;
}
