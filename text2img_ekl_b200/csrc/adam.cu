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

__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, uint2* __restrict__ shadow, int64_t n4,
                                                        const float* __restrict__ state, float lr, float beta1, float beta2,
                                                        float eps) {
  const float step_size = lr / state[1];
  const float inv_bc2 = 1.f / state[2];
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
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

}  // namespace

// n must be a multiple of 4 and all buffers 16-byte aligned.  state: 3 device floats {step, 1-b1^t, sqrt(1-b2^t)},
// zero-initialised by the caller; every call advances the step (device side, so the call is CUDA-graph capturable).
extern "C" int ekl_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float* state,
                             float lr, float beta1, float beta2, float eps, void* stream) {
  EKL_REQUIRE(n % 4 == 0 && n > 0, "adam_step: n %% 4");
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, st>>>(state, beta1, beta2);
  EKL_LAUNCH_CHECK();
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_step_kernel<<<(int)blocks, 256, 0, st>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)shadow_bf16, n / 4,
                                               state, lr, beta1, beta2, eps);
  EKL_LAUNCH_CHECK();
  return 0;
}
