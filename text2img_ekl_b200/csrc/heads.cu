// Discriminator logit heads and the multi-branch GAN losses as a handful of fused kernels (HBM / latency bound,
// warp-shuffle reductions) instead of ~100 elementwise / reduction launches per discriminator pass.
//
//   heads   : model.py:886-888, 935-952  `logits` / `uncond_logits` = Conv2d(8ndf, 1, 4, stride 4) + Sigmoid on the 4x4
//             code map == one dot of length 16*8ndf per sample; computed here as raw (pre-sigmoid) logits.
//   losses  : cub_trainer_splitz_cap_ca.py:423-448 (train_joint_Dnet) and :470-487 (loss_joint_Gnet):
//             nn.BCELoss against constant 0/1 labels on the match / uncond probabilities of every group
//             (real, wrong, fake), + ce_loss(log_softmax(class logits), soft target) (cub:60-65), torch's
//             log clamp at -100 included.
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = l < 8 ? sh[l] : 0.f;
  return warp_sum(t);      // every warp ends with the full sum
}

// grid = GB blocks of 256 threads; K % 8 == 0.  feat row (bf16) . w (fp32, same element order) + bias
__global__ void __launch_bounds__(256) dhead_dots_kernel(const bf16* __restrict__ x, const bf16* __restrict__ h, const float* __restrict__ wu,
                                                         const float* __restrict__ bu, const float* __restrict__ wm,
                                                         const float* __restrict__ bm, int K, float* __restrict__ lu,
                                                         float* __restrict__ lm) {
  __shared__ float sh[8];
  const int b = blockIdx.x;
  float au = 0.f, am = 0.f;
  for (int k = threadIdx.x * 8; k < K; k += 256 * 8) {
    const float4 w0 = *reinterpret_cast<const float4*>(wu + k), w1 = *reinterpret_cast<const float4*>(wu + k + 4);
    const uint4 u = *reinterpret_cast<const uint4*>(x + (int64_t)b * K + k);
    au += bf16_lo(u.x) * w0.x + bf16_hi(u.x) * w0.y + bf16_lo(u.y) * w0.z + bf16_hi(u.y) * w0.w +
          bf16_lo(u.z) * w1.x + bf16_hi(u.z) * w1.y + bf16_lo(u.w) * w1.z + bf16_hi(u.w) * w1.w;
    if (h != nullptr) {
      const float4 m0 = *reinterpret_cast<const float4*>(wm + k), m1 = *reinterpret_cast<const float4*>(wm + k + 4);
      const uint4 v = *reinterpret_cast<const uint4*>(h + (int64_t)b * K + k);
      am += bf16_lo(v.x) * m0.x + bf16_hi(v.x) * m0.y + bf16_lo(v.y) * m0.z + bf16_hi(v.y) * m0.w +
            bf16_lo(v.z) * m1.x + bf16_hi(v.z) * m1.y + bf16_lo(v.w) * m1.z + bf16_hi(v.w) * m1.w;
    }
  }
  au = block_sum_256(au, sh);
  if (h != nullptr) am = block_sum_256(am, sh);
  if (threadIdx.x == 0) {
    lu[b] = au + bu[0];
    if (h != nullptr) lm[b] = am + bm[0];
  }
}

// dx[b,k] = gu[b]*wu[k] (bf16), dwu[k] += sum_b gu[b]*x[b,k]; likewise (h, wm, gm).  A thread owns 8 consecutive feature
// indices (16-byte accesses) of the ROWS rows of its batch chunk (blockIdx.y): 4 x (GB / ROWS) blocks instead of the 16
// blocks x GB serial rows of the one-thread-per-index version (29 us at GB = 72 on the discriminator branch's critical
// path); the weight gradients of the chunks meet in red.global.add.v4.f32.  Block (0, 0) accumulates the bias gradients.
constexpr int DOTS_ROWS = 6;
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__global__ void __launch_bounds__(256) dhead_dots_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ h,
                                                             const float* __restrict__ wu, const float* __restrict__ wm,
                                                             const float* __restrict__ gu, const float* __restrict__ gm, int GB,
                                                             int K, bf16* __restrict__ dx, bf16* __restrict__ dh,
                                                             float* dwu, float* dbu, float* dwm, float* dbm) {
  const int k = (blockIdx.x * 256 + threadIdx.x) * 8;
  const int b0 = blockIdx.y * DOTS_ROWS;
  const int b1 = b0 + DOTS_ROWS < GB ? b0 + DOTS_ROWS : GB;
  if (k < K) {
    float w_u[8], w_m[8], a[8], c[8];
    *reinterpret_cast<float4*>(w_u) = *reinterpret_cast<const float4*>(wu + k);
    *reinterpret_cast<float4*>(w_u + 4) = *reinterpret_cast<const float4*>(wu + k + 4);
    if (h != nullptr) {
      *reinterpret_cast<float4*>(w_m) = *reinterpret_cast<const float4*>(wm + k);
      *reinterpret_cast<float4*>(w_m + 4) = *reinterpret_cast<const float4*>(wm + k + 4);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = c[i] = 0.f;
#pragma unroll 2
    for (int b = b0; b < b1; ++b) {
      const float g = gu[b];
      const uint4 u = *reinterpret_cast<const uint4*>(x + (int64_t)b * K + k);
      a[0] += g * bf16_lo(u.x); a[1] += g * bf16_hi(u.x); a[2] += g * bf16_lo(u.y); a[3] += g * bf16_hi(u.y);
      a[4] += g * bf16_lo(u.z); a[5] += g * bf16_hi(u.z); a[6] += g * bf16_lo(u.w); a[7] += g * bf16_hi(u.w);
      if (dx != nullptr)
        *reinterpret_cast<uint4*>(dx + (int64_t)b * K + k) =
            make_uint4(pack_bf16x2(g * w_u[0], g * w_u[1]), pack_bf16x2(g * w_u[2], g * w_u[3]),
                       pack_bf16x2(g * w_u[4], g * w_u[5]), pack_bf16x2(g * w_u[6], g * w_u[7]));
      if (h != nullptr) {
        const float q = gm[b];
        const uint4 v = *reinterpret_cast<const uint4*>(h + (int64_t)b * K + k);
        c[0] += q * bf16_lo(v.x); c[1] += q * bf16_hi(v.x); c[2] += q * bf16_lo(v.y); c[3] += q * bf16_hi(v.y);
        c[4] += q * bf16_lo(v.z); c[5] += q * bf16_hi(v.z); c[6] += q * bf16_lo(v.w); c[7] += q * bf16_hi(v.w);
        if (dh != nullptr)
          *reinterpret_cast<uint4*>(dh + (int64_t)b * K + k) =
              make_uint4(pack_bf16x2(q * w_m[0], q * w_m[1]), pack_bf16x2(q * w_m[2], q * w_m[3]),
                         pack_bf16x2(q * w_m[4], q * w_m[5]), pack_bf16x2(q * w_m[6], q * w_m[7]));
      }
    }
    if (dwu != nullptr) { red_add_v4(dwu + k, a[0], a[1], a[2], a[3]); red_add_v4(dwu + k + 4, a[4], a[5], a[6], a[7]); }
    if (h != nullptr && dwm != nullptr) { red_add_v4(dwm + k, c[0], c[1], c[2], c[3]); red_add_v4(dwm + k + 4, c[4], c[5], c[6], c[7]); }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32) {
    float su = 0.f, sm = 0.f;
    for (int b = threadIdx.x; b < GB; b += 32) { su += gu[b]; if (h != nullptr) sm += gm[b]; }
    su = warp_sum(su); sm = warp_sum(sm);
    if (threadIdx.x == 0) {
      if (dbu != nullptr) dbu[0] += su;
      if (h != nullptr && dbm != nullptr) dbm[0] += sm;
    }
  }
}

struct LossCfg {
  int groups, B, E1;
  int t_match[3], t_uncond[3], cls_tgt[3];   // 0/1 labels per group; class-target set per group (-1 none, 0, 1)
  float uncond_coeff;
};

// BCE against a constant label with torch's clamp: -max(log(p or 1-p), -100); p = sigmoid(z)
__device__ __forceinline__ float bce_term(float z, int t, float* p_out) {
  const float p = 1.f / (1.f + expf(-z));
  *p_out = p;
  const float lg = t ? logf(p) : log1pf(-p);
  return -fmaxf(lg, -100.f);
}
// d/dz of the above (through the sigmoid): autograd of -clamp(log p, -100): zero where the clamp is active
__device__ __forceinline__ float bce_grad(float p, int t) {
  if (t) return logf(p) >= -100.f ? -(1.f - p) : 0.f;
  return log1pf(-p) >= -100.f ? p : 0.f;
}

// grid = ceil(G*B / 8) blocks of 256 threads: one row per warp.  losses[4] must be ZERO on entry; every block adds its
// share: losses = {total, match, uncond_coeff*uncond, cls}; probs: pm/pu [G*B]; logp [G*B, E1] = log_softmax(cls)
__global__ void __launch_bounds__(256) dloss_fwd_kernel(LossCfg c, const float* __restrict__ lm, const float* __restrict__ lu,
                                                        const float* __restrict__ cls, const float* __restrict__ cp0,
                                                        const float* __restrict__ cp1, float* __restrict__ losses,
                                                        float* __restrict__ pm, float* __restrict__ pu, float* __restrict__ logp) {
  __shared__ float sh[8];
  const int GB = c.groups * c.B;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + w;
  float lmatch = 0.f, lunc = 0.f, lcls = 0.f;
  if (r < GB) {
    const int g = r / c.B;
    if (l == 0) {
      float p;
      lmatch = bce_term(lm[r], c.t_match[g], &p); pm[r] = p;
      lunc = bce_term(lu[r], c.t_uncond[g], &p); pu[r] = p;
    }
    if (cls != nullptr) {
      const float* row = cls + (int64_t)r * c.E1;
      float mx = -INFINITY;
      for (int e = l; e < c.E1; e += 32) mx = fmaxf(mx, row[e]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float se = 0.f;
      for (int e = l; e < c.E1; e += 32) se += expf(row[e] - mx);
      se = warp_sum(se);
      const float lse = mx + logf(se);
      const int tg = c.cls_tgt[g];
      const float* tgt = tg < 0 ? nullptr : (tg == 0 ? cp0 : cp1) + (int64_t)(r - g * c.B) * c.E1;
      for (int e = l; e < c.E1; e += 32) {
        const float lq = row[e] - lse;
        logp[(int64_t)r * c.E1 + e] = lq;
        if (tgt != nullptr) lcls -= tgt[e] * lq;
      }
    }
  }
  lmatch = block_sum_256(lmatch, sh);
  lunc = block_sum_256(lunc, sh);
  lcls = block_sum_256(lcls, sh);
  if (threadIdx.x == 0) {
    const float inv = 1.f / (float)c.B;      // every BCE term is a mean over its group of B; ce_loss divides by B
    const float m = lmatch * inv, u = c.uncond_coeff * lunc * inv, k = lcls * inv;
    atomicAdd(losses + 0, m + u + k); atomicAdd(losses + 1, m); atomicAdd(losses + 2, u); atomicAdd(losses + 3, k);
  }
}

// gradients of losses[0] w.r.t. the raw logits, scaled by the upstream scalar go[0]; same grid as the forward
__global__ void __launch_bounds__(256) dloss_bwd_kernel(LossCfg c, const float* __restrict__ go, const float* __restrict__ pm,
                                                        const float* __restrict__ pu, const float* __restrict__ logp,
                                                        const float* __restrict__ cp0, const float* __restrict__ cp1,
                                                        float* __restrict__ gm, float* __restrict__ gu, float* __restrict__ gcls) {
  const int GB = c.groups * c.B;
  const float s = go[0] / (float)c.B;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + w;
  if (r >= GB) return;
  const int g = r / c.B;
  if (l == 0) {
    gm[r] = s * bce_grad(pm[r], c.t_match[g]);
    gu[r] = s * c.uncond_coeff * bce_grad(pu[r], c.t_uncond[g]);
  }
  if (gcls == nullptr) return;
  const int tg = c.cls_tgt[g];
  const float* tgt = tg < 0 ? nullptr : (tg == 0 ? cp0 : cp1) + (int64_t)(r - g * c.B) * c.E1;
  float tsum = 0.f;
  if (tgt != nullptr) {
    for (int e = l; e < c.E1; e += 32) tsum += tgt[e];
    tsum = warp_sum(tsum);
  }
  for (int e = l; e < c.E1; e += 32) {
    // d/dlogit of -sum_e t_e * log_softmax_e = softmax_e * sum(t) - t_e
    const float v = tgt != nullptr ? (expf(logp[(int64_t)r * c.E1 + e]) * tsum - tgt[e]) * s : 0.f;
    gcls[(int64_t)r * c.E1 + e] = v;
  }
}

int make_cfg(LossCfg* c, int groups, int B, int E1, const int* t_match, const int* t_uncond, const int* cls_tgt, float uncond_coeff) {
  EKL_REQUIRE(groups >= 1 && groups <= 3 && B > 0 && E1 >= 0, "dloss: groups must be 1..3");
  c->groups = groups; c->B = B; c->E1 = E1; c->uncond_coeff = uncond_coeff;
  for (int g = 0; g < 3; ++g) {
    c->t_match[g] = g < groups ? t_match[g] : 0;
    c->t_uncond[g] = g < groups ? t_uncond[g] : 0;
    c->cls_tgt[g] = g < groups ? cls_tgt[g] : -1;
  }
  return 0;
}

}  // namespace

extern "C" int ekl_dhead_dots(const void* x_code, const void* h_c, const float* w_u, const float* b_u, const float* w_m,
                              const float* b_m, int GB, int K, float* logit_u, float* logit_m, void* stream) {
  EKL_REQUIRE(K % 8 == 0 && GB > 0, "dhead_dots: K %% 8");
  dhead_dots_kernel<<<GB, 256, 0, (cudaStream_t)stream>>>((const bf16*)x_code, (const bf16*)h_c, w_u, b_u, w_m, b_m, K, logit_u,
                                                         logit_m);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_dhead_dots_bwd(const void* x_code, const void* h_c, const float* w_u, const float* w_m, const float* g_u,
                                  const float* g_m, int GB, int K, void* dx_code, void* dh_c, float* dw_u, float* db_u,
                                  float* dw_m, float* db_m, void* stream) {
  EKL_REQUIRE(K % 8 == 0 && GB > 0, "dhead_dots_bwd: K %% 8");
  dhead_dots_bwd_kernel<<<dim3(ekl_cdiv(K / 8, 256), ekl_cdiv(GB, DOTS_ROWS)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x_code, (const bf16*)h_c, w_u, w_m, g_u, g_m, GB, K, (bf16*)dx_code, (bf16*)dh_c, dw_u, db_u, dw_m, db_m);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_dloss_fwd(int groups, int B, int E1, const int* t_match, const int* t_uncond, const int* cls_tgt,
                             float uncond_coeff, const float* logit_m, const float* logit_u, const float* cls_logits,
                             const float* cp0, const float* cp1, float* losses, float* p_m, float* p_u, float* logp,
                             void* stream) {
  LossCfg c;
  if (int rc = make_cfg(&c, groups, B, E1, t_match, t_uncond, cls_tgt, uncond_coeff)) return rc;
  EKL_CHECK_CUDA(cudaMemsetAsync(losses, 0, 4 * sizeof(float), (cudaStream_t)stream));
  dloss_fwd_kernel<<<ekl_cdiv(groups * B, 8), 256, 0, (cudaStream_t)stream>>>(c, logit_m, logit_u, cls_logits, cp0, cp1, losses, p_m, p_u, logp);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_dloss_bwd(int groups, int B, int E1, const int* t_match, const int* t_uncond, const int* cls_tgt,
                             float uncond_coeff, const float* go, const float* p_m, const float* p_u, const float* logp,
                             const float* cp0, const float* cp1, float* g_m, float* g_u, float* g_cls, void* stream) {
  LossCfg c;
  if (int rc = make_cfg(&c, groups, B, E1, t_match, t_uncond, cls_tgt, uncond_coeff)) return rc;
  dloss_bwd_kernel<<<ekl_cdiv(groups * B, 8), 256, 0, (cudaStream_t)stream>>>(c, go, p_m, p_u, logp, cp0, cp1, g_m, g_u, g_cls);
  EKL_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- conditioning augmentation: reparameterisation + KL
// CA_NET / VC_NET (model.py:145-152, 182-184, 198) + KL_loss (cub_trainer_splitz_cap_ca.py:54-58), one pass:
//   std = exp(0.5*logvar);  c = eps*std + mu;  kl = -0.5 * mean(1 + logvar - mu^2 - exp(logvar))
// mu / logvar are [B][D] with row strides (they are column halves of one GLU / Linear output).
namespace {

__global__ void __launch_bounds__(256) reparam_kl_fwd_kernel(const float* __restrict__ mu, int64_t mu_rs, const float* __restrict__ lv,
                                                             int64_t lv_rs, const float* __restrict__ eps, int B, int D,
                                                             float* __restrict__ c, float* __restrict__ stdv, float* __restrict__ kl) {
  __shared__ float sh[8];
  const int n = B * D;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int b = i / D, d = i - b * D;
    const float m = mu[b * mu_rs + d], l = lv[b * lv_rs + d];
    const float s = expf(0.5f * l);
    stdv[i] = s;
    c[i] = eps[i] * s + m;
    acc += 1.f + l - m * m - expf(l);
  }
  acc = block_sum_256(acc, sh);
  if (threadIdx.x == 0) kl[0] = -0.5f * acc / (float)n;
}

// dmu = dc + dkl*mu/n ; dlogvar = dc*eps*0.5*std + dstd*0.5*std - dkl*0.5*(1 - exp(logvar))/n   (dc / dstd / dkl may be null)
__global__ void __launch_bounds__(256) reparam_kl_bwd_kernel(const float* __restrict__ mu, int64_t mu_rs, const float* __restrict__ lv,
                                                             int64_t lv_rs, const float* __restrict__ eps, int B, int D,
                                                             const float* __restrict__ dc, const float* __restrict__ dstd,
                                                             const float* __restrict__ dkl, float* __restrict__ dmu,
                                                             float* __restrict__ dlv) {
  const int n = B * D;
  const float gk = dkl != nullptr ? dkl[0] / (float)n : 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int b = i / D, d = i - b * D;
    const float m = mu[b * mu_rs + d], l = lv[b * lv_rs + d];
    const float s = expf(0.5f * l);
    const float g = dc != nullptr ? dc[i] : 0.f;
    const float gs = dstd != nullptr ? dstd[i] : 0.f;
    dmu[i] = g + gk * m;
    dlv[i] = (g * eps[i] + gs) * 0.5f * s - gk * 0.5f * (1.f - expf(l));
  }
}

}  // namespace

extern "C" int ekl_reparam_kl_fwd(const float* mu, int64_t mu_row_stride, const float* logvar, int64_t lv_row_stride,
                                  const float* eps, int B, int D, float* c, float* stdv, float* kl, void* stream) {
  EKL_REQUIRE(B > 0 && D > 0, "reparam_kl: bad shape");
  reparam_kl_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mu, mu_row_stride, logvar, lv_row_stride, eps, B, D, c, stdv, kl);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_reparam_kl_bwd(const float* mu, int64_t mu_row_stride, const float* logvar, int64_t lv_row_stride,
                                  const float* eps, int B, int D, const float* dc, const float* dstd, const float* dkl,
                                  float* dmu, float* dlogvar, void* stream) {
  EKL_REQUIRE(B > 0 && D > 0, "reparam_kl: bad shape");
  reparam_kl_bwd_kernel<<<ekl_cdiv(B * D, 256), 256, 0, (cudaStream_t)stream>>>(mu, mu_row_stride, logvar, lv_row_stride, eps, B, D,
                                                                                dc, dstd, dkl, dmu, dlogvar);
  EKL_LAUNCH_CHECK();
  return 0;
}
