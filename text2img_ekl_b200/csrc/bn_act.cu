// Train-mode BatchNorm + activation family (HBM-bound, vectorised 16-byte accesses, fp32 statistics):
//   statistics  : per-(group, channel) sum / sum-of-squares partials  (also produced by the conv epilogue)
//   finalize    : partials -> mean, rstd (+ running_mean/var update, momentum 0.1, unbiased variance)
//   forward     : z = gamma*(y-mean)*rstd + beta ; out = GLU(z) | LeakyReLU(z, 0.2) | z (+ residual)
//   backward    : dz = act'(z)*dout ; partial sums S1 = sum dz, S2 = sum dz*xhat ; dy = gamma*rstd*(dz - S1/n - xhat*S2/n)
// Activations are bf16 [M rows][C channels] (NHWC flattened); rows are split into `groups` equal contiguous
// groups with independent batch statistics (real / wrong / fake discriminator passes batched into one tensor;
// reference: three separate netD(...) calls, cub_trainer_splitz_cap_ca.py:418-420).
#include "ekl_common.cuh"

namespace {

enum { ACT_NONE = 0, ACT_GLU = 1, ACT_LRELU = 2, ACT_RELU = 3 };

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// Common thread mapping: block = 256 threads = RT row-threads x CT octet-threads; blockIdx.x = octet strip,
// blockIdx.y = row chunk (chunks never straddle a statistics group).
struct Tile {
  int oct, rt, RT, CT;    // octet index (8 channels), row-thread id, row-threads / octet-threads per block
  int64_t r0, r1;         // row range of this block
  int g;                  // statistics group
  bool active;
};
__device__ __forceinline__ Tile make_tile(int64_t M, int noct, int groups) {
  Tile t;
  const int CT = noct < 256 ? noct : 256;
  t.RT = 256 / CT;
  t.CT = CT;
  const int ct = threadIdx.x % CT;
  t.rt = threadIdx.x / CT;
  t.oct = blockIdx.x * CT + ct;
  const int chunks_per_group = gridDim.y / groups;
  t.g = blockIdx.y / chunks_per_group;
  const int k = blockIdx.y - t.g * chunks_per_group;
  const int64_t Mg = M / groups;
  const int64_t per = (Mg + chunks_per_group - 1) / chunks_per_group;
  t.r0 = t.g * Mg + k * per;
  t.r1 = t.r0 + per < (t.g + 1) * Mg ? t.r0 + per : (t.g + 1) * Mg;
  t.active = t.oct < noct && t.rt < t.RT;
  return t;
}

// ---------------------------------------------------------------- statistics of a stored tensor
__global__ void __launch_bounds__(256) col_stats_kernel(const bf16* __restrict__ y, int64_t M, int C, int groups,
                                                        float* __restrict__ partial /*[gridDim.y][2][C]*/) {
  __shared__ float red[256 * 16];
  const Tile t = make_tile(M, C / 8, groups);
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (t.active)
    for (int64_t r = t.r0 + t.rt; r < t.r1; r += t.RT) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(y + r * C + t.oct * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] += f[i]; s2[i] += f[i] * f[i]; }
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[i]; red[threadIdx.x * 16 + 8 + i] = s2[i]; }
  __syncthreads();
  if (t.active && t.rt == 0) {
    const int CT = t.CT;
    for (int k = 1; k < t.RT; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] += red[(threadIdx.x + k * CT) * 16 + i]; s2[i] += red[(threadIdx.x + k * CT) * 16 + 8 + i]; }
    float* dst = partial + (size_t)blockIdx.y * 2 * C + t.oct * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[i] = s1[i]; dst[C + i] = s2[i]; }
  }
}

// ---------------------------------------------------------------- finalize
// partial [rows][2][C]; rows = groups * rows_per_group (group-contiguous). One block (256 thr) per 32 channels.
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int rows_per_group, int C,
                                                           int groups, float count, float eps, float momentum,
                                                           float* __restrict__ mean, float* __restrict__ rstd,
                                                           float* running_mean, float* running_var) {
  __shared__ double sh[2][32][33];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  for (int g = 0; g < groups; ++g) {
    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f, a3 = 0.f, b3 = 0.f;
    if (c < C) {
      const float* base = partial + (size_t)g * rows_per_group * 2 * C + c;
      int r = sl;
      for (; r + 96 < rows_per_group; r += 128) {       // 4 independent row loads in flight
        const float* p0 = base + (size_t)r * 2 * C;
        const float* p1 = p0 + (size_t)64 * C;
        const float* p2 = p1 + (size_t)64 * C;
        const float* p3 = p2 + (size_t)64 * C;
        a0 += p0[0]; b0 += p0[C]; a1 += p1[0]; b1 += p1[C]; a2 += p2[0]; b2 += p2[C]; a3 += p3[0]; b3 += p3[C];
      }
      for (; r < rows_per_group; r += 32) { const float* p0 = base + (size_t)r * 2 * C; a0 += p0[0]; b0 += p0[C]; }
    }
    sh[0][sl][cl] = (double)a0 + (double)a1 + (double)a2 + (double)a3;
    sh[1][sl][cl] = (double)b0 + (double)b1 + (double)b2 + (double)b3;
    __syncthreads();
    if (sl == 0 && c < C) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < 32; ++k) { a += sh[0][k][cl]; b += sh[1][k][cl]; }
      const double m = a / count;
      double var = b / count - m * m;
      if (var < 0.0) var = 0.0;
      mean[g * C + c] = (float)m;
      rstd[g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean != nullptr) {   // sequential updates, one per group == one per reference forward call
        const double unb = count > 1.f ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- forward
template <int ACT>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const bf16* __restrict__ y, int64_t M, int Cy, int groups,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const bf16* __restrict__ residual, bf16* __restrict__ out) {
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const Tile t = make_tile(M, Co / 8, groups);
  if (!t.active) return;
  float sc[8], sh[8], sc2[8], sh2[8];
  const int c0 = t.oct * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float s = gamma[c0 + i] * rstd[t.g * Cy + c0 + i];
    sc[i] = s; sh[i] = beta[c0 + i] - mean[t.g * Cy + c0 + i] * s;
    if (ACT == ACT_GLU) {
      const int c = Co + c0 + i;
      const float s_ = gamma[c] * rstd[t.g * Cy + c];
      sc2[i] = s_; sh2[i] = beta[c] - mean[t.g * Cy + c] * s_;
    }
  }
  constexpr int U = 4;                      // rows in flight per thread (memory-level parallelism)
  for (int64_t rb = t.r0 + t.rt; rb < t.r1; rb += (int64_t)U * t.RT) {
    uint4 ua[U], ub[U], uq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r < t.r1) {
        ua[u] = *reinterpret_cast<const uint4*>(y + r * Cy + c0);
        if (ACT == ACT_GLU) ub[u] = *reinterpret_cast<const uint4*>(y + r * Cy + Co + c0);
        if (residual != nullptr) uq[u] = *reinterpret_cast<const uint4*>(residual + r * Co + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r >= t.r1) break;
      float a[8], o[8];
      unpack8(ua[u], a);
      if (ACT == ACT_GLU) {
        float b[8];
        unpack8(ub[u], b);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (a[i] * sc[i] + sh[i]) * sigmoidf_(b[i] * sc2[i] + sh2[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = a[i] * sc[i] + sh[i];
          o[i] = ACT == ACT_LRELU ? (z > 0.f ? z : 0.2f * z) : (ACT == ACT_RELU ? fmaxf(z, 0.f) : z);
        }
      }
      if (residual != nullptr) {
        float q[8];
        unpack8(uq[u], q);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += q[i];
      }
      *reinterpret_cast<uint4*>(out + r * Co + c0) = pack8(o);
    }
  }
}

// dz for the 8 (GLU: 8+8) pre-activation channels a thread owns; also returns xhat
template <int ACT>
__device__ __forceinline__ void act_bwd8(const float* ya, const float* yb, const float* d, const float* sc, const float* sh,
                                         const float* sc2, const float* sh2, float* dza, float* dzb) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float z = ya[i] * sc[i] + sh[i];
    if (ACT == ACT_GLU) {
      const float s = sigmoidf_(yb[i] * sc2[i] + sh2[i]);
      dza[i] = d[i] * s;
      dzb[i] = d[i] * z * s * (1.f - s);
    } else if (ACT == ACT_LRELU) {
      dza[i] = z > 0.f ? d[i] : 0.2f * d[i];
    } else if (ACT == ACT_RELU) {
      dza[i] = z > 0.f ? d[i] : 0.f;
    } else {
      dza[i] = d[i];
    }
  }
}

// ---------------------------------------------------------------- backward, pass 1: partial sums
template <int ACT>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout,
                                                                int64_t M, int Cy, int groups,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ partial /*[gridDim.y][2][Cy]*/) {
  __shared__ float red[256 * 16];
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const Tile t = make_tile(M, Co / 8, groups);
  const int c0 = t.oct * 8;
  constexpr int NH = ACT == ACT_GLU ? 2 : 1;
  float sc[8], sh[8], sc2[8], sh2[8], mu[8], rs[8], mu2[8], rs2[8];
  float s1[NH][8], s2[NH][8];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[h][i] = s2[h][i] = 0.f;
  if (t.active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mu[i] = mean[t.g * Cy + c0 + i]; rs[i] = rstd[t.g * Cy + c0 + i];
      sc[i] = gamma[c0 + i] * rs[i]; sh[i] = beta[c0 + i] - mu[i] * sc[i];
      if (ACT == ACT_GLU) {
        const int c = Co + c0 + i;
        mu2[i] = mean[t.g * Cy + c]; rs2[i] = rstd[t.g * Cy + c];
        sc2[i] = gamma[c] * rs2[i]; sh2[i] = beta[c] - mu2[i] * sc2[i];
      }
    }
    constexpr int U = 4;
    for (int64_t rb = t.r0 + t.rt; rb < t.r1; rb += (int64_t)U * t.RT) {
      uint4 ua[U], ub[U], ud[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = rb + (int64_t)u * t.RT;
        if (r < t.r1) {
          ua[u] = *reinterpret_cast<const uint4*>(y + r * Cy + c0);
          if (ACT == ACT_GLU) ub[u] = *reinterpret_cast<const uint4*>(y + r * Cy + Co + c0);
          ud[u] = *reinterpret_cast<const uint4*>(dout + r * Co + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = rb + (int64_t)u * t.RT;
        if (r >= t.r1) break;
        float ya[8], yb[8], d[8], dza[8], dzb[8];
        unpack8(ua[u], ya);
        if (ACT == ACT_GLU) unpack8(ub[u], yb);
        unpack8(ud[u], d);
        act_bwd8<ACT>(ya, yb, d, sc, sh, sc2, sh2, dza, dzb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1[0][i] += dza[i]; s2[0][i] += dza[i] * (ya[i] - mu[i]) * rs[i];
          if (ACT == ACT_GLU) { s1[NH - 1][i] += dzb[i]; s2[NH - 1][i] += dzb[i] * (yb[i] - mu2[i]) * rs2[i]; }
        }
      }
    }
  }
  const int CT = t.CT;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[h][i]; red[threadIdx.x * 16 + 8 + i] = s2[h][i]; }
    __syncthreads();
    if (t.active && t.rt == 0) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = s1[h][i]; b[i] = s2[h][i]; }
      for (int k = 1; k < t.RT; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += red[(threadIdx.x + k * CT) * 16 + i]; b[i] += red[(threadIdx.x + k * CT) * 16 + 8 + i]; }
      float* dst = partial + (size_t)blockIdx.y * 2 * Cy + h * Co + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { dst[i] = a[i]; dst[Cy + i] = b[i]; }
    }
  }
}

// partial [groups*rows_per_group][2][Cy] -> sums [groups][2][Cy]; dgamma[c] += sum_g S2, dbeta[c] += sum_g S1
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows_per_group, int Cy,
                                                              int groups, float* __restrict__ sums, float* dgamma, float* dbeta) {
  __shared__ float sh[2][8][32];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float tg = 0.f, tb = 0.f;
  for (int g = 0; g < groups; ++g) {
    float a = 0.f, b = 0.f;
    if (c < Cy)
      for (int r = sl; r < rows_per_group; r += 8) {
        const float* p = partial + ((size_t)(g * rows_per_group + r)) * 2 * Cy;
        a += p[c]; b += p[Cy + c];
      }
    sh[0][sl][cl] = a; sh[1][sl][cl] = b;
    __syncthreads();
    if (sl == 0 && c < Cy) {
      for (int k = 1; k < 8; ++k) { a += sh[0][k][cl]; b += sh[1][k][cl]; }
      sums[(g * 2 + 0) * Cy + c] = a;
      sums[(g * 2 + 1) * Cy + c] = b;
      tb += a; tg += b;
    }
    __syncthreads();
  }
  if (sl == 0 && c < Cy) {
    if (dgamma) dgamma[c] += tg;
    if (dbeta) dbeta[c] += tb;
  }
}

// ---------------------------------------------------------------- backward, pass 2: dy
template <int ACT>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout,
                                                               int64_t M, int Cy, int groups,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ sums /*[groups][2][Cy]*/,
                                                               bf16* __restrict__ dy) {
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const Tile t = make_tile(M, Co / 8, groups);
  if (!t.active) return;
  const int c0 = t.oct * 8;
  const float inv_n = 1.f / (float)(M / groups);
  float sc[8], sh[8], sc2[8], sh2[8], mu[8], rs[8], mu2[8], rs2[8], m1[8], m2[8], m1b[8], m2b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean[t.g * Cy + c0 + i]; rs[i] = rstd[t.g * Cy + c0 + i];
    sc[i] = gamma[c0 + i] * rs[i]; sh[i] = beta[c0 + i] - mu[i] * sc[i];
    m1[i] = sums[(t.g * 2 + 0) * Cy + c0 + i] * inv_n; m2[i] = sums[(t.g * 2 + 1) * Cy + c0 + i] * inv_n;
    if (ACT == ACT_GLU) {
      const int c = Co + c0 + i;
      mu2[i] = mean[t.g * Cy + c]; rs2[i] = rstd[t.g * Cy + c];
      sc2[i] = gamma[c] * rs2[i]; sh2[i] = beta[c] - mu2[i] * sc2[i];
      m1b[i] = sums[(t.g * 2 + 0) * Cy + c] * inv_n; m2b[i] = sums[(t.g * 2 + 1) * Cy + c] * inv_n;
    }
  }
  constexpr int U = 4;
  for (int64_t rb = t.r0 + t.rt; rb < t.r1; rb += (int64_t)U * t.RT) {
    uint4 ua[U], ub[U], ud[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r < t.r1) {
        ua[u] = *reinterpret_cast<const uint4*>(y + r * Cy + c0);
        if (ACT == ACT_GLU) ub[u] = *reinterpret_cast<const uint4*>(y + r * Cy + Co + c0);
        ud[u] = *reinterpret_cast<const uint4*>(dout + r * Co + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r >= t.r1) break;
      float ya[8], yb[8], d[8], dza[8], dzb[8], o[8];
      unpack8(ua[u], ya);
      if (ACT == ACT_GLU) unpack8(ub[u], yb);
      unpack8(ud[u], d);
      act_bwd8<ACT>(ya, yb, d, sc, sh, sc2, sh2, dza, dzb);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = sc[i] * (dza[i] - m1[i] - (ya[i] - mu[i]) * rs[i] * m2[i]);
      *reinterpret_cast<uint4*>(dy + r * Cy + c0) = pack8(o);
      if (ACT == ACT_GLU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sc2[i] * (dzb[i] - m1b[i] - (yb[i] - mu2[i]) * rs2[i] * m2b[i]);
        *reinterpret_cast<uint4*>(dy + r * Cy + Co + c0) = pack8(o);
      }
    }
  }
}

// ---------------------------------------------------------------- plain LeakyReLU backward / concat helpers
__global__ void lrelu_bwd_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, bf16* __restrict__ dx, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float o[8], d[8];
    unpack8(reinterpret_cast<const uint4*>(out)[i], o);
    unpack8(reinterpret_cast<const uint4*>(dout)[i], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = o[k] > 0.f ? d[k] : 0.2f * d[k];
    reinterpret_cast<uint4*>(dx)[i] = pack8(d);
  }
}

// out[b,h,w,:] = cat(code[b,:], x[b,h,w,:])   (model.py:411-414: c_code tiled spatially, then cat along channels)
__global__ void cat_code_kernel(const float* __restrict__ code, int Cc, const bf16* __restrict__ x, int Cx, int64_t HW,
                                int64_t rows, bf16* __restrict__ out) {
  const int Ct = Cc + Cx;
  const int noct = Ct / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * noct; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / noct;
    const int c0 = (int)(i % noct) * 8;
    uint4 v;
    if (c0 < Cc) {
      const float* src = code + (r / HW) * Cc + c0;
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = src[k];
      v = pack8(f);
    } else {
      v = *reinterpret_cast<const uint4*>(x + r * Cx + (c0 - Cc));
    }
    *reinterpret_cast<uint4*>(out + r * Ct + c0) = v;
  }
}

// backward of cat_code: dcode[b,c] = sum_{hw} dcat[b,hw,c] (fp32); dx = dcat[..., Cc:]
__global__ void __launch_bounds__(256) cat_code_bwd_kernel(const bf16* __restrict__ dcat, int Cc, int Cx, int64_t HW,
                                                           float* __restrict__ dcode, bf16* __restrict__ dx) {
  const int Ct = Cc + Cx;
  const int b = blockIdx.x;
  const bf16* src = dcat + (int64_t)b * HW * Ct;
  // dx copy
  const int nox = Cx / 8;
  for (int64_t i = threadIdx.x; i < HW * nox; i += 256) {
    const int64_t r = i / nox;
    const int c0 = (int)(i % nox) * 8;
    *reinterpret_cast<uint4*>(dx + ((int64_t)b * HW + r) * Cx + c0) = *reinterpret_cast<const uint4*>(src + r * Ct + Cc + c0);
  }
  // dcode reduction: thread c over rows (coalesced across c)
  for (int c = threadIdx.x; c < Cc; c += 256) {
    float acc = 0.f;
    for (int64_t r = 0; r < HW; ++r) acc += __bfloat162float(src[r * Ct + c]);
    dcode[(int64_t)b * Cc + c] += acc;
  }
}

int grid_rows(int64_t M, int noct, int groups, dim3* grid) {
  const int CT = noct < 256 ? noct : 256;
  const int RT = 256 / CT;
  const int xs = ekl_cdiv(noct, CT);
  const int64_t Mg = M / groups;
  // enough chunks to fill the machine ~4x, at least 4*RT rows per chunk
  // ~3 fat blocks per SM: per-block prologue (per-channel parameter loads) is amortised over >= 16 rows per thread
  int64_t chunks = (148 * 3 + xs * groups - 1) / (xs * groups);
  const int64_t maxc = (Mg + 16 * RT - 1) / (16 * RT);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  *grid = dim3(xs, (unsigned)(chunks * groups));
  return (int)chunks;
}

}  // namespace

#include "../../include/ekl_b200.h"

extern "C" int ekl_col_stats_rows(int64_t M, int C, int groups) {
  dim3 grid;
  return grid_rows(M, C / 8, groups, &grid) * groups;
}

extern "C" int ekl_col_stats(const void* y, int64_t M, int C, int groups, float* partial, void* stream) {
  EKL_REQUIRE(C % 8 == 0 && M % groups == 0, "col_stats: C %% 8 and M %% groups required (C=%d)", C);
  dim3 grid;
  grid_rows(M, C / 8, groups, &grid);
  col_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, M, C, groups, partial);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_bn_finalize(const float* partial, int rows_per_group, int C, int groups, float count, float eps,
                               float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                               void* stream) {
  bn_finalize_kernel<<<ekl_cdiv(C, 32), 1024, 0, (cudaStream_t)stream>>>(partial, rows_per_group, C, groups, count, eps,
                                                                        momentum, mean, rstd, running_mean, running_var);
  EKL_LAUNCH_CHECK();
  return 0;
}

#define EKL_ACT_SWITCH(act, CALL)                       \
  switch (act) {                                        \
    case ACT_NONE: { constexpr int A = ACT_NONE; CALL; } break;   \
    case ACT_GLU: { constexpr int A = ACT_GLU; CALL; } break;     \
    case ACT_LRELU: { constexpr int A = ACT_LRELU; CALL; } break; \
    case ACT_RELU: { constexpr int A = ACT_RELU; CALL; } break;   \
    default: return ekl_fail(-1, "bad act %d", act);    \
  }

extern "C" int ekl_bn_act_fwd(const void* y, int64_t M, int Cy, int groups, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, int act, const void* residual, void* out,
                              void* stream) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  EKL_REQUIRE(Co % 8 == 0 && M % groups == 0, "bn_act_fwd: bad shape Cy=%d", Cy);
  dim3 grid;
  grid_rows(M, Co / 8, groups, &grid);
  EKL_ACT_SWITCH(act, (bn_act_fwd_kernel<A><<<grid, 256, 0, (cudaStream_t)stream>>>(
                          (const bf16*)y, M, Cy, groups, mean, rstd, gamma, beta, (const bf16*)residual, (bf16*)out)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_bn_act_bwd_rows(int64_t M, int Cy, int groups, int act) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  dim3 grid;
  return grid_rows(M, Co / 8, groups, &grid) * groups;
}

// partial: [ekl_bn_act_bwd_rows][2][Cy] scratch; sums: [groups][2][Cy] scratch; dgamma/dbeta accumulated (+=).
extern "C" int ekl_bn_act_bwd(const void* y, const void* dout, int64_t M, int Cy, int groups, const float* mean,
                              const float* rstd, const float* gamma, const float* beta, int act, float* partial,
                              float* sums, float* dgamma, float* dbeta, void* dy, void* stream) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  EKL_REQUIRE(Co % 8 == 0 && M % groups == 0, "bn_act_bwd: bad shape Cy=%d", Cy);
  dim3 grid;
  const int chunks = grid_rows(M, Co / 8, groups, &grid);
  cudaStream_t st = (cudaStream_t)stream;
  EKL_ACT_SWITCH(act, (bn_act_bwd_reduce_kernel<A><<<grid, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M, Cy, groups,
                                                                         mean, rstd, gamma, beta, partial)));
  EKL_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<ekl_cdiv(Cy, 32), 256, 0, st>>>(partial, chunks, Cy, groups, sums, dgamma, dbeta);
  EKL_LAUNCH_CHECK();
  EKL_ACT_SWITCH(act, (bn_act_bwd_apply_kernel<A><<<grid, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M, Cy, groups,
                                                                        mean, rstd, gamma, beta, sums, (bf16*)dy)));
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_lrelu_bwd(const void* out, const void* dout, void* dx, int64_t n, void* stream) {
  EKL_REQUIRE(n % 8 == 0, "lrelu_bwd: n %% 8");
  int blocks = (int)((n / 8 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  lrelu_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)out, (const bf16*)dout, (bf16*)dx, n / 8);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_cat_code(const float* code, int Cc, const void* x, int Cx, int B, int HW, void* out, void* stream) {
  EKL_REQUIRE(Cc % 8 == 0 && Cx % 8 == 0, "cat_code: channels %% 8");
  const int64_t rows = (int64_t)B * HW;
  int64_t total = rows * ((Cc + Cx) / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cat_code_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(code, Cc, (const bf16*)x, Cx, HW, rows, (bf16*)out);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_cat_code_bwd(const void* dcat, int Cc, int Cx, int B, int HW, float* dcode, void* dx, void* stream) {
  EKL_REQUIRE(Cc % 8 == 0 && Cx % 8 == 0, "cat_code_bwd: channels %% 8");
  cat_code_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const bf16*)dcat, Cc, Cx, HW, dcode, (bf16*)dx);
  EKL_LAUNCH_CHECK();
  return 0;
}
