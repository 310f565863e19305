// Train-mode BatchNorm + activation family (HBM-bound, vectorised 16-byte accesses, fp32 statistics):
//   statistics  : per-(group, channel) sum / sum-of-squares, accumulated with fp64 red.global.add into ONE [groups][2][C]
//                 double buffer (zero on entry) by whoever produces the tensor: the conv epilogue, splitk_finish, or
//                 col_stats for a stored tensor.  No per-tile partial rows and no finalize launch: every consumer block
//                 derives mean / rstd of its own channels from the two sums.
//   forward     : z = gamma*(y-mean)*rstd + beta ; out = GLU(z) | LeakyReLU(z, 0.2) | z (+ residual); the blocks of the
//                 first row chunk also store mean / rstd for the backward pass and update the running statistics
//   backward    : dz = act'(z)*dout ; partial sums S1 = sum dz, S2 = sum dz*xhat ; dy = gamma*rstd*(dz - S1/n - xhat*S2/n)
// Activations are bf16 [M rows][C channels] (NHWC flattened); rows are split into `groups` equal contiguous
// groups with independent batch statistics (real / wrong / fake discriminator passes batched into one tensor;
// reference: three separate netD(...) calls, cub_trainer_splitz_cap_ca.py:418-420).
#include <stdlib.h>

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
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// Common thread mapping: block = 256 threads = RT row-threads x CT octet-threads; blockIdx.x = octet strip,
// blockIdx.y = row chunk (chunks never straddle a statistics group).
struct Tile {
  int oct, rt, RT, CT;    // octet index (8 channels), row-thread id, row-threads / octet-threads per block
  int64_t r0, r1;         // row range of this block
  int g;                  // statistics group
  int k;                  // row-chunk index inside the group (chunk 0 does the per-group bookkeeping)
  bool active;
};
// rev: row chunks are walked from the END of the tensor (blockIdx.y = 0 takes the last chunk).  CTAs are dispatched in
// blockIdx order, so a reversed pass starts on the rows its producer / the previous pass touched LAST -- the ones still
// resident in the 126 MB L2.
__device__ __forceinline__ Tile make_tile(int64_t M, int noct, int groups, bool rev = false) {
  Tile t;
  const int by = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int CT = noct < 256 ? noct : 256;
  t.RT = 256 / CT;
  t.CT = CT;
  const int ct = threadIdx.x % CT;
  t.rt = threadIdx.x / CT;
  t.oct = blockIdx.x * CT + ct;
  const int chunks_per_group = gridDim.y / groups;
  t.g = by / chunks_per_group;
  const int k = by - t.g * chunks_per_group;
  t.k = k;
  const int64_t Mg = M / groups;
  const int64_t per = (Mg + chunks_per_group - 1) / chunks_per_group;
  t.r0 = t.g * Mg + k * per;
  t.r1 = t.r0 + per < (t.g + 1) * Mg ? t.r0 + per : (t.g + 1) * Mg;
  t.active = t.oct < noct && t.rt < t.RT;
  return t;
}

// fp64 accumulation of block-level partial sums into the [groups][2][C] statistics buffer
__device__ __forceinline__ void stat_add(double* dst, float v) { atomicAdd(dst, (double)v); }

// ---------------------------------------------------------------- statistics of a stored tensor
__global__ void __launch_bounds__(256) col_stats_kernel(const bf16* __restrict__ y, int64_t M, int C, int groups,
                                                        double* __restrict__ sums /*[groups][2][C], accumulated*/) {
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
    double* dst = sums + (size_t)t.g * 2 * C + t.oct * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { stat_add(dst + i, s1[i]); stat_add(dst + C + i, s2[i]); }
  }
}

// ---------------------------------------------------------------- split-K finish (conv_tc.cu split plans)
// fp32 scratch [M][C] (sum of the split work items' partial tiles) -> bf16 y, BatchNorm statistics of the ROUNDED
// values (same contract as the conv epilogue), and the scratch is re-zeroed for its next use.
__global__ void __launch_bounds__(256) splitk_finish_kernel(float* __restrict__ scratch, int64_t M, int C, int groups,
                                                            bf16* __restrict__ y, double* __restrict__ sums) {
  __shared__ float red[256 * 16];
  const Tile t = make_tile(M, C / 8, groups);
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (t.active)
    for (int64_t r = t.r0 + t.rt; r < t.r1; r += t.RT) {
      float4* src = reinterpret_cast<float4*>(scratch + r * C + t.oct * 8);
      const float4 a = src[0], b = src[1];
      src[0] = make_float4(0.f, 0.f, 0.f, 0.f); src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint4 o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
      *reinterpret_cast<uint4*>(y + r * C + t.oct * 8) = o;
      float f[8];
      unpack8(o, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] += f[i]; s2[i] += f[i] * f[i]; }
    }
  if (sums == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[i]; red[threadIdx.x * 16 + 8 + i] = s2[i]; }
  __syncthreads();
  if (t.active && t.rt == 0) {
    const int CT = t.CT;
    for (int k = 1; k < t.RT; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] += red[(threadIdx.x + k * CT) * 16 + i]; s2[i] += red[(threadIdx.x + k * CT) * 16 + 8 + i]; }
    double* dst = sums + (size_t)t.g * 2 * C + t.oct * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { stat_add(dst + i, s1[i]); stat_add(dst + C + i, s2[i]); }
  }
}

// ---------------------------------------------------------------- statistics -> coefficients
// mean / rstd of channel c of group g from the two fp64 sums (count = rows per group); the subtraction E[y^2] - mean^2
// is done in fp64, the reciprocal square root in fp32 like torch's batch_norm
struct BnStat { float mean, rstd, var; };
__device__ __forceinline__ BnStat bn_stat(const double* __restrict__ sums, int g, int C, int c, double inv_n, float eps) {
  const double m = sums[((size_t)g * 2 + 0) * C + c] * inv_n;
  double v = sums[((size_t)g * 2 + 1) * C + c] * inv_n - m * m;
  v = v < 0.0 ? 0.0 : v;
  BnStat o;
  o.mean = (float)m; o.var = (float)v; o.rstd = 1.f / sqrtf((float)v + eps);
  return o;
}

// ---------------------------------------------------------------- vector helpers for the streaming passes
// A thread owns VEC consecutive channels (GLU: VEC channels of EACH half): VEC = 4 (8-byte accesses) for GLU, whose
// two halves double the per-channel coefficients held in registers, VEC = 8 (16-byte accesses) otherwise.  Register
// budget <= 85 (3 blocks of 256 threads per SM) keeps >= 70 KB of loads in flight per SM.
template <int VEC> struct VecIO;
template <> struct VecIO<8> {
  typedef uint4 T;
  static __device__ __forceinline__ void unpack(const T& u, float* f) { unpack8(u, f); }
  static __device__ __forceinline__ T pack(const float* f) { return pack8(f); }
};
template <> struct VecIO<4> {
  typedef uint2 T;
  static __device__ __forceinline__ void unpack(const T& u, float* f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  }
  static __device__ __forceinline__ T pack(const float* f) { return make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3])); }
};

template <int VEC>
__device__ __forceinline__ Tile make_tile_v(int64_t M, int nvec, int groups) { return make_tile(M, nvec, groups); }

template <int ACT> struct ActVec { static constexpr int V = ACT == ACT_GLU ? 4 : 8; };

// ---------------------------------------------------------------- forward
// sums != null (training): statistics from the fp64 sums; mean_io / rstd_io [groups][Cy] are WRITTEN by chunk 0 of each
// group (saved for the backward pass) and block row 0 applies one running-statistics momentum update per group, in
// group order (one per reference forward call).  sums == null (inference): mean_io / rstd_io are inputs.
template <int ACT, int U>
__global__ void __launch_bounds__(256, U == 2 ? 3 : 2) bn_act_fwd_kernel(const bf16* __restrict__ y, int64_t M, int Cy, int groups,
                                                            const double* __restrict__ sums, float eps, float momentum,
                                                            float* mean_io, float* rstd_io, float* running_mean,
                                                            float* running_var,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const bf16* __restrict__ residual, bf16* __restrict__ out, int rev) {
  constexpr int VEC = ActVec<ACT>::V;
  using IO = VecIO<VEC>;
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const Tile t = make_tile(M, Co / VEC, groups, rev != 0);
  if (!t.active) return;
  const int c0 = t.oct * VEC;
  // software pipeline: the loads of the next U rows are in flight while the current U rows are computed; the first
  // batch is issued before the per-channel coefficient loads so the block prologue overlaps the stream
  typename IO::T ca[U], cb[U], cq[U], na[U], nb[U], nq[U];
  auto load = [&](typename IO::T* a, typename IO::T* b, typename IO::T* q, int64_t rb) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r < t.r1) {
        a[u] = *reinterpret_cast<const typename IO::T*>(y + r * Cy + c0);
        if (ACT == ACT_GLU) b[u] = *reinterpret_cast<const typename IO::T*>(y + r * Cy + Co + c0);
        if (residual != nullptr) q[u] = *reinterpret_cast<const typename IO::T*>(residual + r * Co + c0);
      }
    }
  };
  int64_t rb = t.r0 + t.rt;
  load(ca, cb, cq, rb);
  float sc[VEC], sh[VEC], sc2[VEC], sh2[VEC];
  const double inv_n = 1.0 / (double)(M / groups);
  constexpr int NHF = ACT == ACT_GLU ? 2 : 1;
#pragma unroll
  for (int h = 0; h < NHF; ++h)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = h * Co + c0 + i;
      float m, r;
      if (sums != nullptr) {
        const BnStat st = bn_stat(sums, t.g, Cy, c, inv_n, eps);
        m = st.mean; r = st.rstd;
        if (t.k == 0 && t.rt == 0) { mean_io[t.g * Cy + c] = m; rstd_io[t.g * Cy + c] = r; }
      } else {
        m = mean_io[t.g * Cy + c]; r = rstd_io[t.g * Cy + c];
      }
      const float s = gamma[c] * r;
      if (h == 0) { sc[i] = s; sh[i] = beta[c] - m * s; } else { sc2[i] = s; sh2[i] = beta[c] - m * s; }
    }
  if (sums != nullptr && running_mean != nullptr && blockIdx.y == 0 && t.rt == 0) {
    const float n = (float)(M / groups);
#pragma unroll
    for (int h = 0; h < NHF; ++h)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int c = h * Co + c0 + i;
        float rm = running_mean[c], rv = running_var[c];
        for (int g = 0; g < groups; ++g) {
          const BnStat st = bn_stat(sums, g, Cy, c, inv_n, eps);
          const float unb = n > 1.f ? st.var * n / (n - 1.f) : st.var;
          rm = (1.f - momentum) * rm + momentum * st.mean;
          rv = (1.f - momentum) * rv + momentum * unb;
        }
        running_mean[c] = rm; running_var[c] = rv;
      }
  }
  for (; rb < t.r1; rb += (int64_t)U * t.RT) {
    load(na, nb, nq, rb + (int64_t)U * t.RT);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r >= t.r1) break;
      float a[VEC], o[VEC];
      IO::unpack(ca[u], a);
      if (ACT == ACT_GLU) {
        float b[VEC];
        IO::unpack(cb[u], b);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = (a[i] * sc[i] + sh[i]) * sigmoidf_(b[i] * sc2[i] + sh2[i]);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float z = a[i] * sc[i] + sh[i];
          o[i] = ACT == ACT_LRELU ? (z > 0.f ? z : 0.2f * z) : (ACT == ACT_RELU ? fmaxf(z, 0.f) : z);
        }
      }
      if (residual != nullptr) {
        float q[VEC];
        IO::unpack(cq[u], q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] += q[i];
      }
      *reinterpret_cast<typename IO::T*>(out + r * Co + c0) = IO::pack(o);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { ca[u] = na[u]; cb[u] = nb[u]; cq[u] = nq[u]; }
  }
}

// dz (pre-activation gradient) for the VEC (GLU: VEC+VEC) channels a thread owns
template <int ACT, int VEC>
__device__ __forceinline__ void act_bwd(const float* ya, const float* yb, const float* d, const float* sc, const float* sh,
                                        const float* sc2, const float* sh2, float* dza, float* dzb) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
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
// S1 = sum dz, S2 = sum dz * xhat with xhat = y*rstd - mean*rstd (one FMA; coefficients rs / nmr)
template <int ACT, int U>
__global__ void __launch_bounds__(256, U == 2 ? 3 : 2) bn_act_bwd_reduce_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout,
                                                                   int64_t M, int Cy, int groups,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   double* __restrict__ sums /*[groups][2][Cy], zero on entry*/, int rev) {
  constexpr int VEC = ActVec<ACT>::V;
  using IO = VecIO<VEC>;
  __shared__ float red[256 * 2 * VEC];
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const Tile t = make_tile(M, Co / VEC, groups, rev != 0);
  const int c0 = t.oct * VEC;
  constexpr int NH = ACT == ACT_GLU ? 2 : 1;
  float sc[VEC], sh[VEC], sc2[VEC], sh2[VEC], rs[VEC], nmr[VEC], rs2[VEC], nmr2[VEC];
  float s1[NH][VEC], s2[NH][VEC];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int i = 0; i < VEC; ++i) s1[h][i] = s2[h][i] = 0.f;
  if (t.active) {
    typename IO::T ca[U], cb[U], cd[U], na[U], nb[U], nd[U];      // software pipeline, see bn_act_fwd_kernel
    auto load = [&](typename IO::T* a, typename IO::T* b, typename IO::T* d, int64_t rb) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = rb + (int64_t)u * t.RT;
        if (r < t.r1) {
          a[u] = *reinterpret_cast<const typename IO::T*>(y + r * Cy + c0);
          if (ACT == ACT_GLU) b[u] = *reinterpret_cast<const typename IO::T*>(y + r * Cy + Co + c0);
          d[u] = *reinterpret_cast<const typename IO::T*>(dout + r * Co + c0);
        }
      }
    };
    int64_t rb = t.r0 + t.rt;
    load(ca, cb, cd, rb);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float mu = mean[t.g * Cy + c0 + i];
      rs[i] = rstd[t.g * Cy + c0 + i]; nmr[i] = -mu * rs[i];
      sc[i] = gamma[c0 + i] * rs[i]; sh[i] = beta[c0 + i] - mu * sc[i];
      if (ACT == ACT_GLU) {
        const int c = Co + c0 + i;
        const float mu_ = mean[t.g * Cy + c];
        rs2[i] = rstd[t.g * Cy + c]; nmr2[i] = -mu_ * rs2[i];
        sc2[i] = gamma[c] * rs2[i]; sh2[i] = beta[c] - mu_ * sc2[i];
      }
    }
    for (; rb < t.r1; rb += (int64_t)U * t.RT) {
      load(na, nb, nd, rb + (int64_t)U * t.RT);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = rb + (int64_t)u * t.RT;
        if (r >= t.r1) break;
        float ya[VEC], yb[VEC], d[VEC], dza[VEC], dzb[VEC];
        IO::unpack(ca[u], ya);
        if (ACT == ACT_GLU) IO::unpack(cb[u], yb);
        IO::unpack(cd[u], d);
        act_bwd<ACT, VEC>(ya, yb, d, sc, sh, sc2, sh2, dza, dzb);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          s1[0][i] += dza[i]; s2[0][i] += dza[i] * (ya[i] * rs[i] + nmr[i]);
          if (ACT == ACT_GLU) { s1[NH - 1][i] += dzb[i]; s2[NH - 1][i] += dzb[i] * (yb[i] * rs2[i] + nmr2[i]); }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { ca[u] = na[u]; cb[u] = nb[u]; cd[u] = nd[u]; }
    }
  }
  const int CT = t.CT;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VEC; ++i) { red[threadIdx.x * 2 * VEC + i] = s1[h][i]; red[threadIdx.x * 2 * VEC + VEC + i] = s2[h][i]; }
    __syncthreads();
    if (t.active && t.rt == 0) {
      float a[VEC], b[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) { a[i] = s1[h][i]; b[i] = s2[h][i]; }
      for (int k = 1; k < t.RT; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          a[i] += red[(threadIdx.x + k * CT) * 2 * VEC + i];
          b[i] += red[(threadIdx.x + k * CT) * 2 * VEC + VEC + i];
        }
      double* dst = sums + (size_t)t.g * 2 * Cy + h * Co + c0;
#pragma unroll
      for (int i = 0; i < VEC; ++i) { stat_add(dst + i, a[i]); stat_add(dst + Cy + i, b[i]); }
    }
  }
}

// ---------------------------------------------------------------- backward, pass 2: dy
// dy = k*(dz - S1/n - xhat*S2/n), k = gamma*rstd  ==  k*dz + A*y + Bc  with  A = -k*rstd*S2/n,  Bc = -k*S1/n - A*mean
template <int ACT, int U>
__global__ void __launch_bounds__(256, U == 2 ? 3 : 2) bn_act_bwd_apply_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout,
                                                                  int64_t M, int Cy, int groups,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const double* __restrict__ sums /*[groups][2][Cy]*/,
                                                                  float* dgamma, float* dbeta, bf16* __restrict__ dy, int rev) {
  constexpr int VEC = ActVec<ACT>::V;
  using IO = VecIO<VEC>;
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const Tile t = make_tile(M, Co / VEC, groups, rev != 0);
  if (!t.active) return;
  const int c0 = t.oct * VEC;
  const float inv_n = 1.f / (float)(M / groups);
  typename IO::T ca[U], cb[U], cd[U], na[U], nb[U], nd[U];        // software pipeline, see bn_act_fwd_kernel
  auto load = [&](typename IO::T* a, typename IO::T* b, typename IO::T* d, int64_t rb) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r < t.r1) {
        a[u] = *reinterpret_cast<const typename IO::T*>(y + r * Cy + c0);
        if (ACT == ACT_GLU) b[u] = *reinterpret_cast<const typename IO::T*>(y + r * Cy + Co + c0);
        d[u] = *reinterpret_cast<const typename IO::T*>(dout + r * Co + c0);
      }
    }
  };
  int64_t rb = t.r0 + t.rt;
  load(ca, cb, cd, rb);
  float sc[VEC], sh[VEC], sc2[VEC], sh2[VEC], A[VEC], Bc[VEC], A2[VEC], Bc2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    {
      const int c = c0 + i;
      const float mu = mean[t.g * Cy + c], r = rstd[t.g * Cy + c];
      sc[i] = gamma[c] * r; sh[i] = beta[c] - mu * sc[i];
      A[i] = -sc[i] * r * (float)sums[(t.g * 2 + 1) * Cy + c] * inv_n;
      Bc[i] = -sc[i] * (float)sums[(t.g * 2 + 0) * Cy + c] * inv_n - A[i] * mu;
    }
    if (ACT == ACT_GLU) {
      const int c = Co + c0 + i;
      const float mu = mean[t.g * Cy + c], r = rstd[t.g * Cy + c];
      sc2[i] = gamma[c] * r; sh2[i] = beta[c] - mu * sc2[i];
      A2[i] = -sc2[i] * r * (float)sums[(t.g * 2 + 1) * Cy + c] * inv_n;
      Bc2[i] = -sc2[i] * (float)sums[(t.g * 2 + 0) * Cy + c] * inv_n - A2[i] * mu;
    }
  }
  // parameter gradients (accumulated): dgamma[c] += sum_g S2, dbeta[c] += sum_g S1 -- one thread per channel
  if (blockIdx.y == 0 && t.rt == 0 && (dgamma != nullptr || dbeta != nullptr)) {
#pragma unroll
    for (int h = 0; h < (ACT == ACT_GLU ? 2 : 1); ++h)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int c = h * Co + c0 + i;
        double tb = 0.0, tg = 0.0;
        for (int g = 0; g < groups; ++g) { tb += sums[(g * 2 + 0) * Cy + c]; tg += sums[(g * 2 + 1) * Cy + c]; }
        if (dgamma != nullptr) dgamma[c] += (float)tg;
        if (dbeta != nullptr) dbeta[c] += (float)tb;
      }
  }
  for (; rb < t.r1; rb += (int64_t)U * t.RT) {
    load(na, nb, nd, rb + (int64_t)U * t.RT);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * t.RT;
      if (r >= t.r1) break;
      float ya[VEC], yb[VEC], d[VEC], dza[VEC], dzb[VEC], o[VEC];
      IO::unpack(ca[u], ya);
      if (ACT == ACT_GLU) IO::unpack(cb[u], yb);
      IO::unpack(cd[u], d);
      act_bwd<ACT, VEC>(ya, yb, d, sc, sh, sc2, sh2, dza, dzb);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = sc[i] * dza[i] + (A[i] * ya[i] + Bc[i]);
      *reinterpret_cast<typename IO::T*>(dy + r * Cy + c0) = IO::pack(o);
      if (ACT == ACT_GLU) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = sc2[i] * dzb[i] + (A2[i] * yb[i] + Bc2[i]);
        *reinterpret_cast<typename IO::T*>(dy + r * Cy + Co + c0) = IO::pack(o);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { ca[u] = na[u]; cb[u] = nb[u]; cd[u] = nd[u]; }
  }
}

// ---------------------------------------------------------------- small layers: backward in one launch
// Layers with few rows per statistics group (the 4x4 / 8x8 discriminator tails, the generator stem's BatchNorm1d) are
// launch-latency bound.  Here a block owns a strip of 8 channel vectors of one group for ALL rows, so the backward sums
// never leave the block: rows -> S1, S2 (warp shuffles + shared memory) -> dy (second sweep hits L1/L2); dgamma / dbeta
// by atomics.  Block = 256 threads = 8 vector-threads x 32 row-threads.  (The forward pass needs no such variant: the
// statistics arrive as two sums per channel, so the streaming kernel is a single launch for every size.)
constexpr int SM_CT = 8, SM_RT = 32;

// reduce NV per-thread values over the 32 row-threads that share a vector-thread; result valid in every thread
template <int NV>
__device__ __forceinline__ void strip_reduce(float* v, float* sh /*[8 warps][SM_CT][NV]*/, float* tot /*[SM_CT][NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ct = threadIdx.x % SM_CT;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 8);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
  __syncthreads();
  if (lane < SM_CT) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[(warp * SM_CT + ct) * NV + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < SM_CT * NV) {
    const int c = threadIdx.x / NV, i = threadIdx.x % NV;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += sh[(w * SM_CT + c) * NV + i];
    tot[c * NV + i] = a;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = tot[ct * NV + i];
}

template <int ACT>
__global__ void __launch_bounds__(256) bn_act_bwd_small_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout, int64_t Mg,
                                                               int Cy, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* dgamma, float* dbeta, bf16* __restrict__ dy) {
  constexpr int VEC = ActVec<ACT>::V;
  constexpr int NH = ACT == ACT_GLU ? 2 : 1;
  constexpr int NV = 2 * NH * VEC;
  using IO = VecIO<VEC>;
  __shared__ float sh[8 * SM_CT * NV];
  __shared__ float tot[SM_CT * NV];
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy;
  const int ct = threadIdx.x % SM_CT, rt = threadIdx.x / SM_CT;
  const int g = blockIdx.y;
  const int c0 = (blockIdx.x * SM_CT + ct) * VEC;
  const bf16* yg = y + (int64_t)g * Mg * Cy;
  const bf16* dg = dout + (int64_t)g * Mg * Co;
  bf16* dyg = dy + (int64_t)g * Mg * Cy;
  float sc[VEC], shf[VEC], sc2[VEC], sh2[VEC], rs[VEC], nmr[VEC], rs2[VEC], nmr2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float mu = mean[g * Cy + c0 + i];
    rs[i] = rstd[g * Cy + c0 + i]; nmr[i] = -mu * rs[i];
    sc[i] = gamma[c0 + i] * rs[i]; shf[i] = beta[c0 + i] - mu * sc[i];
    if (ACT == ACT_GLU) {
      const int c = Co + c0 + i;
      const float mu_ = mean[g * Cy + c];
      rs2[i] = rstd[g * Cy + c]; nmr2[i] = -mu_ * rs2[i];
      sc2[i] = gamma[c] * rs2[i]; sh2[i] = beta[c] - mu_ * sc2[i];
    }
  }
  float acc[NV];       // [S1 a | S2 a | S1 b | S2 b]
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  constexpr int U = 4;        // rows in flight per thread
  for (int64_t rb = rt; rb < Mg; rb += (int64_t)U * SM_RT) {
    typename IO::T ua[U], ub[U], ud[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * SM_RT;
      if (r < Mg) {
        ua[u] = *reinterpret_cast<const typename IO::T*>(yg + r * Cy + c0);
        if (ACT == ACT_GLU) ub[u] = *reinterpret_cast<const typename IO::T*>(yg + r * Cy + Co + c0);
        ud[u] = *reinterpret_cast<const typename IO::T*>(dg + r * Co + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (rb + (int64_t)u * SM_RT >= Mg) break;
      float ya[VEC], yb[VEC], d[VEC], dza[VEC], dzb[VEC];
      IO::unpack(ua[u], ya);
      if (ACT == ACT_GLU) IO::unpack(ub[u], yb);
      IO::unpack(ud[u], d);
      act_bwd<ACT, VEC>(ya, yb, d, sc, shf, sc2, sh2, dza, dzb);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        acc[i] += dza[i]; acc[VEC + i] += dza[i] * (ya[i] * rs[i] + nmr[i]);
        if (ACT == ACT_GLU) { acc[2 * VEC + i] += dzb[i]; acc[3 * VEC + i] += dzb[i] * (yb[i] * rs2[i] + nmr2[i]); }
      }
    }
  }
  strip_reduce<NV>(acc, sh, tot);
  if (rt == 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (dbeta != nullptr) atomicAdd(dbeta + c0 + i, acc[i]);
      if (dgamma != nullptr) atomicAdd(dgamma + c0 + i, acc[VEC + i]);
      if (ACT == ACT_GLU) {
        if (dbeta != nullptr) atomicAdd(dbeta + Co + c0 + i, acc[2 * VEC + i]);
        if (dgamma != nullptr) atomicAdd(dgamma + Co + c0 + i, acc[3 * VEC + i]);
      }
    }
  }
  const float inv_n = 1.f / (float)Mg;
  float A[VEC], Bc[VEC], A2[VEC], Bc2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    A[i] = -sc[i] * rs[i] * acc[VEC + i] * inv_n;
    Bc[i] = -sc[i] * acc[i] * inv_n + A[i] * (nmr[i] / rs[i]);          // - A*mean, mean = -nmr/rs
    if (ACT == ACT_GLU) {
      A2[i] = -sc2[i] * rs2[i] * acc[3 * VEC + i] * inv_n;
      Bc2[i] = -sc2[i] * acc[2 * VEC + i] * inv_n + A2[i] * (nmr2[i] / rs2[i]);
    }
  }
  for (int64_t rb = rt; rb < Mg; rb += (int64_t)U * SM_RT) {
    typename IO::T ua[U], ub[U], ud[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * SM_RT;
      if (r < Mg) {
        ua[u] = *reinterpret_cast<const typename IO::T*>(yg + r * Cy + c0);
        if (ACT == ACT_GLU) ub[u] = *reinterpret_cast<const typename IO::T*>(yg + r * Cy + Co + c0);
        ud[u] = *reinterpret_cast<const typename IO::T*>(dg + r * Co + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * SM_RT;
      if (r >= Mg) break;
      float ya[VEC], yb[VEC], d[VEC], dza[VEC], dzb[VEC], o[VEC];
      IO::unpack(ua[u], ya);
      if (ACT == ACT_GLU) IO::unpack(ub[u], yb);
      IO::unpack(ud[u], d);
      act_bwd<ACT, VEC>(ya, yb, d, sc, shf, sc2, sh2, dza, dzb);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = sc[i] * dza[i] + (A[i] * ya[i] + Bc[i]);
      *reinterpret_cast<typename IO::T*>(dyg + r * Cy + c0) = IO::pack(o);
      if (ACT == ACT_GLU) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = sc2[i] * dzb[i] + (A2[i] * yb[i] + Bc2[i]);
        *reinterpret_cast<typename IO::T*>(dyg + r * Cy + Co + c0) = IO::pack(o);
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

// backward of cat_code: dcode[b,c] += sum_{hw} dcat[b,hw,c] (fp32); dx = dcat[..., Cc:].
// grid = (row chunks, B); block = RT row-threads x (Ct/8) octet-threads: every thread streams 16-byte octets of its
// column strip (code octets are accumulated in registers, feature octets are copied), then the row-threads are
// combined through shared memory and one atomicAdd per (block, code channel) lands in dcode.
__global__ void __launch_bounds__(256) cat_code_bwd_kernel(const bf16* __restrict__ dcat, int Cc, int Cx, int64_t HW,
                                                           int rows_per_block, float* __restrict__ dcode,
                                                           bf16* __restrict__ dx) {
  __shared__ float red[256 * 8];
  const int Ct = Cc + Cx;
  const int noct = Ct / 8, ncode = Cc / 8;
  const int RT = blockDim.x / noct;
  const int oct = threadIdx.x % noct, rt = threadIdx.x / noct;
  const int b = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block) < HW ? (r0 + rows_per_block) : HW;
  const bf16* src = dcat + (int64_t)b * HW * Ct;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rt < RT) {
    constexpr int U = 4;
    for (int64_t rb = r0 + rt; rb < r1; rb += (int64_t)U * RT) {
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t r = rb + (int64_t)k * RT;
        if (r < r1) u[k] = *reinterpret_cast<const uint4*>(src + r * Ct + oct * 8);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t r = rb + (int64_t)k * RT;
        if (r >= r1) break;
        if (oct < ncode) {
          float f[8];
          unpack8(u[k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
        } else {
          *reinterpret_cast<uint4*>(dx + ((int64_t)b * HW + r) * Cx + (oct - ncode) * 8) = u[k];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  if (rt == 0 && oct < ncode) {
    for (int k = 1; k < RT; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += red[(threadIdx.x + k * noct) * 8 + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dcode + (int64_t)b * Cc + oct * 8 + i, acc[i]);
  }
}

int act_vec(int act) { return act == ACT_GLU ? 4 : 8; }

// small-layer single-launch path: few rows per statistics group, whole strips of 8 channel vectors
bool bn_small(int64_t M, int Co, int groups, int act) {
  return M / groups <= 768 && Co % (act_vec(act) * 8) == 0;
}

int grid_rows(int64_t M, int noct, int groups, dim3* grid) {
  const int CT = noct < 256 ? noct : 256;
  const int RT = 256 / CT;
  const int xs = ekl_cdiv(noct, CT);
  const int64_t Mg = M / groups;
  // ~3 fat blocks per SM: the per-block prologue (per-channel coefficients) is amortised over >= 16 rows per thread on
  // large layers; small layers (fewer than 148*3 such chunks) go down to 4 rows per thread to get more blocks in flight
  int64_t chunks = (148 * 3 + xs * groups - 1) / (xs * groups);
  int64_t maxc = (Mg + 16 * RT - 1) / (16 * RT);
  if (maxc < chunks) maxc = (Mg + 4 * RT - 1) / (4 * RT);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  *grid = dim3(xs, (unsigned)(chunks * groups));
  return (int)chunks;
}

}  // namespace

#include "../../include/ekl_b200.h"

// split-K outputs are small (<= a few MB): fine chunks (4 rows per thread) so the pass is not latency bound
static int finish_grid(int64_t M, int noct, int groups, dim3* grid) {
  const int CT = noct < 256 ? noct : 256;
  const int RT = 256 / CT;
  const int xs = ekl_cdiv(noct, CT);
  const int64_t Mg = M / groups;
  int64_t chunks = (Mg + 4 * RT - 1) / (4 * RT);
  if (chunks > 256) chunks = 256;
  if (chunks < 1) chunks = 1;
  *grid = dim3(xs, (unsigned)(chunks * groups));
  return (int)chunks;
}

int ekl_splitk_finish(float* scratch, int64_t M, int C, int groups, void* y, double* sums, cudaStream_t st) {
  EKL_REQUIRE(C % 8 == 0 && M % groups == 0, "splitk_finish: C %% 8 and M %% groups required (C=%d)", C);
  dim3 grid;
  finish_grid(M, C / 8, groups, &grid);
  splitk_finish_kernel<<<grid, 256, 0, st>>>(scratch, M, C, groups, (bf16*)y, sums);
  EKL_LAUNCH_CHECK();
  return 0;
}

extern "C" int ekl_col_stats(const void* y, int64_t M, int C, int groups, double* sums, void* stream) {
  EKL_REQUIRE(y != nullptr && sums != nullptr, "col_stats: null pointer argument");
  EKL_REQUIRE(C % 8 == 0 && M % groups == 0, "col_stats: C %% 8 and M %% groups required (C=%d)", C);
  dim3 grid;
  grid_rows(M, C / 8, groups, &grid);
  col_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, M, C, groups, sums);
  EKL_LAUNCH_CHECK();
  return 0;
}

// rows in flight per thread of the streaming kernels: 2 (3 blocks / SM) or 4 (2 blocks / SM); EKL_BN_U selects (experiments)
static int bn_u() {
  static int u = 0;
  if (u == 0) { const char* e = getenv("EKL_BN_U"); u = (e && e[0] == '4') ? 4 : 2; }
  return u;
}

// traversal direction of the streaming passes (bit 0 forward pass, bit 1 backward reduce, bit 2 backward apply run from the
// end of the tensor); EKL_BN_REV selects (experiments)
static int bn_rev() {
  static int r = -1;
  if (r < 0) { const char* e = getenv("EKL_BN_REV"); r = e ? atoi(e) : 0; }
  return r;
}

#define EKL_ACT_SWITCH(act, CALL)                       \
  switch (act) {                                        \
    case ACT_NONE: { constexpr int A = ACT_NONE; CALL; } break;   \
    case ACT_GLU: { constexpr int A = ACT_GLU; CALL; } break;     \
    case ACT_LRELU: { constexpr int A = ACT_LRELU; CALL; } break; \
    case ACT_RELU: { constexpr int A = ACT_RELU; CALL; } break;   \
    default: return ekl_fail(-1, "bad act %d", act);    \
  }

extern "C" int ekl_bn_act_fwd(const void* y, int64_t M, int Cy, int groups, const double* sums, float eps, float momentum,
                              float* mean, float* rstd, float* running_mean, float* running_var, const float* gamma,
                              const float* beta, int act, const void* residual, void* out, void* stream) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  EKL_REQUIRE(Co % 8 == 0 && M % groups == 0, "bn_act_fwd: bad shape Cy=%d", Cy);
  EKL_REQUIRE(y != nullptr && out != nullptr && mean != nullptr && rstd != nullptr, "bn_act_fwd: null pointer argument");
  dim3 grid;
  grid_rows(M, Co / act_vec(act), groups, &grid);
  if (bn_u() == 4) {
    EKL_ACT_SWITCH(act, (bn_act_fwd_kernel<A, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(
                            (const bf16*)y, M, Cy, groups, sums, eps, momentum, mean, rstd, running_mean, running_var, gamma, beta,
                            (const bf16*)residual, (bf16*)out, bn_rev() & 1)));
  } else {
    EKL_ACT_SWITCH(act, (bn_act_fwd_kernel<A, 2><<<grid, 256, 0, (cudaStream_t)stream>>>(
                            (const bf16*)y, M, Cy, groups, sums, eps, momentum, mean, rstd, running_mean, running_var, gamma, beta,
                            (const bf16*)residual, (bf16*)out, bn_rev() & 1)));
  }
  EKL_LAUNCH_CHECK();
  return 0;
}

// sums: [groups][2][Cy] doubles of caller scratch, ZERO on entry (unused by single-launch small layers); dgamma/dbeta
// accumulated (+=).
extern "C" int ekl_bn_act_bwd(const void* y, const void* dout, int64_t M, int Cy, int groups, const float* mean,
                              const float* rstd, const float* gamma, const float* beta, int act, double* sums,
                              float* dgamma, float* dbeta, void* dy, void* stream) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  EKL_REQUIRE(Co % 8 == 0 && M % groups == 0, "bn_act_bwd: bad shape Cy=%d", Cy);
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_small(M, Co, groups, act)) {
    dim3 sg(Co / act_vec(act) / SM_CT, groups);
    EKL_ACT_SWITCH(act, (bn_act_bwd_small_kernel<A><<<sg, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M / groups, Cy, mean, rstd,
                                                                      gamma, beta, dgamma, dbeta, (bf16*)dy)));
    EKL_LAUNCH_CHECK();
    return 0;
  }
  EKL_REQUIRE(sums != nullptr, "bn_act_bwd: sums scratch required");
  dim3 grid;
  grid_rows(M, Co / act_vec(act), groups, &grid);
  if (bn_u() == 4) {
    EKL_ACT_SWITCH(act, (bn_act_bwd_reduce_kernel<A, 4><<<grid, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M, Cy, groups,
                                                                              mean, rstd, gamma, beta, sums, bn_rev() & 2)));
    EKL_LAUNCH_CHECK();
    EKL_ACT_SWITCH(act, (bn_act_bwd_apply_kernel<A, 4><<<grid, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M, Cy, groups,
                                                                             mean, rstd, gamma, beta, sums, dgamma, dbeta, (bf16*)dy, bn_rev() & 4)));
  } else {
    EKL_ACT_SWITCH(act, (bn_act_bwd_reduce_kernel<A, 2><<<grid, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M, Cy, groups,
                                                                              mean, rstd, gamma, beta, sums, bn_rev() & 2)));
    EKL_LAUNCH_CHECK();
    EKL_ACT_SWITCH(act, (bn_act_bwd_apply_kernel<A, 2><<<grid, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M, Cy, groups,
                                                                             mean, rstd, gamma, beta, sums, dgamma, dbeta, (bf16*)dy, bn_rev() & 4)));
  }
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
  EKL_REQUIRE(Cc % 8 == 0 && Cx % 8 == 0 && (Cc + Cx) / 8 <= 256, "cat_code_bwd: channels %% 8, <= 2048 in total");
  const int noct = (Cc + Cx) / 8;
  const int RT = 256 / noct;
  // ~4 blocks per SM over the whole batch, at least 4*RT rows each
  int chunks = ekl_cdiv(148 * 4, B);
  int rows = ekl_cdiv(HW, chunks);
  if (rows < 4 * RT) rows = 4 * RT;
  chunks = ekl_cdiv(HW, rows);
  cat_code_bwd_kernel<<<dim3(chunks, B), RT * noct, 0, (cudaStream_t)stream>>>((const bf16*)dcat, Cc, Cx, HW, rows, dcode,
                                                                              (bf16*)dx);
  EKL_LAUNCH_CHECK();
  return 0;
}
