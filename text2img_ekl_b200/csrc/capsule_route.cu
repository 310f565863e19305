// Capsule routing-by-agreement for the discriminator class head (JOINT_D_NET64/128 fc_ac_cap, model.py:943,967-971,1082:
// CapsuleLinear(out_capsules = ENTITY_DIM + 1 = 201, in_length = 8*ndf = 512, out_length = 16) on the 16 positions of the
// 4x4 trunk output).
//
// *Parity unpinned*: the reference imports the un-vendored `capsule_layer` package; the arithmetic restated in
// oracle/capsule_ref.py (dynamic routing, 3 iterations: softmax over OUT capsules, squash) is what these kernels implement.
//
// Wide in_length, few in-capsules: here the prior tensor prior[b,i,o,:] = W[o] x[b,i] is SMALL (16 x 201 x 16 floats =
// 206 KB per sample) and comes from one plain library GEMM [B*16, 512] x [512, 201*16]; everything after it -- three
// routing iterations of softmax-over-o, weighted sum over i, squash, agreement update, and the final capsule norm -- is
// ONE kernel per direction (the eager form is ~25 launches forward and ~60 backward per discriminator call).
//
// One CTA of 512 threads per sample.  A half-warp owns an out-capsule o at a time (o = h, h+32, ...); its lane j owns
// in-capsule i = j and streams prior[b,j,o,0:16] (64 contiguous bytes) from L2 once per pass.  Per (o, i) scalars
// (agreement logits, coupling coefficients, their gradients) live in registers for the whole kernel; per-o vectors
// (s, v, their gradients) live in shared memory, component l held by lane l.  Sums over i are 16-lane shuffle
// reduce-scatters, softmax statistics over o are lane-local partials combined across the 32 half-warps through shared
// memory.  The backward recomputes the forward on chip: nothing but the prior is saved.
//   forward : 3 passes over the prior; backward: 5 passes + 1 write of its gradient.  Latency / L2 bound, ~1.2 MB per sample.
#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

namespace {

constexpr int RT = 512;              // threads per CTA
constexpr int NH = RT / 16;          // half-warps per CTA
constexpr int CL = 16;               // out_length == in_capsules == 16 (one lane per component / in-capsule)
constexpr float CAPS_EPS = 1e-8f;    // oracle/capsule_ref.py squash epsilon
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float hsum(float v) {        // sum over the 16 lanes of a half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// w[l] (l = 0..15) per lane -> lane j returns sum over the 16 lanes of w[j]   (15 shuffles)
__device__ __forceinline__ float reduce_scatter16(const float (&w)[16], int j) {
  float a8[8], a4[4], a2[2];
  const bool b3 = (j & 8) != 0, b2 = (j & 4) != 0, b1 = (j & 2) != 0, b0 = (j & 1) != 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float mine = b3 ? w[8 + k] : w[k], other = b3 ? w[k] : w[8 + k];
    a8[k] = mine + __shfl_xor_sync(FULL, other, 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float mine = b2 ? a8[4 + k] : a8[k], other = b2 ? a8[k] : a8[4 + k];
    a4[k] = mine + __shfl_xor_sync(FULL, other, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float mine = b1 ? a4[2 + k] : a4[k], other = b1 ? a4[k] : a4[2 + k];
    a2[k] = mine + __shfl_xor_sync(FULL, other, 2);
  }
  const float mine = b0 ? a2[1] : a2[0], other = b0 ? a2[0] : a2[1];
  return mine + __shfl_xor_sync(FULL, other, 1);
}

__device__ __forceinline__ void load_prior(const float* __restrict__ prior, int b, int j, int O, int o, bool live, float (&p)[16]) {
  if (live) {
    const float4* src = reinterpret_cast<const float4*>(prior + (((size_t)b * CL + j) * O + o) * CL);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = __ldg(src + q);
      p[4 * q] = v.x; p[4 * q + 1] = v.y; p[4 * q + 2] = v.z; p[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int l = 0; l < 16; ++l) p[l] = 0.f;
  }
}

__device__ __forceinline__ float dot16(const float (&p)[16], const float* __restrict__ row) {      // row: 16 floats in smem
  float a = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
    a += p[4 * q] * v.x + p[4 * q + 1] * v.y + p[4 * q + 2] * v.z + p[4 * q + 3] * v.w;
  }
  return a;
}

// squash of a lane-distributed vector (component j in lane j)
__device__ __forceinline__ float squash_lane(float s) {
  const float n2 = hsum(s * s);
  return s * (n2 / (1.f + n2) * rsqrtf(n2 + CAPS_EPS));
}

// backward of v = squash(s): gs = gv*f + s * (2 f'(n2) <gv, s>), lane-distributed
__device__ __forceinline__ float squash_bwd_lane(float s, float gv) {
  const float n2 = hsum(s * s), dot = hsum(gv * s);
  const float r = rsqrtf(n2 + CAPS_EPS), q = 1.f / (1.f + n2);
  const float f = n2 * q * r;
  const float fp = q * r - n2 * q * q * r - 0.5f * n2 * q * r * r * r;
  return gv * f + s * (2.f * fp * dot);
}

// column-wise (per in-capsule j) combine of one partial per thread across the NH half-warps
template <bool MAX>
__device__ __forceinline__ float block_col(float local, float* red, int h, int j) {
  red[h * 16 + j] = local;
  __syncthreads();
  float r = red[j];
#pragma unroll 8
  for (int q = 1; q < NH; ++q) r = MAX ? fmaxf(r, red[q * 16 + j]) : r + red[q * 16 + j];
  __syncthreads();
  return r;
}

// coupling coefficients of the next iteration from the logits a[k]: c = softmax over ALL out-capsules, per in-capsule j
template <int NO>
__device__ __forceinline__ void softmax_o(const float (&a)[NO], float (&c)[NO], int O, int h, int j, float* red) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < NO; ++k) if (h + k * NH < O) m = fmaxf(m, a[k]);
  m = block_col<true>(m, red, h, j);
  float z = 0.f;
#pragma unroll
  for (int k = 0; k < NO; ++k) { c[k] = (h + k * NH < O) ? __expf(a[k] - m) : 0.f; z += c[k]; }
  z = block_col<false>(z, red, h, j);
  const float rz = 1.f / z;
#pragma unroll
  for (int k = 0; k < NO; ++k) c[k] *= rz;
}

// prior [B][16][O][16] -> v [B][O][16] (nullable), norms [B][O] = |v| (nullable)
template <int NO>
__global__ void __launch_bounds__(RT) caps_route_fwd_kernel(const float* __restrict__ prior, int O, float* __restrict__ v_out,
                                                            float* __restrict__ norm_out) {
  extern __shared__ __align__(16) float sm[];
  float* V = sm;                      // [NO*NH][16] current v rows
  float* red = sm + NO * NH * 16;     // [NH][16]
  const int b = blockIdx.x, h = threadIdx.x >> 4, j = threadIdx.x & 15;
  const float invO = 1.f / (float)O;
  float a[NO], c[NO], p[16], w[16];
  // iteration 0: zero logits -> uniform coupling 1/O
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    const int o = h + k * NH;
    const bool live = o < O;
    load_prior(prior, b, j, O, o, live, p);
#pragma unroll
    for (int l = 0; l < 16; ++l) w[l] = p[l] * invO;
    const float v = squash_lane(reduce_scatter16(w, j));
    V[o * 16 + j] = v;
    __syncwarp();
    a[k] = dot16(p, V + o * 16);
  }
#pragma unroll
  for (int r = 1; r < 3; ++r) {
    softmax_o<NO>(a, c, O, h, j, red);
#pragma unroll
    for (int k = 0; k < NO; ++k) {
      const int o = h + k * NH;
      const bool live = o < O;
      load_prior(prior, b, j, O, o, live, p);
#pragma unroll
      for (int l = 0; l < 16; ++l) w[l] = p[l] * c[k];
      const float v = squash_lane(reduce_scatter16(w, j));
      if (r < 2) {
        __syncwarp();
        V[o * 16 + j] = v;
        __syncwarp();
        a[k] += dot16(p, V + o * 16);
      } else {
        const float n = sqrtf(hsum(v * v));
        if (live) {
          if (v_out != nullptr) v_out[((size_t)b * O + o) * 16 + j] = v;
          if (norm_out != nullptr && j == 0) norm_out[(size_t)b * O + o] = n;
        }
      }
    }
  }
}

// gradient of the routing wrt the prior.  g_v [B][O][16] and / or g_norm [B][O] (gradient of |v|); g_prior like prior.
template <int NO>
__global__ void __launch_bounds__(RT) caps_route_bwd_kernel(const float* __restrict__ prior, const float* __restrict__ g_v,
                                                            const float* __restrict__ g_norm, int O, float* __restrict__ g_prior) {
  extern __shared__ __align__(16) float sm[];
  constexpr int ROWS = NO * NH * 16;
  float* V0 = sm;
  float* V1 = V0 + ROWS;
  float* S0 = V1 + ROWS;
  float* S1 = S0 + ROWS;
  float* GS2 = S1 + ROWS;
  float* GS1 = GS2 + ROWS;
  float* GS0 = GS1 + ROWS;
  float* red = GS0 + ROWS;
  const int b = blockIdx.x, h = threadIdx.x >> 4, j = threadIdx.x & 15;
  const float invO = 1.f / (float)O;
  float a[NO], c1[NO], c2[NO], ga2[NO], ga1[NO], p[16], w[16];
  // ---- forward recompute
#pragma unroll
  for (int k = 0; k < NO; ++k) {                              // F0
    const int o = h + k * NH;
    load_prior(prior, b, j, O, o, o < O, p);
#pragma unroll
    for (int l = 0; l < 16; ++l) w[l] = p[l] * invO;
    const float s = reduce_scatter16(w, j);
    S0[o * 16 + j] = s;
    V0[o * 16 + j] = squash_lane(s);
    __syncwarp();
    a[k] = dot16(p, V0 + o * 16);
  }
  softmax_o<NO>(a, c1, O, h, j, red);
#pragma unroll
  for (int k = 0; k < NO; ++k) {                              // F1
    const int o = h + k * NH;
    load_prior(prior, b, j, O, o, o < O, p);
#pragma unroll
    for (int l = 0; l < 16; ++l) w[l] = p[l] * c1[k];
    const float s = reduce_scatter16(w, j);
    S1[o * 16 + j] = s;
    V1[o * 16 + j] = squash_lane(s);
    __syncwarp();
    a[k] += dot16(p, V1 + o * 16);
  }
  softmax_o<NO>(a, c2, O, h, j, red);
  // ---- F2 fused with the backward of iteration 2
  float dpart = 0.f;
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    const int o = h + k * NH;
    const bool live = o < O;
    load_prior(prior, b, j, O, o, live, p);
#pragma unroll
    for (int l = 0; l < 16; ++l) w[l] = p[l] * c2[k];
    const float s = reduce_scatter16(w, j);
    const float v = squash_lane(s);
    float gv = (live && g_v != nullptr) ? g_v[((size_t)b * O + o) * 16 + j] : 0.f;
    if (g_norm != nullptr) {
      const float n = sqrtf(hsum(v * v));
      if (live && n > 0.f) gv += g_norm[(size_t)b * O + o] * v / n;
    }
    GS2[o * 16 + j] = squash_bwd_lane(s, gv);
    __syncwarp();
    ga2[k] = dot16(p, GS2 + o * 16);                          // gc2 for now
    dpart += c2[k] * ga2[k];
  }
  {
    const float D2 = block_col<false>(dpart, red, h, j);
#pragma unroll
    for (int k = 0; k < NO; ++k) ga2[k] = c2[k] * (ga2[k] - D2);
  }
  // ---- backward of iteration 1
  dpart = 0.f;
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    const int o = h + k * NH;
    load_prior(prior, b, j, O, o, o < O, p);
#pragma unroll
    for (int l = 0; l < 16; ++l) w[l] = p[l] * ga2[k];
    const float gv1 = reduce_scatter16(w, j);
    GS1[o * 16 + j] = squash_bwd_lane(S1[o * 16 + j], gv1);
    __syncwarp();
    ga1[k] = dot16(p, GS1 + o * 16);                          // gc1 for now
    dpart += c1[k] * ga1[k];
  }
  {
    const float D1 = block_col<false>(dpart, red, h, j);
#pragma unroll
    for (int k = 0; k < NO; ++k) ga1[k] = ga2[k] + c1[k] * (ga1[k] - D1);
  }
  // ---- backward of iteration 0 and the gradient of the prior
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    const int o = h + k * NH;
    const bool live = o < O;
    load_prior(prior, b, j, O, o, live, p);
#pragma unroll
    for (int l = 0; l < 16; ++l) w[l] = p[l] * ga1[k];
    const float gv0 = reduce_scatter16(w, j);
    GS0[o * 16 + j] = squash_bwd_lane(S0[o * 16 + j], gv0);
    __syncwarp();
    if (live) {
      float4* dst = reinterpret_cast<float4*>(g_prior + (((size_t)b * CL + j) * O + o) * CL);
      const float k2 = c2[k], k1 = c1[k], q2 = ga2[k], q1 = ga1[k];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g2 = *reinterpret_cast<const float4*>(GS2 + o * 16 + 4 * q), g1 = *reinterpret_cast<const float4*>(GS1 + o * 16 + 4 * q),
                     g0 = *reinterpret_cast<const float4*>(GS0 + o * 16 + 4 * q), v1 = *reinterpret_cast<const float4*>(V1 + o * 16 + 4 * q),
                     v0 = *reinterpret_cast<const float4*>(V0 + o * 16 + 4 * q);
        float4 r;
        r.x = k2 * g2.x + q2 * v1.x + k1 * g1.x + q1 * v0.x + invO * g0.x;
        r.y = k2 * g2.y + q2 * v1.y + k1 * g1.y + q1 * v0.y + invO * g0.y;
        r.z = k2 * g2.z + q2 * v1.z + k1 * g1.z + q1 * v0.z + invO * g0.z;
        r.w = k2 * g2.w + q2 * v1.w + k1 * g1.w + q1 * v0.w + invO * g0.w;
        dst[q] = r;
      }
    }
  }
}

template <int NO>
int launch_route(const float* prior, const float* g_v, const float* g_norm, int B, int O, float* v, float* norms, float* g_prior,
                 bool bwd, cudaStream_t st) {
  if (!bwd) {
    const size_t smem = (size_t)(NO * NH * 16 + NH * 16) * sizeof(float);
    caps_route_fwd_kernel<NO><<<B, RT, smem, st>>>(prior, O, v, norms);
  } else {
    const size_t smem = (size_t)(7 * NO * NH * 16 + NH * 16) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      EKL_CHECK_CUDA(cudaFuncSetAttribute(caps_route_bwd_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    caps_route_bwd_kernel<NO><<<B, RT, smem, st>>>(prior, g_v, g_norm, O, g_prior);
  }
  EKL_LAUNCH_CHECK();
  return 0;
}

int dispatch_route(const float* prior, const float* g_v, const float* g_norm, int B, int O, float* v, float* norms, float* g_prior,
                   bool bwd, cudaStream_t st) {
  const int no = ekl_cdiv(O, NH);
  switch (no) {
    case 1: return launch_route<1>(prior, g_v, g_norm, B, O, v, norms, g_prior, bwd, st);
    case 2: return launch_route<2>(prior, g_v, g_norm, B, O, v, norms, g_prior, bwd, st);
    case 3: case 4: return launch_route<4>(prior, g_v, g_norm, B, O, v, norms, g_prior, bwd, st);
    case 5: case 6: case 7: return launch_route<7>(prior, g_v, g_norm, B, O, v, norms, g_prior, bwd, st);
    default: return ekl_fail(-1, "caps_route: at most 224 out-capsules (got %d)", O);
  }
}

}  // namespace

extern "C" int ekl_caps_route_supported(int I, int O, int Lh, int iters) {
  return I == CL && Lh == CL && iters == 3 && O >= 1 && O <= 7 * NH;
}

extern "C" int ekl_caps_route_fwd(const float* prior, int B, int I, int O, int Lh, int iters, float* v, float* norms, void* stream) {
  EKL_REQUIRE(prior != nullptr && (v != nullptr || norms != nullptr) && B > 0, "caps_route_fwd: null pointer argument");
  EKL_REQUIRE(ekl_caps_route_supported(I, O, Lh, iters), "caps_route: needs 16 in-capsules, out_length 16, 3 iterations, <= 224 out-capsules (I=%d L=%d it=%d O=%d)",
              I, Lh, iters, O);
  return dispatch_route(prior, nullptr, nullptr, B, O, v, norms, nullptr, false, (cudaStream_t)stream);
}

extern "C" int ekl_caps_route_bwd(const float* prior, const float* g_v, const float* g_norm, int B, int I, int O, int Lh, int iters,
                                  float* g_prior, void* stream) {
  EKL_REQUIRE(prior != nullptr && g_prior != nullptr && (g_v != nullptr || g_norm != nullptr) && B > 0, "caps_route_bwd: null pointer argument");
  EKL_REQUIRE(ekl_caps_route_supported(I, O, Lh, iters), "caps_route: needs 16 in-capsules, out_length 16, 3 iterations, <= 224 out-capsules (I=%d L=%d it=%d O=%d)",
              I, Lh, iters, O);
  return dispatch_route(prior, g_v, g_norm, B, O, nullptr, nullptr, g_prior, true, (cudaStream_t)stream);
}
