// Train-mode BatchNorm + activation family (HBM-bound; fp32 arithmetic, fp64 sums):
//   statistics  : per-(group, channel) sum / sum-of-squares, accumulated with fp64 red.global.add into ONE [groups][2][C]
//                 double buffer (zero on entry) by whoever produces the tensor: the conv epilogue, splitk_finish, or
//                 col_stats for a stored tensor.  No per-tile partial rows and no finalize launch: every consumer block
//                 derives mean / rstd of its own channels from the two sums.
//   forward     : z = gamma*(y-mean)*rstd + beta ; out = GLU(z) | LeakyReLU(z, 0.2) | z (+ residual); the blocks of the
//                 first row chunk also store mean / rstd for the backward pass and update the running statistics
//   backward    : dz = act'(z)*dout ; partial sums S1 = sum dz, S2 = sum dz*xhat ; dy = gamma*rstd*(dz - S1/n - xhat*S2/n)
//   data path   : the three streaming passes read their rows through a shared-memory ring of bulk async copies (struct
//                 Ring); layers with few rows per group run the backward as one launch (bn_act_bwd_small_kernel) and,
//                 after a split-K conv, the finishing pass and the forward as one launch (splitk_bn_act_fwd_kernel)
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

template <int ACT> struct ActVec { static constexpr int V = ACT == ACT_GLU ? 4 : 8; };

// ---------------------------------------------------------------- streaming passes: common structure
// Grid = (channel strips, row chunks per group, groups); block = 256 threads = RT row-threads x CT vector-threads; a
// thread owns VEC consecutive channels (of EACH half for GLU) of the rows rt, rt + RT, ...
//
// Data path (round 2, after the ncu captures in profiles/r02_bn_summary.md): the rows of a block's chunk stream through a
// ring of RING_S shared-memory stages filled by 1-D bulk async copies (cp.async.bulk ... mbarrier::complete_tx; thread 0
// issues them, one copy per tensor and tile when the block spans all channels -- the usual case -- else one per row
// segment).  The bytes in flight per SM are then RING_S tiles of every resident block (~48-64 KB per block) instead of
// what fits in registers next to the coefficients and partial sums (two rows per thread, ~37 KB per SM: the previous
// register-pipelined versions ran at 1.8-3.5 TB/s and spilled in the hot loop when widened).  Also from those captures:
//   * geometry is 32-bit and comes from the host (rows per group / per chunk, CT); groups are the grid's z dimension --
//     a third of the old kernels' executed instructions were 64-bit divisions, an fp64 reciprocal and the fp64
//     statistics of 8 channels recomputed by EVERY row-thread;
//   * per-channel coefficients are computed ONCE per block into shared memory (one thread per channel) while the first
//     tiles are in flight, and each thread then takes its 4..8 channels from there;
//   * GLU's gate is rcp(1 + ex2(b * sc' + sh')) with -log2(e) folded into sc' / sh' (FFMA, MUFU.EX2, FADD, MUFU.RCP);
//   * the backward sums live one per 128-byte line (EKL_BN_BWD_SPREAD doubles apart), so the fp64 reds of a block land
//     on as many L2 slices as it has channels (444 blocks x 256 reds onto 16 lines used to drain for ~8 us after the last
//     block had finished), and the block-level reduction issues ONE red per (channel, sum) from CT * NV threads.
constexpr int BWD_SPREAD = 16;                  // == EKL_BN_BWD_SPREAD (include/ekl_b200.h)
constexpr float NEG_LOG2E = -1.4426950408889634f;
// RING_S stages of the shared-memory ring, RPT rows per thread and tile (a tile is RT * RPT rows): template parameters
// <RING_S, RPT> of the kernels below; measured choice in ring_cfg().

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// sigmoid of the pre-scaled gate argument t = -log2(e) * z:  1 / (1 + 2^t)   (t -> +inf gives 0, t -> -inf gives 1)
__device__ __forceinline__ float gate_sigmoid(float t) { return rcp_approx(1.f + ex2_approx(t)); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int VEC> struct SmemIO;
template <> struct SmemIO<8> { static __device__ __forceinline__ void load(uint32_t a, float* f) { unpack8(lds128(a), f); } };
template <> struct SmemIO<4> {
  static __device__ __forceinline__ void load(uint32_t a, float* f) {
    const uint2 u = lds64(a);
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  }
};

// Producer side of the ring (thread 0 of a block, once per tile): see struct Ring below for the stage layout.
template <int VEC, int NH, int RING_S, int RPT>
__device__ __noinline__ void ring_issue(int j, int Mg, int per, int CT, int Cy, int has1, const bf16* t0, const bf16* t1,
                                        uint32_t sbase, uint32_t bars, int stage_bytes, int RT, int nrows, int g, int k,
                                        int strip0, int segv) {
  const int Co = NH == 2 ? Cy / 2 : Cy;
  const int seg = CT * VEC, TR = RT * RPT;
  const int s = j % RING_S;
  const uint32_t st = sbase + (uint32_t)s * stage_bytes, fb = bars + 8 * s;
  const int first = j * TR;
  const int rows = nrows - first < TR ? nrows - first : TR;
  const int64_t grow = (int64_t)g * Mg + k * per + first;
  const uint32_t y_bytes = (uint32_t)TR * NH * seg * 2;
  if (gridDim.x == 1) {
    const uint32_t b0 = (uint32_t)rows * Cy * 2, b1 = has1 ? (uint32_t)rows * Co * 2 : 0u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(b0 + b1) : "memory");
    bulk_g2s(st, t0 + grow * Cy, b0, fb);
    if (has1) bulk_g2s(st + y_bytes, t1 + grow * Co, b1, fb);
  } else {
    const uint32_t sb = (uint32_t)segv * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)rows * sb * (NH + has1)) : "memory");
#pragma unroll 1
    for (int r = 0; r < rows; ++r) {
      bulk_g2s(st + (uint32_t)(r * NH * seg * 2), t0 + (grow + r) * Cy + strip0, sb, fb);
      if (NH == 2) bulk_g2s(st + (uint32_t)((r * NH + 1) * seg * 2), t0 + (grow + r) * Cy + Co + strip0, sb, fb);
      if (has1) bulk_g2s(st + y_bytes + (uint32_t)(r * seg * 2), t1 + (grow + r) * Co + strip0, sb, fb);
    }
  }
}

// One block's view of a streaming pass.  Up to two tensors stream through the ring: tensor 0 is y ([rows][Cy], NH halves
// of the block's channel strip per row), tensor 1 (optional) has Co channels per row (dout / residual).  A stage holds
// TR = RT * RPT rows: [TR][NH][seg] of y, then [TR][seg] of tensor 1 (seg = CT * VEC channels; with one strip seg == Co
// and a row is laid out exactly as in global memory, so a tile is ONE contiguous copy per tensor).
// The struct keeps only what the consumer loop needs; the producer (thread 0, once per tile) re-derives its geometry.
template <int VEC, int NH, int RING_S, int RPT>
struct Ring {
  uint32_t sbase, bars;           // shared-window addresses: stage 0 / full[0] (empty[s] = bars + 8 * (RING_S + s))
  uint32_t off_a, step_a;         // byte offset of this thread's y vector in row rt of a stage / stride of RT rows
  uint32_t off_1, step_1;         // the same for tensor 1
  uint32_t seg2;                  // bytes between the two halves of a row (GLU gate half)
  int stage_bytes;
  int rt, RT, nrows, ntiles;
  int g, k, ct, strip0, segv;     // statistics group, chunk, vector-thread; strip origin / valid width (channels, per half)
  bool active;

  __device__ __forceinline__ void init(int Mg, int per, int CT, int Cy, bool rev, bool has1, uint8_t* smem_ring) {
    const int Co = NH == 2 ? Cy / 2 : Cy;
    RT = 256 / CT;
    rt = (int)threadIdx.x / CT;
    ct = (int)threadIdx.x - rt * CT;
    k = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    g = rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int seg = CT * VEC;
    strip0 = (int)blockIdx.x * seg;
    segv = Co - strip0 < seg ? Co - strip0 : seg;
    active = ct * VEC < segv && rt < RT;
    const int r0 = k * per;
    const int r1 = r0 + per < Mg ? r0 + per : Mg;
    nrows = r1 > r0 ? r1 - r0 : 0;
    const int TR = RT * RPT;
    ntiles = (nrows + TR - 1) / TR;
    bars = smem_u32(smem_ring);
    sbase = bars + 128;
    const int y_bytes = TR * NH * seg * 2;
    stage_bytes = y_bytes + (has1 ? TR * seg * 2 : 0);
    seg2 = (uint32_t)seg * 2;
    off_a = (uint32_t)((rt * NH * seg + ct * VEC) * 2);
    step_a = (uint32_t)(RT * NH * seg * 2);
    off_1 = (uint32_t)(y_bytes + (rt * seg + ct * VEC) * 2);
    step_1 = (uint32_t)(RT * seg * 2);
  }
  static __device__ __forceinline__ void bar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
  }
  __device__ __forceinline__ void barriers() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < RING_S; ++s) { bar_init(bars + 8 * s, 1); bar_init(bars + 8 * (RING_S + s), 8); }
      fence_barrier_init();
    }
    __syncthreads();
  }
  // thread 0: start the copies of tile j into stage j % RING_S (an out-of-line call: the producer's address arithmetic must
  // not hold registers across the consumer loop of all 256 threads)
  __device__ __forceinline__ void issue(int j, int Mg, int per, int CT, int Cy, bool has1, const bf16* t0, const bf16* t1) const {
    ring_issue<VEC, NH, RING_S, RPT>(j, Mg, per, CT, Cy, has1 ? 1 : 0, t0, t1, sbase, bars, stage_bytes, RT, nrows, g, k, strip0, segv);
  }
  static __device__ __forceinline__ void wait_bar(uint32_t a, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(a), "r"(parity) : "memory");
      if (!ok && clock64() - t0 > 8000000000LL) __trap();      // a protocol bug must trap, not hang the GPU box
    }
  }
  __device__ __forceinline__ void wait_tile(int j) const { wait_bar(bars + 8 * (j % RING_S), (uint32_t)(j / RING_S) & 1u); }
  // every warp releases the stage; thread 0 refills it with tile j + RING_S once all 8 warps have
  __device__ __forceinline__ void release(int j, int Mg, int per, int CT, int Cy, bool has1, const bf16* t0, const bf16* t1) const {
    __syncwarp();
    const uint32_t eb = bars + 8 * (RING_S + j % RING_S);
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(eb) : "memory");
    if (threadIdx.x == 0 && j + RING_S < ntiles) {
      wait_bar(eb, (uint32_t)(j / RING_S) & 1u);
      issue(j + RING_S, Mg, per, CT, Cy, has1, t0, t1);
    }
  }
  __device__ __forceinline__ uint32_t stage(int j) const { return sbase + (uint32_t)(j % RING_S) * stage_bytes; }
};

// ---------------------------------------------------------------- forward
// sums != null (training): statistics from the fp64 sums; mean_io / rstd_io [groups][Cy] are WRITTEN by chunk 0 of each
// group (saved for the backward pass) and the first block of every strip applies one running-statistics momentum update
// per group, in group order (one per reference forward call).  sums == null (inference): mean_io / rstd_io are inputs.
// Dynamic shared memory: [NH * CT * VEC] float2 (scale, shift; GLU gate half pre-multiplied by -log2 e) | ring.
template <int ACT, int RING_S, int RPT>
__global__ void __launch_bounds__(256, 3) bn_act_fwd_kernel(const bf16* __restrict__ y, int Mg, int per, int CT, int Cy,
                                                            const double* __restrict__ sums, double inv_n, float eps, float momentum,
                                                            float* mean_io, float* rstd_io, float* running_mean,
                                                            float* running_var,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const bf16* __restrict__ residual, bf16* __restrict__ out, int rev,
                                                            int coef_bytes) {
  constexpr int VEC = ActVec<ACT>::V;
  constexpr int NH = ACT == ACT_GLU ? 2 : 1;
  using IO = VecIO<VEC>;
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  float2* coef = reinterpret_cast<float2*>(smem_dyn);
  Ring<VEC, NH, RING_S, RPT> R;
  const bool has1 = residual != nullptr;
  R.init(Mg, per, CT, Cy, rev != 0, has1, smem_dyn + coef_bytes);
  R.barriers();
  if (threadIdx.x == 0)
    for (int j = 0; j < RING_S && j < R.ntiles; ++j) R.issue(j, Mg, per, CT, Cy, has1, y, residual);
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy, groups = (int)gridDim.z, nstrip = CT * VEC;
  for (int j = threadIdx.x; j < NH * nstrip; j += 256) {
    const int h = j >= nstrip ? 1 : 0;
    const int cc = R.strip0 + j - h * nstrip;
    if (cc >= Co) continue;
    const int c = h * Co + cc;
    float m, r;
    if (sums != nullptr) {
      const BnStat st = bn_stat(sums, R.g, Cy, c, inv_n, eps);
      m = st.mean; r = st.rstd;
      if (R.k == 0) { mean_io[R.g * Cy + c] = m; rstd_io[R.g * Cy + c] = r; }
      if (running_mean != nullptr && blockIdx.y == 0 && blockIdx.z == 0) {
        const float n = (float)Mg;
        float rm = running_mean[c], rv = running_var[c];
        for (int g = 0; g < groups; ++g) {
          const BnStat sg = bn_stat(sums, g, Cy, c, inv_n, eps);
          const float unb = n > 1.f ? sg.var * n / (n - 1.f) : sg.var;
          rm = (1.f - momentum) * rm + momentum * sg.mean;
          rv = (1.f - momentum) * rv + momentum * unb;
        }
        running_mean[c] = rm; running_var[c] = rv;
      }
    } else {
      m = mean_io[R.g * Cy + c]; r = rstd_io[R.g * Cy + c];
    }
    float s = gamma[c] * r, b = beta[c] - m * s;
    if (ACT == ACT_GLU && h == 1) { s *= NEG_LOG2E; b *= NEG_LOG2E; }
    coef[j] = make_float2(s, b);
  }
  __syncthreads();
  float sc[VEC], sh[VEC], sc2[VEC], sh2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float2 v = coef[R.ct * VEC + i];
    sc[i] = v.x; sh[i] = v.y;
    if (ACT == ACT_GLU) { const float2 w = coef[nstrip + R.ct * VEC + i]; sc2[i] = w.x; sh2[i] = w.y; }
  }
  bf16* po = out + ((int64_t)R.g * Mg + R.k * per + R.rt) * Co + R.strip0 + R.ct * VEC;
  const int64_t stepO = (int64_t)R.RT * Co;
  int left = R.active ? R.nrows - R.rt : 0;          // rows of this thread's residue class still to come (> 0: row valid)
  for (int j = 0; j < R.ntiles; ++j) {
    R.wait_tile(j);
    const uint32_t st = R.stage(j);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      if (left - u * R.RT > 0) {
        const uint32_t p = st + R.off_a + u * R.step_a;
        float a[VEC], o[VEC];
        SmemIO<VEC>::load(p, a);
        if (ACT == ACT_GLU) {
          float b[VEC];
          SmemIO<VEC>::load(p + R.seg2, b);
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[i] = fmaf(a[i], sc[i], sh[i]) * gate_sigmoid(fmaf(b[i], sc2[i], sh2[i]));
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float z = fmaf(a[i], sc[i], sh[i]);
            o[i] = ACT == ACT_LRELU ? (z > 0.f ? z : 0.2f * z) : (ACT == ACT_RELU ? fmaxf(z, 0.f) : z);
          }
        }
        if (has1) {
          float q[VEC];
          SmemIO<VEC>::load(st + R.off_1 + u * R.step_1, q);
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[i] += q[i];
        }
        *reinterpret_cast<typename IO::T*>(po + u * stepO) = IO::pack(o);
      }
    }
    po += RPT * stepO;
    left -= RPT * R.RT;
    R.release(j, Mg, per, CT, Cy, has1, y, residual);
  }
}

// ---------------------------------------------------------------- backward, pass 1: partial sums
// S1 = sum dz, S2 = sum dz * xhat with xhat = y*rstd - mean*rstd (one FMA; coefficients rs / nmr), accumulated into
// sums[((g*2 + which) * Cy + c) * BWD_SPREAD] (fp64 reds, one 128-byte line per sum).
// Dynamic shared memory: [NH * CT * VEC] float4 (scale, shift, rstd, -mean*rstd) | ring (re-used for the block reduction).
template <int ACT, int RING_S, int RPT>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_reduce_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout,
                                                                   int Mg, int per, int CT, int Cy,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   double* __restrict__ sums, int rev, int coef_bytes) {
  constexpr int VEC = ActVec<ACT>::V;
  constexpr int NH = ACT == ACT_GLU ? 2 : 1;
  constexpr int NV = 2 * NH * VEC;                 // sums a thread carries: [S1 a | S2 a | S1 b | S2 b]
  using IO = VecIO<VEC>;
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  float4* coef = reinterpret_cast<float4*>(smem_dyn);
  Ring<VEC, NH, RING_S, RPT> R;
  R.init(Mg, per, CT, Cy, rev != 0, true, smem_dyn + coef_bytes);
  R.barriers();
  if (threadIdx.x == 0)
    for (int j = 0; j < RING_S && j < R.ntiles; ++j) R.issue(j, Mg, per, CT, Cy, true, y, dout);
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy, nstrip = CT * VEC;
  for (int j = threadIdx.x; j < NH * nstrip; j += 256) {
    const int h = j >= nstrip ? 1 : 0;
    const int cc = R.strip0 + j - h * nstrip;
    if (cc >= Co) continue;
    const int c = h * Co + cc;
    const float mu = mean[R.g * Cy + c], r = rstd[R.g * Cy + c];
    float s = gamma[c] * r, b = beta[c] - mu * s;
    if (ACT == ACT_GLU && h == 1) { s *= NEG_LOG2E; b *= NEG_LOG2E; }
    coef[j] = make_float4(s, b, r, -mu * r);
  }
  __syncthreads();
  float sc[VEC], sh[VEC], sc2[VEC], sh2[VEC], rs[VEC], nmr[VEC], rs2[VEC], nmr2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 v = coef[R.ct * VEC + i];
    sc[i] = v.x; sh[i] = v.y; rs[i] = v.z; nmr[i] = v.w;
    if (ACT == ACT_GLU) { const float4 w = coef[nstrip + R.ct * VEC + i]; sc2[i] = w.x; sh2[i] = w.y; rs2[i] = w.z; nmr2[i] = w.w; }
  }
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  int left = R.active ? R.nrows - R.rt : 0;
  for (int j = 0; j < R.ntiles; ++j) {
    R.wait_tile(j);
    const uint32_t st = R.stage(j);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      if (left - u * R.RT > 0) {
        const uint32_t p = st + R.off_a + u * R.step_a;
        float ya[VEC], yb[VEC], d[VEC];
        SmemIO<VEC>::load(p, ya);
        if (ACT == ACT_GLU) SmemIO<VEC>::load(p + R.seg2, yb);
        SmemIO<VEC>::load(st + R.off_1 + u * R.step_1, d);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float z = fmaf(ya[i], sc[i], sh[i]);
          float dza;
          if (ACT == ACT_GLU) {
            const float s = gate_sigmoid(fmaf(yb[i], sc2[i], sh2[i]));
            dza = d[i] * s;
            const float dzb = dza * z * (1.f - s);
            acc[2 * VEC + i] += dzb;
            acc[3 * VEC + i] = fmaf(dzb, fmaf(yb[i], rs2[i], nmr2[i]), acc[3 * VEC + i]);
          } else if (ACT == ACT_LRELU) {
            dza = z > 0.f ? d[i] : 0.2f * d[i];
          } else if (ACT == ACT_RELU) {
            dza = z > 0.f ? d[i] : 0.f;
          } else {
            dza = d[i];
          }
          acc[i] += dza;
          acc[VEC + i] = fmaf(dza, fmaf(ya[i], rs[i], nmr[i]), acc[VEC + i]);
        }
      }
    }
    left -= RPT * R.RT;
    R.release(j, Mg, per, CT, Cy, true, y, dout);
  }
  // block reduction over the row-threads (the ring is drained: its memory holds the partials): thread (rt, ct) parks its
  // NV sums, then thread j = ct * NV + v adds the RT partials of ONE (vector-thread, sum) pair and issues one fp64 red
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_dyn + coef_bytes + 128);
#pragma unroll
  for (int v = 0; v < NV; ++v) red[threadIdx.x * NV + v] = acc[v];
  __syncthreads();
  for (int j = threadIdx.x; j < CT * NV; j += 256) {
    const int ct = j / NV, v = j - ct * NV;
    if (ct * VEC >= R.segv) continue;
    float a = 0.f;
    for (int k = 0; k < R.RT; ++k) a += red[(k * CT + ct) * NV + v];
    const int h = v / (2 * VEC), which = (v / VEC) & 1, i = v % VEC;
    const int c = h * Co + R.strip0 + ct * VEC + i;
    stat_add(sums + ((size_t)(R.g * 2 + which) * Cy + c) * BWD_SPREAD, a);
  }
}

// ---------------------------------------------------------------- backward, pass 2: dy
// dy = k*(dz - S1/n - xhat*S2/n), k = gamma*rstd  ==  k*dz + A*y + Bc  with  A = -k*rstd*S2/n,  Bc = -k*S1/n - A*mean
// Dynamic shared memory: [NH * CT * VEC] float4 (scale, shift, A, Bc) | ring.
template <int ACT, int RING_S, int RPT>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_apply_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout,
                                                                  int Mg, int per, int CT, int Cy, float inv_n,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const double* __restrict__ sums,
                                                                  float* dgamma, float* dbeta, bf16* __restrict__ dy, int rev,
                                                                  int coef_bytes) {
  constexpr int VEC = ActVec<ACT>::V;
  constexpr int NH = ACT == ACT_GLU ? 2 : 1;
  using IO = VecIO<VEC>;
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  float4* coef = reinterpret_cast<float4*>(smem_dyn);
  Ring<VEC, NH, RING_S, RPT> R;
  R.init(Mg, per, CT, Cy, rev != 0, true, smem_dyn + coef_bytes);
  R.barriers();
  if (threadIdx.x == 0)
    for (int j = 0; j < RING_S && j < R.ntiles; ++j) R.issue(j, Mg, per, CT, Cy, true, y, dout);
  const int Co = ACT == ACT_GLU ? Cy / 2 : Cy, groups = (int)gridDim.z, nstrip = CT * VEC;
  for (int j = threadIdx.x; j < NH * nstrip; j += 256) {
    const int h = j >= nstrip ? 1 : 0;
    const int cc = R.strip0 + j - h * nstrip;
    if (cc >= Co) continue;
    const int c = h * Co + cc;
    const float mu = mean[R.g * Cy + c], r = rstd[R.g * Cy + c];
    const float s = gamma[c] * r, b = beta[c] - mu * s;
    const float S1 = (float)sums[((size_t)(R.g * 2 + 0) * Cy + c) * BWD_SPREAD];
    const float S2 = (float)sums[((size_t)(R.g * 2 + 1) * Cy + c) * BWD_SPREAD];
    const float A = -s * r * S2 * inv_n;
    const float Bc = -s * S1 * inv_n - A * mu;
    const bool gate = ACT == ACT_GLU && h == 1;
    coef[j] = make_float4(gate ? s * NEG_LOG2E : s, gate ? b * NEG_LOG2E : b, A, Bc);
    // parameter gradients (accumulated): dgamma[c] += sum_g S2, dbeta[c] += sum_g S1 -- one thread per channel
    if (blockIdx.y == 0 && blockIdx.z == 0 && (dgamma != nullptr || dbeta != nullptr)) {
      double tb = 0.0, tg = 0.0;
      for (int g = 0; g < groups; ++g) {
        tb += sums[((size_t)(g * 2 + 0) * Cy + c) * BWD_SPREAD];
        tg += sums[((size_t)(g * 2 + 1) * Cy + c) * BWD_SPREAD];
      }
      if (dgamma != nullptr) dgamma[c] += (float)tg;
      if (dbeta != nullptr) dbeta[c] += (float)tb;
    }
  }
  __syncthreads();
  // k2: the true gamma*rstd of the gate half (its stored scale carries the -log2 e factor)
  float sc[VEC], sh[VEC], sc2[VEC], sh2[VEC], A[VEC], Bc[VEC], A2[VEC], Bc2[VEC], k2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 v = coef[R.ct * VEC + i];
    sc[i] = v.x; sh[i] = v.y; A[i] = v.z; Bc[i] = v.w;
    if (ACT == ACT_GLU) {
      const float4 w = coef[nstrip + R.ct * VEC + i];
      sc2[i] = w.x; sh2[i] = w.y; A2[i] = w.z; Bc2[i] = w.w; k2[i] = w.x * (1.f / NEG_LOG2E);
    }
  }
  bf16* po = dy + ((int64_t)R.g * Mg + R.k * per + R.rt) * Cy + R.strip0 + R.ct * VEC;
  const int64_t stepY = (int64_t)R.RT * Cy;
  int left = R.active ? R.nrows - R.rt : 0;
  for (int j = 0; j < R.ntiles; ++j) {
    R.wait_tile(j);
    const uint32_t st = R.stage(j);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      if (left - u * R.RT > 0) {
        const uint32_t p = st + R.off_a + u * R.step_a;
        float ya[VEC], yb[VEC], d[VEC], oa[VEC], ob[VEC];
        SmemIO<VEC>::load(p, ya);
        if (ACT == ACT_GLU) SmemIO<VEC>::load(p + R.seg2, yb);
        SmemIO<VEC>::load(st + R.off_1 + u * R.step_1, d);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float z = fmaf(ya[i], sc[i], sh[i]);
          float dza;
          if (ACT == ACT_GLU) {
            const float s = gate_sigmoid(fmaf(yb[i], sc2[i], sh2[i]));
            dza = d[i] * s;
            const float dzb = dza * z * (1.f - s);
            ob[i] = fmaf(k2[i], dzb, fmaf(A2[i], yb[i], Bc2[i]));
          } else if (ACT == ACT_LRELU) {
            dza = z > 0.f ? d[i] : 0.2f * d[i];
          } else if (ACT == ACT_RELU) {
            dza = z > 0.f ? d[i] : 0.f;
          } else {
            dza = d[i];
          }
          oa[i] = fmaf(sc[i], dza, fmaf(A[i], ya[i], Bc[i]));
        }
        *reinterpret_cast<typename IO::T*>(po + u * stepY) = IO::pack(oa);
        if (ACT == ACT_GLU) *reinterpret_cast<typename IO::T*>(po + u * stepY + Co) = IO::pack(ob);
      }
    }
    po += RPT * stepY;
    left -= RPT * R.RT;
    R.release(j, Mg, per, CT, Cy, true, y, dout);
  }
}

// dz (pre-activation gradient) for the VEC (GLU: VEC+VEC) channels a thread owns; sc2 / sh2 are the gate half's
// coefficients pre-multiplied by -log2(e) (see gate_sigmoid).  Used by the single-launch small-layer kernel.
template <int ACT, int VEC>
__device__ __forceinline__ void act_bwd(const float* ya, const float* yb, const float* d, const float* sc, const float* sh,
                                        const float* sc2, const float* sh2, float* dza, float* dzb) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (ACT == ACT_GLU) {
      const float z = fmaf(ya[i], sc[i], sh[i]);
      const float s = gate_sigmoid(fmaf(yb[i], sc2[i], sh2[i]));
      dza[i] = d[i] * s;
      dzb[i] = dza[i] * z * (1.f - s);
    } else if (ACT == ACT_LRELU) {
      dza[i] = fmaf(ya[i], sc[i], sh[i]) > 0.f ? d[i] : 0.2f * d[i];
    } else if (ACT == ACT_RELU) {
      dza[i] = fmaf(ya[i], sc[i], sh[i]) > 0.f ? d[i] : 0.f;
    } else {
      dza[i] = d[i];
    }
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
  float sc[VEC], shf[VEC], sc2[VEC], sh2[VEC], g2s[VEC], g2h[VEC], rs[VEC], nmr[VEC], rs2[VEC], nmr2[VEC];
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
      g2s[i] = sc2[i] * NEG_LOG2E; g2h[i] = sh2[i] * NEG_LOG2E;       // gate argument pre-scaled for gate_sigmoid
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
      act_bwd<ACT, VEC>(ya, yb, d, sc, shf, g2s, g2h, dza, dzb);
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
      act_bwd<ACT, VEC>(ya, yb, d, sc, shf, g2s, g2h, dza, dzb);
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

// ---------------------------------------------------------------- split-K finish + BatchNorm forward in one launch
// The split-K convolutions of the step (4x4 / 8x8 discriminator tails) are all followed by train-mode BatchNorm +
// LeakyReLU on a few hundred rows per group: finish (fp32 scratch -> bf16 y + statistics), then the normalise pass, were
// two launch-latency-bound kernels per layer (8-17 us + 7-19 us on the discriminator branch's critical path).  Here a block
// owns a strip of 32 channels of ONE group for all its rows: sweep 1 rounds the scratch to bf16 y, re-zeroes it and
// takes the statistics of the rounded values (same contract as the conv epilogue / splitk_finish); sweep 2 re-reads y
// (L1 / L2) and writes out = act(gamma * (y - mean) * rstd + beta).  mean / rstd are stored for the backward pass.  The
// running statistics need one momentum update per group IN GROUP ORDER: each block leaves its (mean, var) in global
// memory and takes a ticket per strip; the last block of a strip applies the updates of all groups.
// aux: [strips] tickets (uint32, ZERO on entry, left zero) then [groups][C] floats of variances.
// Block = 256 threads = 8 vector-threads (4 channels, 16-byte fp32 reads) x 32 row-threads; grid = (C / 32, groups).
template <int ACT>
__global__ void __launch_bounds__(256) splitk_bn_act_fwd_kernel(float* __restrict__ scratch, int Mg, int C, float eps, float momentum,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                float* running_mean, float* running_var, unsigned int* aux,
                                                                bf16* __restrict__ y, bf16* __restrict__ out) {
  constexpr int NV = 8;                      // [sum x 4 | sum of squares x 4]
  __shared__ float sh[8 * SM_CT * NV];
  __shared__ float tot[SM_CT * NV];
  __shared__ int last_block;
  const int ct = threadIdx.x % SM_CT, rt = threadIdx.x / SM_CT;
  const int g = blockIdx.y, groups = (int)gridDim.y;
  const int c0 = ((int)blockIdx.x * SM_CT + ct) * 4;
  float* sg = scratch + (int64_t)g * Mg * C;
  bf16* yg = y + (int64_t)g * Mg * C;
  bf16* og = out + (int64_t)g * Mg * C;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  constexpr int U = 4;
  for (int rb = rt; rb < Mg; rb += U * 32) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * 32;
      if (r < Mg) v[u] = *reinterpret_cast<const float4*>(sg + (int64_t)r * C + c0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * 32;
      if (r >= Mg) break;
      *reinterpret_cast<float4*>(sg + (int64_t)r * C + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint2 o = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
      *reinterpret_cast<uint2*>(yg + (int64_t)r * C + c0) = o;
      const float f0 = bf16_lo(o.x), f1 = bf16_hi(o.x), f2 = bf16_lo(o.y), f3 = bf16_hi(o.y);
      acc[0] += f0; acc[1] += f1; acc[2] += f2; acc[3] += f3;
      acc[4] += f0 * f0; acc[5] += f1 * f1; acc[6] += f2 * f2; acc[7] += f3 * f3;
    }
  }
  strip_reduce<NV>(acc, sh, tot);
  const double inv_n = 1.0 / (double)Mg;
  float sc[4], sf[4];
  float* var_buf = reinterpret_cast<float*>(aux + gridDim.x);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double m = (double)acc[i] * inv_n;
    double v = (double)acc[4 + i] * inv_n - m * m;
    v = v < 0.0 ? 0.0 : v;
    const float mean = (float)m, rstd = 1.f / sqrtf((float)v + eps);
    const int c = c0 + i;
    if (rt == 0) { mean_out[g * C + c] = mean; rstd_out[g * C + c] = rstd; var_buf[g * C + c] = (float)v; }
    sc[i] = gamma[c] * rstd; sf[i] = beta[c] - mean * sc[i];
  }
  if (running_mean != nullptr) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_block = atomicAdd(aux + blockIdx.x, 1u) == (unsigned)(groups - 1);
    __syncthreads();
    if (last_block) {
      __threadfence();
      if (threadIdx.x < SM_CT * 4) {
        const int c = (int)blockIdx.x * SM_CT * 4 + threadIdx.x;
        const float n = (float)Mg;
        float rm = running_mean[c], rv = running_var[c];
        for (int gg = 0; gg < groups; ++gg) {
          const float m = __ldcg(mean_out + gg * C + c), v = __ldcg(var_buf + gg * C + c);
          const float unb = n > 1.f ? v * n / (n - 1.f) : v;
          rm = (1.f - momentum) * rm + momentum * m;
          rv = (1.f - momentum) * rv + momentum * unb;
        }
        running_mean[c] = rm; running_var[c] = rv;
      }
      if (threadIdx.x == 0) aux[blockIdx.x] = 0u;        // tickets are left zero for the next use of the arena slice
    }
  }
  for (int rb = rt; rb < Mg; rb += U * 32) {
    uint2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * 32;
      if (r < Mg) v[u] = *reinterpret_cast<const uint2*>(yg + (int64_t)r * C + c0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * 32;
      if (r >= Mg) break;
      float z[4] = {bf16_lo(v[u].x), bf16_hi(v[u].x), bf16_lo(v[u].y), bf16_hi(v[u].y)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        z[i] = fmaf(z[i], sc[i], sf[i]);
        z[i] = ACT == ACT_LRELU ? (z[i] > 0.f ? z[i] : 0.2f * z[i]) : (ACT == ACT_RELU ? fmaxf(z[i], 0.f) : z[i]);
      }
      *reinterpret_cast<uint2*>(og + (int64_t)r * C + c0) = make_uint2(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]));
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

// eligibility of the fused finish + BatchNorm forward for an [M][C] split-K output with `groups` statistics groups
int ekl_splitk_bn_fusable(int64_t M, int C, int groups, int act) {
  if (groups <= 0 || M % groups != 0 || C % (SM_CT * 4) != 0) return 0;
  if (!(act == ACT_NONE || act == ACT_LRELU || act == ACT_RELU)) return 0;
  return M / groups <= 768;
}

// aux: C / 32 uint32 tickets + groups * C floats, ZERO on entry (see splitk_bn_act_fwd_kernel)
int ekl_splitk_bn_act_fwd(float* scratch, int64_t M, int C, int groups, float eps, float momentum, float* mean, float* rstd,
                          float* running_mean, float* running_var, const float* gamma, const float* beta, int act, void* y,
                          void* out, void* aux, cudaStream_t st) {
  EKL_REQUIRE(ekl_splitk_bn_fusable(M, C, groups, act), "splitk_bn_act_fwd: unsupported shape M=%lld C=%d groups=%d act=%d",
              (long long)M, C, groups, act);
  EKL_REQUIRE(scratch && mean && rstd && gamma && beta && y && out && aux, "splitk_bn_act_fwd: null pointer argument");
  const dim3 grid(C / (SM_CT * 4), groups);
  const int Mg = (int)(M / groups);
  switch (act) {
    case ACT_LRELU:
      splitk_bn_act_fwd_kernel<ACT_LRELU><<<grid, 256, 0, st>>>(scratch, Mg, C, eps, momentum, gamma, beta, mean, rstd, running_mean,
                                                                running_var, (unsigned int*)aux, (bf16*)y, (bf16*)out);
      break;
    case ACT_RELU:
      splitk_bn_act_fwd_kernel<ACT_RELU><<<grid, 256, 0, st>>>(scratch, Mg, C, eps, momentum, gamma, beta, mean, rstd, running_mean,
                                                               running_var, (unsigned int*)aux, (bf16*)y, (bf16*)out);
      break;
    default:
      splitk_bn_act_fwd_kernel<ACT_NONE><<<grid, 256, 0, st>>>(scratch, Mg, C, eps, momentum, gamma, beta, mean, rstd, running_mean,
                                                               running_var, (unsigned int*)aux, (bf16*)y, (bf16*)out);
  }
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

// Launch geometry of a streaming pass: grid (channel strips, row chunks per group, groups), rows per chunk, vector
// threads per block; dynamic shared memory = `ncoef` floats per channel of the block's strip + the tile ring
// (RING_S stages of RT * RPT rows of y [+ the second tensor]).
// ring shape <stages, rows per thread and tile>: measured on config 2 (profiles/r02_bn_summary.md) <4, 2> 7.73-7.85 ms/step,
// <3, 4> 7.88, <6, 2> 7.95, <8, 1> 7.87-7.98
constexpr int BN_RING_S = 4, BN_RPT = 2;

struct StreamGeom { dim3 grid; int Mg, per, CT, coef_bytes; size_t smem; };
static int stream_geom(int64_t M, int Co, int groups, int act, int ncoef, bool second, StreamGeom* o) {
  const int RING_S = BN_RING_S, RPT = BN_RPT;
  const int VEC = act_vec(act), NH = act == ACT_GLU ? 2 : 1;
  const int nvec = Co / VEC;
  EKL_REQUIRE(M / groups < (int64_t)1 << 31, "bn_act: more than 2^31 rows per group");
  EKL_REQUIRE(groups <= 65535, "bn_act: too many statistics groups");
  dim3 g2;
  const int chunks = grid_rows(M, nvec, groups, &g2);
  o->Mg = (int)(M / groups);
  o->per = (o->Mg + chunks - 1) / chunks;
  o->CT = nvec < 256 ? nvec : 256;
  o->grid = dim3(g2.x, (unsigned)chunks, (unsigned)groups);
  const int RT = 256 / o->CT;
  const size_t seg = (size_t)o->CT * VEC;
  o->coef_bytes = (int)((NH * seg * ncoef * sizeof(float) + 127) / 128 * 128);
  const size_t stage = (size_t)RT * RPT * seg * 2 * (NH + (second ? 1 : 0));
  o->smem = o->coef_bytes + 128 + RING_S * stage;
  return 0;
}

// dynamic shared memory above 48 KB needs an opt-in per kernel; done once per instantiation
template <typename K>
static int allow_smem(K kern, size_t bytes) {
  EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  (void)bytes;
  return 0;
}
#define EKL_BN_LAUNCH(KERN, SG, ST, ...)                                          \
  do {                                                                            \
    static bool attr_done = false;                                                \
    auto kern_ = KERN<A, BN_RING_S, BN_RPT>;                                      \
    if (!attr_done) { if (int rc = allow_smem(kern_, (SG).smem)) return rc; attr_done = true; } \
    kern_<<<(SG).grid, 256, (SG).smem, ST>>>(__VA_ARGS__);                        \
  } while (0)

extern "C" int ekl_bn_act_fwd(const void* y, int64_t M, int Cy, int groups, const double* sums, float eps, float momentum,
                              float* mean, float* rstd, float* running_mean, float* running_var, const float* gamma,
                              const float* beta, int act, const void* residual, void* out, void* stream) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  EKL_REQUIRE(groups > 0 && Co % 8 == 0 && M % groups == 0, "bn_act_fwd: bad shape Cy=%d", Cy);
  EKL_REQUIRE(y != nullptr && out != nullptr && mean != nullptr && rstd != nullptr, "bn_act_fwd: null pointer argument");
  StreamGeom sg;
  if (int rc = stream_geom(M, Co, groups, act, 2, residual != nullptr, &sg)) return rc;
  const double inv_n = 1.0 / (double)sg.Mg;
  cudaStream_t st = (cudaStream_t)stream;
  EKL_ACT_SWITCH(act, EKL_BN_LAUNCH(bn_act_fwd_kernel, sg, st, (const bf16*)y, sg.Mg, sg.per, sg.CT, Cy, sums, inv_n, eps, momentum,
                                    mean, rstd, running_mean, running_var, gamma, beta, (const bf16*)residual, (bf16*)out,
                                    bn_rev() & 1, sg.coef_bytes));
  EKL_LAUNCH_CHECK();
  return 0;
}

// doubles of zeroed scratch ekl_bn_act_bwd needs for its two reductions (0: the layer runs the single-launch kernel)
extern "C" int64_t ekl_bn_bwd_scratch_doubles(int64_t M, int Cy, int groups, int act) {
  if (groups <= 0 || Cy <= 0 || M <= 0 || M % groups != 0) return -1;
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  if (bn_small(M, Co, groups, act)) return 0;
  return (int64_t)groups * 2 * Cy * BWD_SPREAD;
}

// sums: ekl_bn_bwd_scratch_doubles(...) doubles of caller scratch, ZERO on entry (unused by single-launch small layers);
// dgamma/dbeta accumulated (+=).
extern "C" int ekl_bn_act_bwd(const void* y, const void* dout, int64_t M, int Cy, int groups, const float* mean,
                              const float* rstd, const float* gamma, const float* beta, int act, double* sums,
                              float* dgamma, float* dbeta, void* dy, void* stream) {
  const int Co = act == ACT_GLU ? Cy / 2 : Cy;
  EKL_REQUIRE(groups > 0 && Co % 8 == 0 && M % groups == 0, "bn_act_bwd: bad shape Cy=%d", Cy);
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_small(M, Co, groups, act)) {
    dim3 sg(Co / act_vec(act) / SM_CT, groups);
    EKL_ACT_SWITCH(act, (bn_act_bwd_small_kernel<A><<<sg, 256, 0, st>>>((const bf16*)y, (const bf16*)dout, M / groups, Cy, mean, rstd,
                                                                      gamma, beta, dgamma, dbeta, (bf16*)dy)));
    EKL_LAUNCH_CHECK();
    return 0;
  }
  EKL_REQUIRE(sums != nullptr, "bn_act_bwd: sums scratch required");
  StreamGeom sg;
  if (int rc = stream_geom(M, Co, groups, act, 4, true, &sg)) return rc;
  const float inv_n = 1.f / (float)sg.Mg;
  EKL_ACT_SWITCH(act, EKL_BN_LAUNCH(bn_act_bwd_reduce_kernel, sg, st, (const bf16*)y, (const bf16*)dout, sg.Mg, sg.per, sg.CT, Cy,
                                    mean, rstd, gamma, beta, sums, bn_rev() & 2, sg.coef_bytes));
  EKL_LAUNCH_CHECK();
  EKL_ACT_SWITCH(act, EKL_BN_LAUNCH(bn_act_bwd_apply_kernel, sg, st, (const bf16*)y, (const bf16*)dout, sg.Mg, sg.per, sg.CT, Cy,
                                    inv_n, mean, rstd, gamma, beta, sums, dgamma, dbeta, (bf16*)dy, bn_rev() & 4, sg.coef_bytes));
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
