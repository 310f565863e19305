// tcgen05 gather-GEMM: the tensor-core kernel behind every convolution forward / data-gradient of the step.
//
//   out_v[p, n] = sum_t sum_k A_{map(v,t)}[p + (dh,dw), k] * Wp[v][n][t][k]            (conv_plan.h)
//
// One CTA computes a 128-pixel x BN-channel tile.  Roles (192 threads):
//   warp 0      TMA producer: per (tap, 64-channel block) one 4-D tiled box of A (the shifted pixel patch; the
//               conv zero padding is TMA out-of-bounds fill) + one 2-D box of packed weights, SWIZZLE_128B.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction),
//               accumulating fp32 in TMEM; tcgen05.commit releases smem stages / signals the epilogue.
//   warps 2..5  epilogue: tcgen05.ld -> bf16 -> swizzled smem staging -> per-channel sum / sum-of-squares
//               partials for train-mode BatchNorm -> TMA tensor store.
// Roofline: tensor pipe (2*128*BN*K flop per tile); operands are re-used from L2 across taps.
#include "conv_plan.h"
#include "ekl_common.cuh"

namespace {

struct TcParams {
  CUtensorMap a_maps[4];
  CUtensorMap o_maps[EKL_MAX_VAR];
  CUtensorMap w_map;
  EklTap taps[EKL_MAX_VAR][EKL_MAX_TAPS];
  float* stats;       // [(mtile*nvar + v)][2][N] or null
  int ntaps, ncb;     // taps, Cin/KC
  int Cin, N;
  int tb, th, tw, nTh, nTw;
  int rows_valid;     // tb*th*tw
};

template <int BN, int KC>
struct TcCfg {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = 128 * BN * 2;
  // two CTAs per SM: <= ~100 KB of stages each
  static constexpr int STAGES_RAW = (96 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_MAIN = PIPE_BYTES > OUT_BYTES ? PIPE_BYTES : OUT_BYTES;
  static constexpr int SMEM_BYTES = SMEM_MAIN + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);   // SW128 / SW64 / SW32
  static constexpr uint32_t SBO = 8 * KC * 2;
  static constexpr int OBOX = BN < 64 ? BN : 64;   // channels per output TMA box
};

template <int BN, int KC>
__global__ void __launch_bounds__(192) conv_gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using C = TcCfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + C::SMEM_MAIN);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.z;
  const int n0 = blockIdx.y * BN;
  int mt = blockIdx.x;
  const int twi = mt % p.nTw; mt /= p.nTw;
  const int thi = mt % p.nTh; mt /= p.nTh;
  const int w0 = twi * p.tw, h0 = thi * p.th, b0 = mt * p.tb;
  const int n_iters = p.ntaps * p.ncb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&p.w_map);
      const uint32_t tx = (uint32_t)(p.rows_valid * KC * 2 + C::B_BYTES);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int t = it / p.ncb, cb = it - t * p.ncb;
        const EklTap tap = p.taps[v][t];
        uint8_t* sa = smem + s * C::STAGE_BYTES;
        mbar_expect_tx(&full[s], tx);
        tma_load_4d(&p.a_maps[tap.map], &full[s], sa, cb * KC, w0 + tap.dw, h0 + tap.dh, b0);
        tma_load_2d(&p.w_map, &full[s], sa + C::A_BYTES, t * p.Cin + cb * KC, v * p.N + n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % C::STAGES;
      const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          const uint64_t da = umma_desc(sa + k * 32, 16, C::SBO, C::LAYOUT);
          const uint64_t db = umma_desc(sb + k * 32, 16, C::SBO, C::LAYOUT);
          tc_mma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty[s]);                         // smem stage reusable once these MMAs retire
        if (it == n_iters - 1) tc_commit(tmem_full);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ---------------- epilogue (warps 2..5); TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;   // 0..127
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    uint8_t* stage = smem;             // pipeline smem is idle now (all TMA loads consumed, all MMAs retired)
    if constexpr (BN == 16) {
      // one box of [128][16] bf16, 32-byte rows, no swizzle (a warp's 32 rows are 1 KB contiguous: conflict-free)
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      uint8_t* box = stage + row * 32;
      *reinterpret_cast<uint4*>(box) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(box + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      if constexpr (BN >= 64) {
        // staging = BN/64 boxes of [128 rows][64 ch] bf16, 128-byte rows, SWIZZLE_128B (16-B chunk index ^= row & 7)
        uint8_t* box = stage + (c0 >> 6) * (128 * 128) + row * 128;
        const int ch0 = (c0 & 63) >> 3;   // first 16-B chunk of this 32-column group: 0 or 4
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 val = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          *reinterpret_cast<uint4*>(box + (((ch0 + j) ^ (row & 7)) << 4)) = val;
        }
      } else {
        // BN == 32: one box of [128][32] bf16, 64-byte rows, SWIZZLE_64B (chunk index ^= (row >> 1) & 3)
        uint8_t* box = stage + row * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 val = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          *reinterpret_cast<uint4*>(box + ((j ^ ((row >> 1) & 3)) << 4)) = val;
        }
      }
    }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (et == 0) {
#pragma unroll
      for (int j = 0; j < BN / C::OBOX; ++j)
        tma_store_4d(&p.o_maps[v], stage + j * (128 * 128), n0 + j * C::OBOX, w0, h0, b0);
      tma_store_commit();
    }
    if (p.stats != nullptr) {
      // per-channel partial sums over this tile's valid rows, from the bf16 values actually stored
      float* dst = p.stats + ((size_t)blockIdx.x * gridDim.z + v) * 2 * p.N + n0;
      for (int c = et; c < BN; c += 128) {
        float s1 = 0.f, s2 = 0.f;
        const uint8_t* base;
        int chunk, sub = (c & 7) * 2;
        if constexpr (BN >= 64) { base = stage + (c >> 6) * (128 * 128); chunk = (c & 63) >> 3; }
        else { base = stage; chunk = c >> 3; }
        for (int rr = 0; rr < p.rows_valid; ++rr) {
          const uint8_t* ptr;
          if constexpr (BN >= 64) ptr = base + rr * 128 + ((chunk ^ (rr & 7)) << 4) + sub;
          else if constexpr (BN == 32) ptr = base + rr * 64 + ((chunk ^ ((rr >> 1) & 3)) << 4) + sub;
          else ptr = base + rr * 32 + (chunk << 4) + sub;
          const float x = __bfloat162float(*reinterpret_cast<const bf16*>(ptr));
          s1 += x; s2 += x * x;
        }
        dst[c] = s1;
        dst[p.N + c] = s2;
      }
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

int make_view_map(CUtensorMap* m, const EklView& v, int boxC, int tw, int th, int tb, int swz) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
  uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
  uint32_t box[4] = {(uint32_t)boxC, (uint32_t)tw, (uint32_t)th, (uint32_t)tb};
  return ekl_make_tmap(m, v.base, 4, dims, strides, box, swz, 2);
}

template <int BN, int KC>
int launch_tc(const EklGather* g, TcParams& p, int mtiles, cudaStream_t st) {
  using C = TcCfg<BN, KC>;
  auto kern = conv_gemm_tc_kernel<BN, KC>;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(mtiles, g->N / BN, g->nvar);
  kern<<<grid, 192, C::SMEM_BYTES, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

static int floor_pow2(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

// Tile geometry over the M grid: 128 rows = tb x th x tw.  `group_b`: batch extent that a tile may not straddle
// (per-group BatchNorm statistics), 0 = whole batch.
void ekl_tc_geometry(const EklGather* g, int group_b, int* tb, int* th, int* tw) {
  *tw = g->mW < 128 ? g->mW : 128;
  int rest = 128 / *tw;
  *th = g->mH < rest ? g->mH : rest;
  rest /= *th;
  int gb = group_b > 0 ? group_b : g->mB;
  int b = floor_pow2(gb < rest ? gb : rest);
  while (b > 1 && gb % b != 0) b /= 2;
  *tb = b;
}

int ekl_tc_supported(const EklGather* g) {
  auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
  if (g->Cin % 16 != 0 || g->N % 16 != 0) return 0;
  if (!pow2(g->mW) || !pow2(g->mH)) return 0;
  for (int i = 0; i < g->n_a; ++i)
    if (g->a[i].f32 || g->a[i].sC != 1) return 0;
  for (int i = 0; i < g->nvar; ++i)
    if (g->o[i].f32 || g->o[i].sC != 1) return 0;
  return 1;
}

// stats: [(mtile*nvar+v)][2][N] fp32 partials or null. Returns the number of M tiles through *mtiles_out.
int ekl_gather_gemm_tc(const EklGather* g, const void* w_packed, float* stats, int group_b, int* mtiles_out,
                       cudaStream_t st) {
  EKL_REQUIRE(ekl_tc_supported(g), "gather_gemm_tc: unsupported shape Cin=%d N=%d mH=%d mW=%d", g->Cin, g->N, g->mH, g->mW);
  TcParams p;
  memset(&p, 0, sizeof(p));
  int tb, th, tw;
  ekl_tc_geometry(g, group_b, &tb, &th, &tw);
  p.tb = tb; p.th = th; p.tw = tw;
  p.nTw = ekl_cdiv(g->mW, tw); p.nTh = ekl_cdiv(g->mH, th);
  const int nTb = ekl_cdiv(g->mB, tb);
  const int mtiles = p.nTw * p.nTh * nTb;
  if (mtiles_out) *mtiles_out = mtiles;
  p.rows_valid = tb * th * tw;
  p.ntaps = g->ntaps; p.Cin = g->Cin; p.N = g->N; p.stats = stats;
  memcpy(p.taps, g->taps, sizeof(p.taps));
  const int KC = g->Cin % 64 == 0 ? 64 : (g->Cin % 32 == 0 ? 32 : 16);
  p.ncb = g->Cin / KC;
  const int swz = KC == 64 ? 3 : (KC == 32 ? 2 : 1);
  int BN = g->N % 256 == 0 ? 256 : (g->N % 128 == 0 ? 128 : (g->N % 64 == 0 ? 64 : (g->N % 32 == 0 ? 32 : 16)));
  // prefer more CTAs when the grid would not fill the machine
  while (BN > 64 && (int64_t)mtiles * (g->N / BN) * g->nvar < 148) BN /= 2;
  for (int i = 0; i < g->n_a; ++i) {
    int rc = make_view_map(&p.a_maps[i], g->a[i], KC, tw, th, tb, swz);
    if (rc) return rc;
  }
  const int obox = BN < 64 ? BN : 64;
  for (int i = 0; i < g->nvar; ++i) {
    int rc = make_view_map(&p.o_maps[i], g->o[i], obox, tw, th, tb, BN >= 64 ? 3 : (BN == 32 ? 2 : 0));
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g->ntaps * g->Cin, (uint64_t)g->nvar * g->N};
    uint64_t strides[1] = {(uint64_t)g->ntaps * g->Cin * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    int rc = ekl_make_tmap(&p.w_map, w_packed, 2, dims, strides, box, swz, 2);
    if (rc) return rc;
  }
#define EKL_TC_CASE(bn, kc) if (BN == bn && KC == kc) return launch_tc<bn, kc>(g, p, mtiles, st);
  EKL_TC_CASE(256, 64) EKL_TC_CASE(128, 64) EKL_TC_CASE(64, 64) EKL_TC_CASE(32, 64) EKL_TC_CASE(16, 64)
  EKL_TC_CASE(256, 32) EKL_TC_CASE(128, 32) EKL_TC_CASE(64, 32) EKL_TC_CASE(32, 32) EKL_TC_CASE(16, 32)
  EKL_TC_CASE(256, 16) EKL_TC_CASE(128, 16) EKL_TC_CASE(64, 16) EKL_TC_CASE(32, 16) EKL_TC_CASE(16, 16)
#undef EKL_TC_CASE
  return ekl_fail(-1, "gather_gemm_tc: no kernel for BN=%d KC=%d", BN, KC);
}
