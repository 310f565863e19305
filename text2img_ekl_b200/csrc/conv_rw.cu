// tcgen05 3x3 / stride-1 convolution for SMALL channel counts (Cin in {16, 32, 64}: one K block per tap) with the
// filter RESIDENT in shared memory and the input fetched once per tile as three column-shifted row-halo boxes.
//
// Why: the generic gather-GEMM kernel (conv_tc.cu) re-fetches the 128-pixel A tile and the filter tile for each of the
// 9 taps; for the generator's high-resolution layers (64..256^2 maps, 16..64 channels, model.py:107-123 ResBlock,
// :426-437 GET_IMAGE_G, the folded jointConv, the discriminator stem) that makes the kernel L2->SMEM bound at a
// fraction of the tensor pipe.  Here a CTA tile is 16 image rows x 8 columns of one image:
//   * per tile, 3 TMA boxes [Cin][8 w][18 h] at column offsets -1, 0, +1 (rows h0-1 .. h0+16; out-of-bounds = the conv
//     zero padding) -- 432 pixel rows instead of 9 x 128, and no filter traffic at all;
//   * because a tile row is exactly 8 pixels = one 8-row swizzle atom, the A operand of tap (dh, dw) is the box of
//     column shift dw advanced by (dh+1) atoms: an atom-aligned descriptor start, no partial-atom addressing;
//   * the 9 filter taps [BN][Cin] of the current output-channel tile stay in shared memory for all tiles of a round.
// Pipeline roles / TMEM double buffering / epilogue (bf16 pack, BatchNorm partial sums, TMA store, optional fused
// activation and border-class bias) follow conv_tc.cu.  Forward and data-gradient of EKL_S1 (taps differ only in sign).
#include <stdlib.h>

#include "conv_plan.h"
#include "ekl_common.cuh"

// -DEKL_RW_TIMING: per-role cycle accounting of CTA 0, printed at kernel exit (development aid)
#ifdef EKL_RW_TIMING
#define RW_T0() long long t0_ = clock64()
#define RW_ACC(x) do { const long long t1_ = clock64(); (x) += t1_ - t0_; t0_ = t1_; } while (0)
#else
#define RW_T0() do {} while (0)
#define RW_ACC(x) do {} while (0)
#endif

namespace {

struct RwParams {
  CUtensorMap a_map, o_map, w_map;
  EklTap taps[9];
  double* stats;         // [2][N] per-channel sum / sum-of-squares, accumulated with fp64 red.global.add; or null
  const float* bias9;    // [B][9][N] or null
  int N, ntn, H, W, nTh, nTw, tiles, act;
  int halo1;             // experiment: ONE [Cin][10 w][18 h] halo box per tile, taps = row-shifted descriptors
};

template <int BN, int KC>
struct RwCfg {
  static constexpr int ROWB = KC * 2;                                  // bytes per pixel row
  static constexpr int BOX_BYTES = ((144 * ROWB + 1023) / 1024) * 1024; // 18 x 8 pixel rows, atom aligned
  static constexpr int STAGE_BYTES = 3 * BOX_BYTES;
  static constexpr int WT_BYTES = ((BN * ROWB + 1023) / 1024) * 1024;   // one tap's filter tile
  static constexpr int W_BYTES = 9 * WT_BYTES;
  static constexpr int OUT_BYTES = 128 * BN * 2;
  static constexpr int STAGES_RAW = (212 * 1024 - W_BYTES - OUT_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static_assert(STAGES >= 2, "not enough shared memory for two stages");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + W_BYTES + OUT_BYTES + 1024 + 512 + 2 * 2 * BN * 4 * (256 / BN);
  static constexpr int NBUF = 4;                                        // TMEM accumulators in flight (4 x BN <= 256 columns)
  static constexpr int TMEM_COLS = NBUF * BN < 32 ? 32 : NBUF * BN;
  static constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);
  static constexpr uint32_t SBO = 8 * ROWB;
  static constexpr int QUADS = BN / 4;
  static constexpr int SLICES = 128 / QUADS;
};

template <int BN>
__device__ __forceinline__ uint32_t rw_stage_off(int r, int c) {
  if constexpr (BN == 64) return (uint32_t)(r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
  else if constexpr (BN == 32) return (uint32_t)(r * 64 + (((c >> 3) ^ ((r >> 1) & 3)) << 4) + (c & 7) * 2);
  else return (uint32_t)(r * 32 + c * 2);
}

__device__ __forceinline__ float rw_act(float v, int act) {
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  if (act == 4) return tanhf(v);
  return v;
}

template <int BN, int KC>
__global__ void __launch_bounds__(192, 1) conv3x3_rw_kernel(const __grid_constant__ RwParams p) {
  using C = RwCfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem + C::STAGES * C::STAGE_BYTES;
  uint8_t* stage_out = wsm + C::W_BYTES;
  uint64_t* full = (uint64_t*)(stage_out + C::OUT_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;      // [NBUF]
  uint64_t* tmem_empty = tmem_full + C::NBUF;   // [NBUF]
  uint64_t* wfull = tmem_empty + C::NBUF;
  uint64_t* wfree = wfull + 1;
  uint32_t* tmem_slot = (uint32_t*)(wfree + 1);
  float* red = (float*)(stage_out + C::OUT_BYTES + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < C::NBUF; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
    mbar_init(wfull, 1); mbar_init(wfree, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_tiles = cta < p.tiles;

  auto tile_origin = [&](int tile, int& w0, int& h0, int& b0) {
    const int twi = tile % p.nTw; tile /= p.nTw;
    const int thi = tile % p.nTh; tile /= p.nTh;
    w0 = twi * 8; h0 = thi * 16; b0 = tile;
  };

  if (warp == 0) {
    if (has_tiles && elect_one()) {
      tma_prefetch_desc(&p.w_map);
      tma_prefetch_desc(&p.a_map);
      uint32_t kit = 0;
#ifdef EKL_RW_TIMING
      long long tp_wait = 0, tp_issue = 0;
#endif
      RW_T0();
      for (int n = 0; n < p.ntn; ++n) {
        // the filter tile may be overwritten once every MMA of the previous round has retired
        if (n > 0) mbar_wait(wfree, (uint32_t)(n - 1) & 1u);
        mbar_expect_tx(wfull, (uint32_t)(9 * BN * C::ROWB));
        for (int t = 0; t < 9; ++t) tma_load_2d(&p.w_map, wfull, wsm + t * C::WT_BYTES, t * KC, n * BN);
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const int s = kit % C::STAGES;
          const uint32_t ph = (kit / C::STAGES) & 1u;
          int w0, h0, b0;
          tile_origin(tile, w0, h0, b0);
          mbar_wait(&empty[s], ph ^ 1u);
          RW_ACC(tp_wait);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          if (p.halo1) {
            mbar_expect_tx(&full[s], (uint32_t)(180 * C::ROWB));
            tma_load_4d(&p.a_map, &full[s], sa, 0, w0 - 1, h0 - 1, b0);
          } else {
            mbar_expect_tx(&full[s], (uint32_t)(3 * 144 * C::ROWB));
#pragma unroll
            for (int j = 0; j < 3; ++j) tma_load_4d(&p.a_map, &full[s], sa + j * C::BOX_BYTES, 0, w0 + j - 1, h0 - 1, b0);
          }
          RW_ACC(tp_issue);
        }
      }
#ifdef EKL_RW_TIMING
      if (cta == 0) printf("rw producer: wait_empty %lld issue %lld cycles, %u tiles\n", tp_wait, tp_issue, kit);
#endif
    }
  } else if (warp == 1) {
    if (has_tiles) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      uint32_t kit = 0;
      // descriptors are affine in the shared address: desc(base + off) = desc(base) + (off >> 4) (no carry out of the
      // 14-bit start field for offsets inside one CTA's shared memory).  Per-tap A offsets are tile-invariant.
      uint32_t a_off[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const EklTap tap = p.taps[t];
        a_off[t] = (p.halo1 ? (uint32_t)((tap.dh + 1) * 10 + tap.dw + 1) * C::ROWB
                            : (uint32_t)(tap.dw + 1) * C::BOX_BYTES + (uint32_t)(tap.dh + 1) * C::SBO) >> 4;
      }
      const uint32_t a_sbo = p.halo1 ? 10u * C::ROWB : C::SBO;
      const uint64_t bdesc0 = umma_desc(smem_u32(wsm), 16, C::SBO, C::LAYOUT);
#ifdef EKL_RW_TIMING
      long long tm_tmem = 0, tm_full = 0, tm_issue = 0, tm_w = 0;
#endif
      RW_T0();
      for (int n = 0; n < p.ntn; ++n) {
        mbar_wait(wfull, (uint32_t)n & 1u);
        tc_fence_after();
        RW_ACC(tm_w);
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const int s = kit % C::STAGES;
          const uint32_t ph = (kit / C::STAGES) & 1u;
          const uint32_t buf = kit % C::NBUF, use = kit / C::NBUF;
          mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);
          RW_ACC(tm_tmem);
          mbar_wait(&full[s], ph);
          tc_fence_after();
          RW_ACC(tm_full);
          if (elect_one()) {
            const uint64_t adesc0 = umma_desc(smem_u32(smem + s * C::STAGE_BYTES), 16, a_sbo, C::LAYOUT);
            const uint32_t tacc = tmem_base + buf * BN;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc_mma_bf16(tacc, adesc0 + (uint64_t)(a_off[t] + 2 * k), bdesc0 + (uint64_t)((t * C::WT_BYTES + k * 32) >> 4), idesc,
                            (t | k) != 0 ? 1u : 0u);
            }
            tc_commit(&empty[s]);
            tc_commit(&tmem_full[buf]);
          }
          __syncwarp();
          RW_ACC(tm_issue);
        }
        if (elect_one()) tc_commit(wfree);       // all MMAs of this round retired -> filter tile reusable
        __syncwarp();
      }
#ifdef EKL_RW_TIMING
      if (cta == 0 && lane == 0) printf("rw mma: wait_w %lld wait_tmem_empty %lld wait_full %lld issue %lld\n", tm_w, tm_tmem, tm_full, tm_issue);
#endif
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int quad = et % C::QUADS, slice = et / C::QUADS;
    const uint32_t so = smem_u32(stage_out);
    uint32_t kit = 0;
    bool store_pending = false;
#ifdef EKL_RW_TIMING
    long long te_full = 0, te_store = 0, te_ld = 0, te_bar = 0, te_stats = 0;
#endif
    RW_T0();
    for (int n = 0; n < p.ntn; ++n) {
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      if (has_tiles)
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const uint32_t buf = kit % C::NBUF, use = kit / C::NBUF;
          int w0, h0, b0;
          tile_origin(tile, w0, h0, b0);
          mbar_wait(&tmem_full[buf], use & 1u);
          tc_fence_after();
          RW_ACC(te_full);
          if (store_pending && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
          RW_ACC(te_store);
          const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
          const float* brow = nullptr;
          if (p.bias9 != nullptr) {
            const int hh = h0 + (row >> 3), ww = w0 + (row & 7);
            const int cls = (hh == 0 ? 0 : (hh == p.H - 1 ? 2 : 1)) * 3 + (ww == 0 ? 0 : (ww == p.W - 1 ? 2 : 1));
            brow = p.bias9 + ((size_t)b0 * 9 + cls) * p.N + n * BN;
          }
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += (BN < 32 ? 16 : 32)) {
            constexpr int CH = BN < 32 ? 16 : 32;
            uint32_t rr[CH];
            if constexpr (CH == 16) tmem_ld16(tacc + (uint32_t)c0, rr); else tmem_ld32(tacc + (uint32_t)c0, rr);
            tmem_ld_wait();
            if (brow != nullptr) {
#pragma unroll
              for (int i = 0; i < CH; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(brow + c0 + i);
                rr[i] = __float_as_uint(__uint_as_float(rr[i]) + bv.x); rr[i + 1] = __float_as_uint(__uint_as_float(rr[i + 1]) + bv.y);
                rr[i + 2] = __float_as_uint(__uint_as_float(rr[i + 2]) + bv.z); rr[i + 3] = __float_as_uint(__uint_as_float(rr[i + 3]) + bv.w);
              }
            }
            if (p.act != 0) {
#pragma unroll
              for (int i = 0; i < CH; ++i) rr[i] = __float_as_uint(rw_act(__uint_as_float(rr[i]), p.act));
            }
#pragma unroll
            for (int j = 0; j < CH / 8; ++j)
              sts128(so + rw_stage_off<BN>(row, c0 + 8 * j),
                     make_uint4(pack_bf16x2(__uint_as_float(rr[8 * j]), __uint_as_float(rr[8 * j + 1])),
                                pack_bf16x2(__uint_as_float(rr[8 * j + 2]), __uint_as_float(rr[8 * j + 3])),
                                pack_bf16x2(__uint_as_float(rr[8 * j + 4]), __uint_as_float(rr[8 * j + 5])),
                                pack_bf16x2(__uint_as_float(rr[8 * j + 6]), __uint_as_float(rr[8 * j + 7]))));
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);
          RW_ACC(te_ld);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_4d(&p.o_map, stage_out, n * BN, w0, h0, b0);
            tma_store_commit();
          }
          store_pending = true;
          RW_ACC(te_bar);
          if (p.stats != nullptr) {
            constexpr int RPS = 128 / C::SLICES;
            const int r0 = slice * RPS;
#pragma unroll
            for (int rr = r0; rr < r0 + RPS; ++rr) {
              const uint2 u = lds64(so + rw_stage_off<BN>(rr, 4 * quad));
              const float x0 = bf16_lo(u.x), x1 = bf16_hi(u.x), x2 = bf16_lo(u.y), x3 = bf16_hi(u.y);
              s1[0] += x0; s2[0] += x0 * x0; s1[1] += x1; s2[1] += x1 * x1;
              s1[2] += x2; s2[2] += x2 * x2; s1[3] += x3; s2[3] += x3 * x3;
            }
          }
          RW_ACC(te_stats);
        }
      if (p.stats != nullptr && has_tiles) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float* my = red + (size_t)slice * 2 * BN;
#pragma unroll
        for (int i = 0; i < 4; ++i) { my[4 * quad + i] = s1[i]; my[BN + 4 * quad + i] = s2[i]; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        double* dst = p.stats + n * BN;
        for (int c = et; c < 2 * BN; c += 128) {
          float acc = 0.f;
#pragma unroll
          for (int sl = 0; sl < C::SLICES; ++sl) acc += red[(size_t)sl * 2 * BN + c];
          atomicAdd(dst + ((c < BN) ? c : (p.N + c - BN)), (double)acc);
        }
      }
    }
    if (store_pending && et == 0) tma_store_wait_all();
#ifdef EKL_RW_TIMING
    if (cta == 0 && et == 0)
      printf("rw epilogue: wait_tmem_full %lld wait_store_read+bar %lld ld/pack/sts %lld fence+bar+store %lld stats %lld\n", te_full,
             te_store, te_ld, te_bar, te_stats);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN, int KC>
int launch_rw(RwParams& p, int grid, cudaStream_t st) {
  using C = RwCfg<BN, KC>;
  auto kern = conv3x3_rw_kernel<BN, KC>;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  kern<<<grid, 192, C::SMEM_BYTES, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int ekl_num_sms();

// The resident-filter kernel applies to: stride-1 3x3 plans, one statistics group, Cin in {16,32,64}, N % 16 == 0,
// H % 16 == 0, W % 8 == 0.
int ekl_rw_supported(const EklGather* g, int group_b) {
  if (g->nvar != 1 || g->ntaps != 9 || g->n_a != 1) return 0;
  if (!(g->Cin == 16 || g->Cin == 32 || g->Cin == 64)) return 0;
  if (g->N % 16 != 0 || g->mH % 16 != 0 || g->mW % 8 != 0) return 0;
  if (group_b > 0 && group_b != g->mB) return 0;
  if (g->a[0].f32 || g->a[0].sC != 1 || g->o[0].f32 || g->o[0].sC != 1) return 0;
  for (int t = 0; t < 9; ++t)
    if (g->taps[0][t].dh < -1 || g->taps[0][t].dh > 1 || g->taps[0][t].dw < -1 || g->taps[0][t].dw > 1) return 0;
  return 1;
}

int ekl_conv3x3_rw(const EklGather* g, const void* w_packed, double* stats, int act, const float* bias9, cudaStream_t st) {
  EKL_REQUIRE(ekl_rw_supported(g, 0), "conv3x3_rw: unsupported plan");
  RwParams p;
  memset(&p, 0, sizeof(p));
  memcpy(p.taps, g->taps[0], sizeof(p.taps));
  p.stats = stats; p.bias9 = bias9; p.N = g->N; p.H = g->mH; p.W = g->mW; p.act = act;
  p.nTh = g->mH / 16; p.nTw = g->mW / 8; p.tiles = g->mB * p.nTh * p.nTw;
  const int KC = g->Cin;
  const int BN = g->N % 64 == 0 ? 64 : (g->N % 32 == 0 ? 32 : 16);
  p.ntn = g->N / BN;
  const int swz = KC == 64 ? 3 : (KC == 32 ? 2 : 1);
  {
    const EklView& v = g->a[0];
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
    uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
    const char* e = getenv("EKL_RW_HALO");
    p.halo1 = (e && e[0] == '1') ? 1 : 0;
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)(p.halo1 ? 10 : 8), 18, 1};
    if (int rc = ekl_make_tmap(&p.a_map, v.base, 4, dims, strides, box, swz, 2)) return rc;
  }
  {
    const EklView& v = g->o[0];
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
    uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
    uint32_t box[4] = {(uint32_t)BN, 8, 16, 1};
    if (int rc = ekl_make_tmap(&p.o_map, v.base, 4, dims, strides, box, BN == 64 ? 3 : (BN == 32 ? 2 : 0), 2)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * g->Cin, (uint64_t)g->N};
    uint64_t strides[1] = {(uint64_t)9 * g->Cin * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (int rc = ekl_make_tmap(&p.w_map, w_packed, 2, dims, strides, box, swz, 2)) return rc;
  }
  const int grid = ekl_num_sms();      // persistent: one CTA per SM
#define EKL_RW_CASE(bn, kc) if (BN == bn && KC == kc) return launch_rw<bn, kc>(p, grid, st);
  EKL_RW_CASE(64, 64) EKL_RW_CASE(32, 64) EKL_RW_CASE(16, 64)
  EKL_RW_CASE(64, 32) EKL_RW_CASE(32, 32) EKL_RW_CASE(16, 32)
  EKL_RW_CASE(64, 16) EKL_RW_CASE(32, 16) EKL_RW_CASE(16, 16)
#undef EKL_RW_CASE
  return ekl_fail(-1, "conv3x3_rw: no kernel for BN=%d KC=%d", BN, KC);
}
