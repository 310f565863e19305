// tcgen05 convolution with the filter RESIDENT in shared memory and the input fetched once per tile as a row-halo patch:
// the small-contraction layers (Cin <= 128) whose taps all lie in the 3x3 neighbourhood of the output pixel --
//   * 3x3 / stride-1 forward and data-gradient, Cin in {16, 32, 64, 128} (model.py:107-123 ResBlock, :426-437 GET_IMAGE_G,
//     the folded jointConv, the discriminator stem);
//   * the sub-pixel plans with 4 output-parity variants x 2x2 taps: nearest-up-x2 + 3x3 forward (model.py:87-94) and the
//     data-gradient of conv4x4/s2 (model.py:816-830), Cin in {16, 32, 64, 128}.
//
// Why: the generic gather-GEMM kernel (conv_tc.cu) re-fetches the 128-pixel A tile and the filter tile for every tap
// (and variant); with a short contraction that makes it L2->SMEM bound at a fraction of the tensor pipe (config 2:
// dgrad of the 64->128 discriminator conv 173 us = 447 TFLOP/s, stage-3 up-conv 111 us = 102 TFLOP/s executed).  Here a
// CTA tile is 16 image rows x 8 columns of one image of the M grid:
//   * A per tile and 64-channel block: either three column-shifted row-halo TMA boxes [KC][8 w][18 h] (an 8-pixel tile
//     row is one 8-row swizzle atom, so tap (dh, dw) is box dw advanced by (dh+1) atoms), or ONE [KC][10 w][18 h] halo box
//     whose taps are pixel-row-shifted UMMA descriptors (start (dh+1)*10 + (dw+1) rows in, stride between 8-row groups
//     = 10 rows; the swizzle is a function of the shared-memory address, verified bit-correct for SW32/64/128) --
//     432 resp. 180 pixel rows instead of taps x 128, out-of-bounds = the conv zero padding;
//   * the ntaps x nkb filter blocks [BN][KC] of the current (variant, output-channel tile) stay in shared memory for all
//     tiles of a round; rounds walk (variant, N tile).
// Pipeline roles / TMEM accumulators in flight / epilogue (bf16 pack, BatchNorm sums, TMA store, optional fused
// activation and border-class bias) follow conv_tc.cu.
#include <stdlib.h>

#include "conv_plan.h"
#include "ekl_common.cuh"

int ekl_num_sms();

namespace {

struct RwParams {
  CUtensorMap a_map, w_map;
  CUtensorMap o_maps[EKL_MAX_VAR];
  EklTap taps[EKL_MAX_VAR][9];
  double* stats;         // [2][N] per-channel sum / sum-of-squares, accumulated with fp64 red.global.add; or null
  const float* bias9;    // [B][9][N] or null
  int N, Cin, ntn, H, W, nTh, nTw, tiles, act;
  int nvar, ntaps, nkb;  // output variants (1 | 4 parity classes), taps per variant (9 | 4), 64-channel blocks of Cin
  int halo1;             // ONE [KC][10 w][18 h] halo box per (tile, channel block); taps = row-shifted descriptors
  int stages;            // A pipeline depth
  int box_bytes;         // bytes of one A box (atom aligned); a stage holds nkb * (halo1 ? 1 : 3) of them
};

template <int BN, int KC>
struct RwCfg {
  static constexpr int ROWB = KC * 2;                                    // bytes per pixel row
  static constexpr int BOX3_BYTES = ((144 * ROWB + 1023) / 1024) * 1024;  // 18 x 8 pixel rows, atom aligned
  static constexpr int BOX1_BYTES = ((180 * ROWB + 1023) / 1024) * 1024;  // 18 x 10 pixel rows
  static constexpr int WT_BYTES = ((BN * ROWB + 1023) / 1024) * 1024;     // one (tap, channel block) filter tile
  static constexpr int OUT_BYTES = 128 * BN * 2;
  static constexpr int TAIL_BYTES = 512 + 2 * 2 * BN * 4 * (256 / BN);    // barriers + cross-slice reduction scratch
  static constexpr int BUDGET = 214 * 1024;
  static constexpr int NBUF = 4;                                          // TMEM accumulators in flight (4 x BN <= 256 columns)
  static constexpr int TMEM_COLS = NBUF * BN < 32 ? 32 : NBUF * BN;
  static constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);
  static constexpr uint32_t SBO = 8 * ROWB;
  static constexpr int QUADS = BN / 4;
  static constexpr int SLICES = 128 / QUADS;
  static constexpr int MAX_STAGES = 4;
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

// NKB: 64-channel blocks of the contraction, NTAPS: taps per variant -- compile-time, so that the single MMA-issuing
// thread runs a fully unrolled instruction stream with constant descriptor offsets (its issue rate bounds the kernel)
template <int BN, int KC, int NKB, int NTAPS>
__global__ void __launch_bounds__(192, 1) conv3x3_rw_kernel(const __grid_constant__ RwParams p) {
  using C = RwCfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = NKB * (p.halo1 ? 1 : 3) * p.box_bytes;
  constexpr int wblocks = NTAPS * NKB;                       // resident filter blocks of a round
  uint8_t* wsm = smem + p.stages * stage_bytes;
  uint8_t* stage_out = wsm + wblocks * C::WT_BYTES;
  uint64_t* full = (uint64_t*)(stage_out + C::OUT_BYTES);
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tmem_full = empty + C::MAX_STAGES;  // [NBUF]
  uint64_t* tmem_empty = tmem_full + C::NBUF;   // [NBUF]
  uint64_t* wfull = tmem_empty + C::NBUF;
  uint64_t* wfree = wfull + 1;
  uint32_t* tmem_slot = (uint32_t*)(wfree + 1);
  float* red = (float*)(stage_out + C::OUT_BYTES + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int rounds = p.nvar * p.ntn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
      for (int r = 0; r < rounds; ++r) {
        const int v = r / p.ntn, n = r - v * p.ntn;
        // the filter blocks may be overwritten once every MMA of the previous round has retired
        if (r > 0) mbar_wait(wfree, (uint32_t)(r - 1) & 1u);
        mbar_expect_tx(wfull, (uint32_t)(wblocks * BN * C::ROWB));
#pragma unroll
        for (int t = 0; t < NTAPS; ++t)
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb)
            tma_load_2d(&p.w_map, wfull, wsm + (t * NKB + kb) * C::WT_BYTES, t * p.Cin + kb * KC, v * p.N + n * BN);
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const int s = kit % p.stages;
          const uint32_t ph = (kit / p.stages) & 1u;
          int w0, h0, b0;
          tile_origin(tile, w0, h0, b0);
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sa = smem + s * stage_bytes;
          if (p.halo1) {
            mbar_expect_tx(&full[s], (uint32_t)(NKB * 180 * C::ROWB));
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) tma_load_4d(&p.a_map, &full[s], sa + kb * p.box_bytes, kb * KC, w0 - 1, h0 - 1, b0);
          } else {
            mbar_expect_tx(&full[s], (uint32_t)(NKB * 3 * 144 * C::ROWB));
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                tma_load_4d(&p.a_map, &full[s], sa + (kb * 3 + j) * p.box_bytes, kb * KC, w0 + j - 1, h0 - 1, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (has_tiles) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      uint32_t kit = 0;
      // descriptors are affine in the shared address: desc(base + off) = desc(base) + (off >> 4) (no carry out of the
      // 14-bit start field for offsets inside one CTA's shared memory).  Per-tap A offsets are tile-invariant.
      const uint32_t a_sbo = p.halo1 ? 10u * C::ROWB : C::SBO;
      const uint32_t kb_stride = (uint32_t)((p.halo1 ? 1 : 3) * p.box_bytes) >> 4;
      const uint64_t bdesc0 = umma_desc(smem_u32(wsm), 16, C::SBO, C::LAYOUT);
      for (int r = 0; r < rounds; ++r) {
        const int v = r / p.ntn;
        uint32_t a_off[NTAPS];
#pragma unroll
        for (int t = 0; t < NTAPS; ++t) {
          const EklTap tap = p.taps[v][t];
          a_off[t] = (p.halo1 ? (uint32_t)((tap.dh + 1) * 10 + tap.dw + 1) * C::ROWB
                              : (uint32_t)(tap.dw + 1) * p.box_bytes + (uint32_t)(tap.dh + 1) * C::SBO) >> 4;
        }
        mbar_wait(wfull, (uint32_t)r & 1u);
        tc_fence_after();
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const int s = kit % p.stages;
          const uint32_t ph = (kit / p.stages) & 1u;
          const uint32_t buf = kit % C::NBUF, use = kit / C::NBUF;
          mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc0 = umma_desc(smem_u32(smem + s * stage_bytes), 16, a_sbo, C::LAYOUT);
            const uint32_t tacc = tmem_base + buf * BN;
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
#pragma unroll
              for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  tc_mma_bf16(tacc, adesc0 + (uint64_t)(a_off[t] + kb * kb_stride + 2 * k),
                              bdesc0 + (uint64_t)((((t * NKB + kb) * C::WT_BYTES) + k * 32) >> 4), idesc,
                              (t | kb | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(&empty[s]);
            tc_commit(&tmem_full[buf]);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(wfree);       // all MMAs of this round retired -> filter blocks reusable
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int quad = et % C::QUADS, slice = et / C::QUADS;
    const uint32_t so = smem_u32(stage_out);
    uint32_t kit = 0;
    bool store_pending = false;
    for (int r = 0; r < rounds; ++r) {
      const int v = r / p.ntn, n = r - v * p.ntn;
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      if (has_tiles)
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const uint32_t buf = kit % C::NBUF, use = kit / C::NBUF;
          int w0, h0, b0;
          tile_origin(tile, w0, h0, b0);
          mbar_wait(&tmem_full[buf], use & 1u);
          tc_fence_after();
          if (store_pending && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
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
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_4d(&p.o_maps[v], stage_out, n * BN, w0, h0, b0);
            tma_store_commit();
          }
          store_pending = true;
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
        }
      if (p.stats != nullptr && has_tiles) {
        // all variants of an up-conv land in the same per-channel sums
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// shared-memory plan of a launch: A boxes per stage, pipeline depth (0: the resident filter does not leave room for two)
template <int BN, int KC>
int rw_smem_plan(int ntaps, int nkb, int halo1, int* stages, int* box_bytes, int* smem_bytes) {
  using C = RwCfg<BN, KC>;
  *box_bytes = halo1 ? C::BOX1_BYTES : C::BOX3_BYTES;
  const int stage = nkb * (halo1 ? 1 : 3) * *box_bytes;
  const int fixed = ntaps * nkb * C::WT_BYTES + C::OUT_BYTES + C::TAIL_BYTES + 1024;
  int st = (C::BUDGET - fixed) / stage;
  if (st > C::MAX_STAGES) st = C::MAX_STAGES;
  *stages = st;
  *smem_bytes = st * stage + fixed;
  return st >= 2;
}

template <int BN, int KC, int NKB, int NTAPS>
int launch_rw(RwParams& p, int grid, cudaStream_t st) {
  int smem_bytes = 0;
  EKL_REQUIRE((rw_smem_plan<BN, KC>(NTAPS, NKB, p.halo1, &p.stages, &p.box_bytes, &smem_bytes)),
              "conv3x3_rw: the resident filter leaves no room for the input pipeline");
  auto kern = conv3x3_rw_kernel<BN, KC, NKB, NTAPS>;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RwCfg<BN, KC>::BUDGET + 8 * 1024));
    attr_done = true;
  }
  kern<<<grid, 192, smem_bytes, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

int rw_kc(const EklGather* g) { return g->Cin % 64 == 0 ? 64 : g->Cin; }

// input-pipeline stages that fit next to the resident filter (same arithmetic as rw_smem_plan)
int rw_fit(int ntaps, int nkb, int KC, int BN, int halo1) {
  const int rowb = KC * 2;
  const int wt = (BN * rowb + 1023) / 1024 * 1024;
  const int box = ((halo1 ? 180 : 144) * rowb + 1023) / 1024 * 1024;
  const int stage = nkb * (halo1 ? 1 : 3) * box;
  const int fixed = ntaps * nkb * wt + 128 * BN * 2 + 512 + 4096 + 1024;
  return (214 * 1024 - fixed) / stage;
}

int rw_halo1(const EklGather* g);

// N-tile width: the widest of 64 / 32 / 16 that divides N and leaves room for two input stages (0: none does).  A 3x3 conv
// over 128 channels keeps 18 filter blocks resident: 32-wide tiles there (generator ResBlock data-gradients).
int rw_bn(const EklGather* g) {
  const int KC = rw_kc(g), nkb = g->Cin / KC, halo1 = rw_halo1(g);
  for (int bn = 64; bn >= 16; bn >>= 1)
    if (g->N % bn == 0 && rw_fit(g->ntaps, nkb, KC, bn, halo1) >= 2) return bn;
  return 0;
}

// single-halo-box mode: forced for the plans that only exist in that form (several channel blocks, parity variants),
// opt-in for the classic 3x3 layers (EKL_RW_HALO=1; measured a wash there: those are not load bound)
int rw_halo1(const EklGather* g) {
  if (g->nvar > 1 || g->Cin > 64) return 1;
  static int v = -1;
  if (v < 0) { const char* e = getenv("EKL_RW_HALO"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

// EKL_RW_SUBPIXEL=0 keeps the 4-variant / 128-channel plans on the generic kernel (A/B measurements)
bool rw_subpixel_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EKL_RW_SUBPIXEL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

}  // namespace

// The resident-filter kernel applies to: one A view, one statistics group, every tap inside the 3x3 neighbourhood,
// (1 variant x 9 taps) or (4 parity variants x 4 taps), Cin in {16, 32, 64, 128}, N % 16 == 0, H % 16 == 0, W % 8 == 0,
// and a filter (taps x Cin x N-tile) that leaves room for a two-stage input pipeline (rw_bn picks the N-tile width).
int ekl_rw_supported(const EklGather* g, int group_b) {
  if (g->n_a != 1) return 0;
  const bool classic = g->nvar == 1 && g->ntaps == 9, sub = g->nvar == 4 && g->ntaps == 4;
  if (!classic && !sub) return 0;
  if (sub && !rw_subpixel_enabled()) return 0;
  if (!(g->Cin == 16 || g->Cin == 32 || g->Cin == 64 || g->Cin == 128)) return 0;
  if (g->N % 16 != 0 || g->mH % 16 != 0 || g->mW % 8 != 0) return 0;
  if (group_b > 0 && group_b != g->mB) return 0;
  if (g->a[0].f32 || g->a[0].sC != 1) return 0;
  for (int v = 0; v < g->nvar; ++v) {
    if (g->o[v].f32 || g->o[v].sC != 1) return 0;
    for (int t = 0; t < g->ntaps; ++t)
      if (g->taps[v][t].map != 0 || g->taps[v][t].dh < -1 || g->taps[v][t].dh > 1 || g->taps[v][t].dw < -1 || g->taps[v][t].dw > 1)
        return 0;
  }
  // sub-pixel plans reload the resident filter once per (variant, N tile) round: worth it only when a round has a few
  // tiles per CTA (the generator's 32x32 -> 64x64 up-conv, 192 tiles, stays on the CTA-pair kernel)
  if (sub && (int64_t)g->mB * (g->mH / 16) * (g->mW / 8) < 2 * (int64_t)ekl_num_sms()) return 0;
  // shared-memory plan: filter blocks + staging + >= 2 input stages for some N-tile width
  if (rw_bn(g) == 0) return 0;
  // a 3x3 conv over 128 channels runs 32-wide N tiles and re-reads the input once per tile: only for one or two of them
  if (classic && g->Cin == 128 && g->N / rw_bn(g) > 2) return 0;
  return 1;
}

int ekl_conv3x3_rw(const EklGather* g, const void* w_packed, double* stats, int act, const float* bias9, cudaStream_t st) {
  EKL_REQUIRE(ekl_rw_supported(g, 0), "conv3x3_rw: unsupported plan");
  EKL_REQUIRE(g->nvar == 1 || (act == 0 && bias9 == nullptr), "conv3x3_rw: no fused activation / bias on sub-pixel plans");
  RwParams p;
  memset(&p, 0, sizeof(p));
  for (int v = 0; v < g->nvar; ++v) memcpy(p.taps[v], g->taps[v], sizeof(EklTap) * g->ntaps);
  p.stats = stats; p.bias9 = bias9; p.N = g->N; p.Cin = g->Cin; p.H = g->mH; p.W = g->mW; p.act = act;
  p.nvar = g->nvar; p.ntaps = g->ntaps;
  p.nTh = g->mH / 16; p.nTw = g->mW / 8; p.tiles = g->mB * p.nTh * p.nTw;
  const int KC = rw_kc(g);
  const int BN = rw_bn(g);
  p.nkb = g->Cin / KC;
  p.ntn = g->N / BN;
  p.halo1 = rw_halo1(g);
  const int swz = KC == 64 ? 3 : (KC == 32 ? 2 : 1);
  {
    const EklView& v = g->a[0];
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
    uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)(p.halo1 ? 10 : 8), 18, 1};
    if (int rc = ekl_make_tmap(&p.a_map, v.base, 4, dims, strides, box, swz, 2)) return rc;
  }
  for (int i = 0; i < g->nvar; ++i) {
    const EklView& v = g->o[i];
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
    uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
    uint32_t box[4] = {(uint32_t)BN, 8, 16, 1};
    if (int rc = ekl_make_tmap(&p.o_maps[i], v.base, 4, dims, strides, box, BN == 64 ? 3 : (BN == 32 ? 2 : 0), 2)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g->ntaps * g->Cin, (uint64_t)g->nvar * g->N};
    uint64_t strides[1] = {(uint64_t)g->ntaps * g->Cin * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (int rc = ekl_make_tmap(&p.w_map, w_packed, 2, dims, strides, box, swz, 2)) return rc;
  }
  const int grid = ekl_num_sms();      // persistent: one CTA per SM
#define EKL_RW_CASE(bn, kc, NKB_, NT_) \
  if (BN == bn && KC == kc && p.nkb == NKB_ && p.ntaps == NT_) return launch_rw<bn, kc, NKB_, NT_>(p, grid, st);
#define EKL_RW_BN(kc, NKB_, NT_) EKL_RW_CASE(64, kc, NKB_, NT_) EKL_RW_CASE(32, kc, NKB_, NT_) EKL_RW_CASE(16, kc, NKB_, NT_)
  EKL_RW_BN(64, 1, 9) EKL_RW_BN(32, 1, 9) EKL_RW_BN(16, 1, 9)
  EKL_RW_BN(64, 1, 4) EKL_RW_BN(32, 1, 4) EKL_RW_BN(16, 1, 4) EKL_RW_BN(64, 2, 4) EKL_RW_BN(64, 2, 9)
#undef EKL_RW_BN
#undef EKL_RW_CASE
  return ekl_fail(-1, "conv3x3_rw: no kernel for BN=%d KC=%d", BN, KC);
}
