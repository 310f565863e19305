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
#include <stdlib.h>

#include "conv_plan.h"
#include "ekl_common.cuh"

namespace {

struct TcParams {
  CUtensorMap a_maps[4];
  CUtensorMap o_maps[EKL_MAX_VAR];
  CUtensorMap w_map;
  EklTap taps[EKL_MAX_VAR][EKL_MAX_TAPS];
  double* stats;      // [group][2][N] per-channel sum / sum-of-squares, accumulated with fp64 red.global.add; or null
  int ntaps, ncb;     // taps, Cin/KC
  int Cin, N;
  int tb, th, tw, nTh, nTw;
  int rows_valid;     // tb*th*tw
  int nvar, ntn;      // variants, N tiles
  int groups, mtg;    // BatchNorm statistic groups, M tiles per group
  int act;            // epilogue activation on the fp32 accumulator: 0 none, 2 LeakyReLU(0.2), 4 tanh
  const float* bias9; // [B][9][N] per-sample, per-border-class bias added to the accumulator (folded code channels) or null
  int H, W;           // M-grid extents (border classes of bias9)
  int ksplit;         // split-K: the (tap, channel-block) iterations of a tile are divided over ksplit work items ...
  float* scratch;     // ... whose fp32 partial tiles are red-added into scratch[pixel][N] (finished by splitk_finish)
  int mB;             // batch extent of the M grid (bounds of the scratch rows)
  int b_mn;           // data-gradient reading the FORWARD-packed filter [Cout][tap][Cin] as an MN-major B operand
  int Nm;             // b_mn: master Cin (row pitch of the forward-packed filter = ntaps_master * Nm)
};

template <int BN, int KC>
struct TcCfg {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = 128 * BN * 2;
  // one persistent CTA per SM: pipeline stages + a dedicated epilogue staging tile within ~220 KB
  static constexpr int STAGES_RAW = (220 * 1024 - OUT_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = PIPE_BYTES + OUT_BYTES + 1024 /*align slack*/ + 512 /*barriers + reduction scratch*/ +
                                    4096 /* SLICES x 2 x BN floats of cross-slice reduction scratch */;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;               // double-buffered accumulator
  static constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);   // SW128 / SW64 / SW32
  static constexpr uint32_t SBO = 8 * KC * 2;
  static constexpr int OBOX = BN < 64 ? BN : 64;   // channels per output TMA box
  // statistics: each epilogue thread owns one channel QUAD (8-byte shared loads) over a slice of the rows
  static constexpr int QUADS = BN / 4;
  static constexpr int SLICES = 128 / QUADS;
};

// byte offset of channel c (even) of row r inside the swizzled staging tile
template <int BN>
__device__ __forceinline__ uint32_t stage_off(int r, int c) {
  if constexpr (BN >= 64) return (uint32_t)((c >> 6) * (128 * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
  else if constexpr (BN == 32) return (uint32_t)(r * 64 + (((c >> 3) ^ ((r >> 1) & 3)) << 4) + (c & 7) * 2);
  else return (uint32_t)(r * 32 + c * 2);
}

__device__ __forceinline__ float epi_act(float v, int act) {
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  if (act == 4) return tanhf(v);
  return v;
}

// Persistent kernel: grid = #SMs; every CTA walks the same static schedule
//   for round (v, n, g):  M tiles of group g with index (cta + rotation) + i*grid
// so that one CTA's tiles of a round share the output-channel range (their BatchNorm partial sums accumulate in
// registers and are flushed once per round) and consecutive rounds land on different CTAs when a round has fewer
// tiles than CTAs.  TMEM holds two accumulators: the epilogue of tile i overlaps the main loop of tile i+1.
template <int BN, int KC>
__global__ void __launch_bounds__(192, 1) conv_gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using C = TcCfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_out = smem + C::PIPE_BYTES;
  uint64_t* full = (uint64_t*)(stage_out + C::OUT_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* red = (float*)(stage_out + C::OUT_BYTES + 512);   // [SLICES][2][BN] cross-slice reduction scratch

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int n_iters = p.ntaps * p.ncb;
  const int rounds = p.nvar * p.ntn * p.groups * p.ksplit;
  const int k_per = (n_iters + p.ksplit - 1) / p.ksplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // schedule helpers (identical in every role)
  auto round_vng = [&](int r, int& v, int& n, int& g) { r /= p.ksplit; g = r % p.groups; int q = r / p.groups; n = q % p.ntn; v = q / p.ntn; };
  auto k_range = [&](int r, int& it0, int& it1) { const int ks = r % p.ksplit; it0 = ks * k_per; it1 = (it0 + k_per) < n_iters ? (it0 + k_per) : n_iters; };
  auto first_tile = [&](int r) { int rot = (int)(((long long)r * p.mtg) % grid); int c = cta - rot; if (c < 0) c += grid; return c; };
  auto tile_origin = [&](int g, int mloc, int& w0, int& h0, int& b0) {
    int mt = g * p.mtg + mloc;
    const int twi = mt % p.nTw; mt /= p.nTw;
    const int thi = mt % p.nTh; mt /= p.nTh;
    w0 = twi * p.tw; h0 = thi * p.th; b0 = mt * p.tb;
  };

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&p.w_map);
      const uint32_t tx = (uint32_t)(p.rows_valid * KC * 2 + C::B_BYTES);
      uint32_t kit = 0;                      // global k-iteration counter -> smem ring position
      for (int r = 0; r < rounds; ++r) {
        int v, n, g, it0, it1;
        round_vng(r, v, n, g);
        k_range(r, it0, it1);
        for (int mloc = first_tile(r); mloc < p.mtg; mloc += grid) {
          int w0, h0, b0;
          tile_origin(g, mloc, w0, h0, b0);
          for (int it = it0; it < it1; ++it, ++kit) {
            const int s = kit % C::STAGES;
            const uint32_t ph = (kit / C::STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            const int t = it / p.ncb, cb = it - t * p.ncb;
            const EklTap tap = p.taps[v][t];
            uint8_t* sa = smem + s * C::STAGE_BYTES;
            mbar_expect_tx(&full[s], tx);
            tma_load_4d(&p.a_maps[tap.map], &full[s], sa, cb * KC, w0 + tap.dw, h0 + tap.dh, b0);
            if (p.b_mn) {
              // K rows = output channels [cb*KC, +KC), N columns = input channels of master tap src: 64-wide boxes
#pragma unroll
              for (int j = 0; j < (BN >= 64 ? BN / 64 : 1); ++j)
                tma_load_2d(&p.w_map, &full[s], sa + C::A_BYTES + j * (KC * 128), tap.src[0] * p.Nm + n * BN + j * 64, cb * KC);
            } else {
              tma_load_2d(&p.w_map, &full[s], sa + C::A_BYTES, t * p.Cin + cb * KC, v * p.N + n * BN);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t idesc_mn = umma_idesc_bf16(128, BN, 0, 1);
    uint32_t kit = 0, tile = 0;
    for (int r = 0; r < rounds; ++r) {
      int it0, it1;
      k_range(r, it0, it1);
      for (int mloc = first_tile(r); mloc < p.mtg; mloc += grid, ++tile) {
        const uint32_t buf = tile & 1u, use = tile >> 1;
        mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);        // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int it = it0; it < it1; ++it, ++kit) {
          const int s = kit % C::STAGES;
          const uint32_t ph = (kit / C::STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            // descriptors are affine in the shared address: +32 bytes along K == +2 in the (addr >> 4) start field
            const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
            const uint64_t da0 = umma_desc(sa, 16, C::SBO, C::LAYOUT);
            if (p.b_mn) {
              // MN-major B: boxes of [KC k-rows][64 n] (128-byte rows, SWIZZLE_128B); 16 k-rows per MMA = +2048 bytes
              const uint64_t db0 = umma_desc(sa + C::A_BYTES, KC * 128, 8 * 128, 2u);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc_mma_bf16(tacc, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(128 * k), idesc_mn, (it != it0 || k != 0) ? 1u : 0u);
            } else {
              const uint64_t db0 = umma_desc(sa + C::A_BYTES, 16, C::SBO, C::LAYOUT);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc_mma_bf16(tacc, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, (it != it0 || k != 0) ? 1u : 0u);
            }
            tc_commit(&empty[s]);                              // smem stage reusable once these MMAs retire
            if (it == it1 - 1) tc_commit(&tmem_full[buf]);     // accumulator complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5); TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;   // 0..127
    const int quad = et % C::QUADS, slice = et / C::QUADS;      // statistics ownership (BN=256: 64 quads x 2 slices)
    const uint32_t so = smem_u32(stage_out);
    uint32_t tile = 0;
    bool store_pending = false;
    for (int r = 0; r < rounds; ++r) {
      int v, n, g;
      round_vng(r, v, n, g);
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      const bool round_has_tiles = first_tile(r) < p.mtg;
      for (int mloc = first_tile(r); mloc < p.mtg; mloc += grid, ++tile) {
        const uint32_t buf = tile & 1u, use = tile >> 1;
        int w0, h0, b0;
        tile_origin(g, mloc, w0, h0, b0);
        mbar_wait(&tmem_full[buf], use & 1u);
        tc_fence_after();
        // staging tile must be free: the previous tile's TMA store has finished READING it
        if (store_pending && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
        const float* brow = nullptr;
        if (p.bias9 != nullptr) {
          // this thread's pixel (b, h, w) -> border class (3 row classes x 3 column classes)
          const int wi = row % p.tw, hi = (row / p.tw) % p.th, bi = row / (p.tw * p.th);
          const int hh = h0 + hi, ww = w0 + wi;
          const int cls = (hh == 0 ? 0 : (hh == p.H - 1 ? 2 : 1)) * 3 + (ww == 0 ? 0 : (ww == p.W - 1 ? 2 : 1));
          brow = p.bias9 + ((size_t)(b0 + bi) * 9 + cls) * p.N + n * BN;
        }
        if (p.scratch != nullptr) {
          // split-K work item: red-add the fp32 partial tile into scratch[pixel][N]; no staging / statistics / store
          const int wi = row % p.tw, hi = (row / p.tw) % p.th, bi = row / (p.tw * p.th);
          const bool live = row < p.rows_valid && (b0 + bi) < p.mB && (h0 + hi) < p.H && (w0 + wi) < p.W;
          float* dst = p.scratch + ((size_t)((b0 + bi) * p.H + h0 + hi) * p.W + w0 + wi) * p.N + n * BN;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t rr[16];
            tmem_ld16(tacc + (uint32_t)c0, rr);
            tmem_ld_wait();
            if (live) {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(__uint_as_float(rr[i])),
                             "f"(__uint_as_float(rr[i + 1])), "f"(__uint_as_float(rr[i + 2])), "f"(__uint_as_float(rr[i + 3]))
                             : "memory");
            }
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);
          continue;
        }
        if constexpr (BN == 16) {
          uint32_t rr[16];
          tmem_ld16(tacc, rr);
          tmem_ld_wait();
          uint32_t pk[8];
          if (brow != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(brow + i);
              rr[i] = __float_as_uint(__uint_as_float(rr[i]) + bv.x); rr[i + 1] = __float_as_uint(__uint_as_float(rr[i + 1]) + bv.y);
              rr[i + 2] = __float_as_uint(__uint_as_float(rr[i + 2]) + bv.z); rr[i + 3] = __float_as_uint(__uint_as_float(rr[i + 3]) + bv.w);
            }
          }
          if (p.act != 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(epi_act(__uint_as_float(rr[i]), p.act));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
          sts128(so + row * 32, make_uint4(pk[0], pk[1], pk[2], pk[3]));
          sts128(so + row * 32 + 16, make_uint4(pk[4], pk[5], pk[6], pk[7]));
        } else {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t rr[32];
            tmem_ld32(tacc + (uint32_t)c0, rr);
            tmem_ld_wait();
            uint32_t pk[16];
            if (brow != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(brow + c0 + i);
                rr[i] = __float_as_uint(__uint_as_float(rr[i]) + bv.x); rr[i + 1] = __float_as_uint(__uint_as_float(rr[i + 1]) + bv.y);
                rr[i + 2] = __float_as_uint(__uint_as_float(rr[i + 2]) + bv.z); rr[i + 3] = __float_as_uint(__uint_as_float(rr[i + 3]) + bv.w);
              }
            }
            if (p.act != 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(epi_act(__uint_as_float(rr[i]), p.act));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(so + stage_off<BN>(row, c0 + 8 * j), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
          }
        }
        // accumulator drained -> the MMA warp may start the tile after next in this buffer
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
#pragma unroll
          for (int j = 0; j < BN / C::OBOX; ++j)
            tma_store_4d(&p.o_maps[v], stage_out + j * (128 * 128), n * BN + j * C::OBOX, w0, h0, b0);
          tma_store_commit();
        }
        store_pending = true;
        if (p.stats != nullptr) {
          // per-channel partial sums of the bf16 values actually stored, over this thread's row slice
          constexpr int RPS = 128 / C::SLICES;
          const int r0 = slice * RPS;
          const int r1 = (r0 + RPS) < p.rows_valid ? (r0 + RPS) : p.rows_valid;
#pragma unroll 4
          for (int rr = r0; rr < r1; ++rr) {
            const uint2 u = lds64(so + stage_off<BN>(rr, 4 * quad));
            const float x0 = bf16_lo(u.x), x1 = bf16_hi(u.x), x2 = bf16_lo(u.y), x3 = bf16_hi(u.y);
            s1[0] += x0; s2[0] += x0 * x0; s1[1] += x1; s2[1] += x1 * x1;
            s1[2] += x2; s2[2] += x2 * x2; s1[3] += x3; s2[3] += x3 * x3;
          }
        }
      }
      if (p.stats != nullptr && round_has_tiles) {
        // flush this round's partial sums: combine the row slices through smem, then ONE fp64 red.global.add per channel
        // sum into the [group][2][N] statistics buffer (all variants of an up-conv land in the same sums)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float* my = red + (size_t)slice * 2 * BN;
#pragma unroll
        for (int i = 0; i < 4; ++i) { my[4 * quad + i] = s1[i]; my[BN + 4 * quad + i] = s2[i]; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        double* dst = p.stats + (size_t)g * 2 * p.N + n * BN;
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

// ------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a 2-CTA cluster compute ONE 256-pixel x BN-channel tile.
// Each CTA loads its own 128-pixel A patch and HALF of the filter tile (BN/2 rows); the leader CTA's MMA thread issues
// M = 256 instructions that read A and B from both CTAs' shared memory and accumulate into both CTAs' TMEM (128 rows
// each).  Per CTA and k-iteration that is A + B/2 instead of A + B bytes through the L2->SM port -- the bound of the
// single-CTA kernel on every layer wider than the ridge (measured ~85-130 GB/s per SM) -- and half the shared-memory
// reads of B per MMA.  Synchronisation: every TMA of the pair completes on the LEADER's full barrier; the MMA thread's
// tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers; the 2 x 128 epilogue threads release an
// accumulator on the leader's tmem_empty barrier.  Everything else (schedule, epilogue variants, statistics, split-K)
// is the single-CTA kernel's, with "M tile" = pair tile * 2 + CTA rank.
template <int BN, int KC>
struct Tc2Cfg {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = (BN / 2) * KC * 2;            // this CTA's half of the filter tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = 128 * BN * 2;
  static constexpr int STAGES_RAW = (220 * 1024 - OUT_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = PIPE_BYTES + OUT_BYTES + 1024 + 512 + 4096;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);
  static constexpr uint32_t SBO = 8 * KC * 2;
  static constexpr int OBOX = BN < 64 ? BN : 64;
  static constexpr int QUADS = BN / 4;
  static constexpr int SLICES = 128 / QUADS;
};

template <int BN, int KC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) conv_gemm_tc2_kernel(const __grid_constant__ TcParams p) {
  using C = Tc2Cfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_out = smem + C::PIPE_BYTES;
  uint64_t* full = (uint64_t*)(stage_out + C::OUT_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]  (the leader's copies are the live ones)
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* red = (float*)(stage_out + C::OUT_BYTES + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int n_iters = p.ntaps * p.ncb;
  const int rounds = p.nvar * p.ntn * p.groups * p.ksplit;
  const int k_per = (n_iters + p.ksplit - 1) / p.ksplit;
  const int ptg = (p.mtg + 1) >> 1;             // pair tiles per group

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 256); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // both CTAs' barriers are initialised before either is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto round_vng = [&](int r, int& v, int& n, int& g) { r /= p.ksplit; g = r % p.groups; int q = r / p.groups; n = q % p.ntn; v = q / p.ntn; };
  auto k_range = [&](int r, int& it0, int& it1) { const int ks = r % p.ksplit; it0 = ks * k_per; it1 = (it0 + k_per) < n_iters ? (it0 + k_per) : n_iters; };
  auto first_pair = [&](int r) { int rot = (int)(((long long)r * ptg) % npairs); int c = pair - rot; if (c < 0) c += npairs; return c; };
  // this CTA's M tile of pair tile ploc; an odd tile count leaves the last pair's second CTA without one: it runs the same
  // protocol on an out-of-range patch (TMA zero fill) and drops its result
  auto tile_origin = [&](int g, int ploc, int& w0, int& h0, int& b0) {
    const int mloc = 2 * ploc + rank;
    if (mloc >= p.mtg) { w0 = 0; h0 = 0; b0 = p.mB + p.tb; return false; }
    int mt = g * p.mtg + mloc;
    const int twi = mt % p.nTw; mt /= p.nTw;
    const int thi = mt % p.nTh; mt /= p.nTh;
    w0 = twi * p.tw; h0 = thi * p.th; b0 = mt * p.tb;
    return true;
  };

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&p.w_map);
      const uint32_t tx = (uint32_t)(p.rows_valid * KC * 2 + C::B_BYTES);
      uint32_t kit = 0;
      for (int r = 0; r < rounds; ++r) {
        int v, n, g, it0, it1;
        round_vng(r, v, n, g);
        k_range(r, it0, it1);
        for (int ploc = first_pair(r); ploc < ptg; ploc += npairs) {
          int w0, h0, b0;
          tile_origin(g, ploc, w0, h0, b0);
          for (int it = it0; it < it1; ++it, ++kit) {
            const int s = kit % C::STAGES;
            const uint32_t ph = (kit / C::STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            const int t = it / p.ncb, cb = it - t * p.ncb;
            const EklTap tap = p.taps[v][t];
            uint8_t* sa = smem + s * C::STAGE_BYTES;
            if (rank == 0) mbar_expect_tx(&full[s], 2u * tx);            // both CTAs' bytes land on the leader's barrier
            tma_load_4d_2sm(&p.a_maps[tap.map], &full[s], sa, cb * KC, w0 + tap.dw, h0 + tap.dh, b0);
            if (p.b_mn) {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_2d_2sm(&p.w_map, &full[s], sa + C::A_BYTES + j * (KC * 128),
                                tap.src[0] * p.Nm + n * BN + (rank * (BN / 128) + j) * 64, cb * KC);
            } else {
              tma_load_2d_2sm(&p.w_map, &full[s], sa + C::A_BYTES, t * p.Cin + cb * KC, v * p.N + n * BN + rank * (BN / 2));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      constexpr uint32_t idesc_mn = umma_idesc_bf16(256, BN, 0, 1);
      uint32_t kit = 0, tile = 0;
      for (int r = 0; r < rounds; ++r) {
        int it0, it1;
        k_range(r, it0, it1);
        for (int ploc = first_pair(r); ploc < ptg; ploc += npairs, ++tile) {
          const uint32_t buf = tile & 1u, use = tile >> 1;
          mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);        // both CTAs' epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          for (int it = it0; it < it1; ++it, ++kit) {
            const int s = kit % C::STAGES;
            const uint32_t ph = (kit / C::STAGES) & 1u;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
              const uint64_t da0 = umma_desc(sa, 16, C::SBO, C::LAYOUT);
              if (p.b_mn) {
                const uint64_t db0 = umma_desc(sa + C::A_BYTES, KC * 128, 8 * 128, 2u);
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  tc_mma2_bf16(tacc, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(128 * k), idesc_mn, (it != it0 || k != 0) ? 1u : 0u);
              } else {
                const uint64_t db0 = umma_desc(sa + C::A_BYTES, 16, C::SBO, C::LAYOUT);
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  tc_mma2_bf16(tacc, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, (it != it0 || k != 0) ? 1u : 0u);
              }
              tc_commit2(&empty[s]);                             // both CTAs' stage s reusable once these MMAs retire
              if (it == it1 - 1) tc_commit2(&tmem_full[buf]);    // accumulator complete in both CTAs
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5 of BOTH CTAs): this CTA's 128 rows of the pair tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int quad = et % C::QUADS, slice = et / C::QUADS;
    const uint32_t so = smem_u32(stage_out);
    uint32_t tile = 0;
    bool store_pending = false;
    for (int r = 0; r < rounds; ++r) {
      int v, n, g;
      round_vng(r, v, n, g);
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      bool any_valid = false;
      for (int ploc = first_pair(r); ploc < ptg; ploc += npairs, ++tile) {
        const uint32_t buf = tile & 1u, use = tile >> 1;
        int w0, h0, b0;
        const bool valid = tile_origin(g, ploc, w0, h0, b0);
        mbar_wait(&tmem_full[buf], use & 1u);
        tc_fence_after();
        if (!valid) {                                            // no tile for this CTA: just release the accumulator
          tc_fence_before();
          mbar_arrive_leader(&tmem_empty[buf]);
          continue;
        }
        any_valid = true;
        if (store_pending && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
        const float* brow = nullptr;
        if (p.bias9 != nullptr) {
          const int wi = row % p.tw, hi = (row / p.tw) % p.th, bi = row / (p.tw * p.th);
          const int hh = h0 + hi, ww = w0 + wi;
          const int cls = (hh == 0 ? 0 : (hh == p.H - 1 ? 2 : 1)) * 3 + (ww == 0 ? 0 : (ww == p.W - 1 ? 2 : 1));
          brow = p.bias9 + ((size_t)(b0 + bi) * 9 + cls) * p.N + n * BN;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t rr[32];
          tmem_ld32(tacc + (uint32_t)c0, rr);
          tmem_ld_wait();
          uint32_t pk[16];
          if (brow != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(brow + c0 + i);
              rr[i] = __float_as_uint(__uint_as_float(rr[i]) + bv.x); rr[i + 1] = __float_as_uint(__uint_as_float(rr[i + 1]) + bv.y);
              rr[i + 2] = __float_as_uint(__uint_as_float(rr[i + 2]) + bv.z); rr[i + 3] = __float_as_uint(__uint_as_float(rr[i + 3]) + bv.w);
            }
          }
          if (p.act != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(epi_act(__uint_as_float(rr[i]), p.act));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(so + stage_off<BN>(row, c0 + 8 * j), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
        }
        tc_fence_before();
        mbar_arrive_leader(&tmem_empty[buf]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
#pragma unroll
          for (int j = 0; j < BN / C::OBOX; ++j)
            tma_store_4d(&p.o_maps[v], stage_out + j * (128 * 128), n * BN + j * C::OBOX, w0, h0, b0);
          tma_store_commit();
        }
        store_pending = true;
        if (p.stats != nullptr) {
          constexpr int RPS = 128 / C::SLICES;
          const int r0 = slice * RPS;
          const int r1 = (r0 + RPS) < p.rows_valid ? (r0 + RPS) : p.rows_valid;
#pragma unroll 4
          for (int rr = r0; rr < r1; ++rr) {
            const uint2 u = lds64(so + stage_off<BN>(rr, 4 * quad));
            const float x0 = bf16_lo(u.x), x1 = bf16_hi(u.x), x2 = bf16_lo(u.y), x3 = bf16_hi(u.y);
            s1[0] += x0; s2[0] += x0 * x0; s1[1] += x1; s2[1] += x1 * x1;
            s1[2] += x2; s2[2] += x2 * x2; s1[3] += x3; s2[3] += x3 * x3;
          }
        }
      }
      if (p.stats != nullptr && any_valid) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float* my = red + (size_t)slice * 2 * BN;
#pragma unroll
        for (int i = 0; i < 4; ++i) { my[4 * quad + i] = s1[i]; my[BN + 4 * quad + i] = s2[i]; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        double* dst = p.stats + (size_t)g * 2 * p.N + n * BN;
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
  cluster_sync_all();           // the leader's MMAs read the peer's shared memory: neither CTA leaves before both are done
  if (warp == 1) tmem_dealloc2<C::TMEM_COLS>(tmem_base);
}

template <int BN, int KC>
int launch_tc2(TcParams& p, int grid, cudaStream_t st) {
  using C = Tc2Cfg<BN, KC>;
  auto kern = conv_gemm_tc2_kernel<BN, KC>;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  kern<<<grid, 192, C::SMEM_BYTES, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

int make_view_map(CUtensorMap* m, const EklView& v, int boxC, int tw, int th, int tb, int swz) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
  uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
  uint32_t box[4] = {(uint32_t)boxC, (uint32_t)tw, (uint32_t)th, (uint32_t)tb};
  return ekl_make_tmap(m, v.base, 4, dims, strides, box, swz, 2);
}

template <int BN, int KC>
int launch_tc(const EklGather* g, TcParams& p, int grid, cudaStream_t st) {
  using C = TcCfg<BN, KC>;
  auto kern = conv_gemm_tc_kernel<BN, KC>;
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

int ekl_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      n = 148;
  }
  return n;
}

// stats: [group][2][N] fp64 sums (accumulated; zero on entry) or null.
// Split-K plan of a gather-GEMM (0 = not split): few tiles and a long contraction (the 4x4 / 8x8 discriminator tails:
// M = 384..1152 rows, K up to 18432) leave most SMs idle while each busy SM is bound by its own operand ingest, so the
// (tap, channel-block) iterations of a tile are spread over up to 4 work items that red-add fp32 partial tiles into a
// scratch buffer; splitk_finish (bn_act.cu) rounds to bf16, takes the BatchNorm statistics and re-zeroes the scratch.
static void tc_tile_plan(const EklGather* g, int group_b, bool allow_split, int* BN_out, int* ks_out) {
  int tb, th, tw;
  ekl_tc_geometry(g, group_b, &tb, &th, &tw);
  const int groups = (group_b > 0 && g->mB % group_b == 0) ? g->mB / group_b : 1;
  const int mtiles = ekl_cdiv(g->mW, tw) * ekl_cdiv(g->mH, th) * ekl_cdiv(g->mB / groups, tb) * groups;
  const int KC = g->Cin % 64 == 0 ? 64 : (g->Cin % 32 == 0 ? 32 : 16);
  const int n_iters = g->ntaps * (g->Cin / KC);
  const int sms = ekl_num_sms();
  const int bn_min = g->N % 64 == 0 ? 64 : 16;      // measured: tiles narrower than 64 only add per-iteration overhead
  double best = -1.0;
  int BN = 16, KS = 1;
  for (int bn = 256; bn >= bn_min; bn >>= 1) {
    if (g->N % bn != 0) continue;
    const int64_t tiles = (int64_t)mtiles * (g->N / bn) * g->nvar;
    int ks = 1;
    // a plan splits when split_min x its work items fit the SMs (EKL_TC_SPLIT_MIN).  2 since the finishing pass is fused
    // with the BatchNorm that follows: config 2, four A/B pairs, 7.642-7.652 ms/step at 3 vs 7.605-7.617 at 2 (the
    // 1024 -> 2048 discriminator conv at 72 images, 72 work items, now runs as 144 half-length ones)
    static int split_min = -1;
    if (split_min < 0) { const char* e = getenv("EKL_TC_SPLIT_MIN"); split_min = e ? atoi(e) : 2; if (split_min < 2) split_min = 2; }
    if (allow_split && g->nvar == 1 && KC == 64 && tiles * split_min <= sms && n_iters >= 64) {
      ks = (int)(sms / tiles);
      if (ks > n_iters / 16) ks = n_iters / 16;
      if (ks > 4) ks = 4;
      if (ks < 1) ks = 1;
    }
    // bytes the busiest CTA fetches (per-SM operand ingest is the bound: ~85 GB/s measured whether 48 or 148 SMs run), but
    // never less than the plan's TOTAL operand traffic divided by rho SM-equivalents of L2 fabric.  The second term makes
    // few-tile layers (the 4x4 / 8x8 tails) take wider N tiles = fewer work items.  Measured (profiles/r02_summary.md):
    // timed alone those layers are a wash (+-10 % either way, +1.8 % in sum), but inside the step, where the
    // discriminator branches run concurrently and a persistent conv CTA owns its SM, fewer busy SMs per kernel let the
    // other branches' kernels in: config 2 8.43 -> 8.19 ms/step for any rho in 20..60, configs 1 / 4 +2 % / +1 %, coco
    // unchanged.  EKL_TC_RHO=0 restores the per-CTA term alone.
    static double rho = -1.0;
    if (rho < 0) { const char* e = getenv("EKL_TC_RHO"); rho = e ? atof(e) : 45.0; }
    double cost = (double)((tiles * ks + sms - 1) / sms) * (128 + bn) / ks;
    if (rho > 0) { const double total = (double)tiles * (128 + bn) / rho; if (total > cost) cost = total; }
    if (best < 0 || cost < best) { best = cost; BN = bn; KS = ks; }
  }
  *BN_out = BN; *ks_out = KS;
}

// fp32 scratch elements the split-K path of this plan needs (0: the plan is not split)
int64_t ekl_tc_split_elems(const EklGather* g, int group_b) {
  int BN, ks;
  tc_tile_plan(g, group_b, true, &BN, &ks);
  return ks > 1 ? (int64_t)g->mB * g->mH * g->mW * g->N : 0;
}

// 1 if the data-gradient plan g can read the FORWARD-packed filter ([Cout][tap][Cin], i.e. the bf16 shadow of a
// channels_last master) as an MN-major B operand: one master tap per packed tap, 64-channel K blocks, N % 64 == 0
int ekl_tc_dgrad_from_fwd_ok(const EklGather* g) {
  if (!g->transposed || g->Cin % 64 != 0 || g->N % 64 != 0) return 0;
  for (int v = 0; v < g->nvar; ++v)
    for (int t = 0; t < g->ntaps; ++t)
      if (g->taps[v][t].nsrc != 1) return 0;
  return 1;
}

int ekl_gather_gemm_tc(const EklGather* g, const void* w_packed, double* stats, int group_b, int act, const float* bias9,
                       float* scratch, int* mtiles_out, cudaStream_t st, int w_is_fwd_packed) {
  EKL_REQUIRE(ekl_tc_supported(g), "gather_gemm_tc: unsupported shape Cin=%d N=%d mH=%d mW=%d", g->Cin, g->N, g->mH, g->mW);
  TcParams p;
  memset(&p, 0, sizeof(p));
  int tb, th, tw;
  ekl_tc_geometry(g, group_b, &tb, &th, &tw);
  p.tb = tb; p.th = th; p.tw = tw;
  p.nTw = ekl_cdiv(g->mW, tw); p.nTh = ekl_cdiv(g->mH, th);
  p.groups = (group_b > 0 && g->mB % group_b == 0) ? g->mB / group_b : 1;
  const int gb = g->mB / p.groups;
  p.mtg = p.nTw * p.nTh * ekl_cdiv(gb, tb);
  const int mtiles = p.mtg * p.groups;
  if (mtiles_out) *mtiles_out = mtiles;
  p.rows_valid = tb * th * tw;
  p.ntaps = g->ntaps; p.Cin = g->Cin; p.N = g->N; p.stats = stats; p.nvar = g->nvar; p.act = act;
  p.bias9 = bias9; p.H = g->mH; p.W = g->mW;
  EKL_REQUIRE(bias9 == nullptr || (g->nvar == 1 && g->ntaps == 9 && g->N % 4 == 0), "bias9: stride-1 3x3 forward only");
  memcpy(p.taps, g->taps, sizeof(p.taps));
  const int KC = g->Cin % 64 == 0 ? 64 : (g->Cin % 32 == 0 ? 32 : 16);
  p.ncb = g->Cin / KC;
  const int swz = KC == 64 ? 3 : (KC == 32 ? 2 : 1);
  int BN, ksplit;
  tc_tile_plan(g, group_b, scratch != nullptr, &BN, &ksplit);
  EKL_REQUIRE(scratch == nullptr || (ksplit > 1 && act == 0 && bias9 == nullptr), "split-K scratch passed to a plan that is not split");
  p.ksplit = ksplit; p.scratch = ksplit > 1 ? scratch : nullptr; p.mB = g->mB;
  if (ksplit > 1) p.stats = nullptr;      // statistics come from splitk_finish
  if (const char* e = getenv("EKL_TC_BN")) {          // experiment knob (unsplit plans only)
    const int v = atoi(e);
    if (ksplit == 1 && (v == 16 || v == 32 || v == 64 || v == 128 || v == 256) && g->N % v == 0) BN = v;
  }
  p.ntn = g->N / BN;
  // CTA-pair kernel (conv_gemm_tc2_kernel): 64-channel K blocks, N tiles of >= 128 channels, no split-K, and at least 12
  // M tiles per (group, variant).  The threshold is measured (profiles/r02_summary.md, per-layer A/B on config 2): layers
  // with >= 12 M tiles gain 6-14 %, the weight-dominated layers below it (3-9 M tiles) lose 40-70 %.
  // EKL_TC2 = 0 never, 1 (default) by the rule above, 2 whenever the kernel can run at all.
  static int tc2_mode = -1;
  if (tc2_mode < 0) { const char* e = getenv("EKL_TC2"); tc2_mode = e ? atoi(e) : 1; }
  const int sms = ekl_num_sms();
  const bool use2 = tc2_mode > 0 && KC == 64 && ksplit == 1 && BN >= 128 && sms % 2 == 0 && (tc2_mode >= 2 || p.mtg >= 12);
  for (int i = 0; i < g->n_a; ++i) {
    int rc = make_view_map(&p.a_maps[i], g->a[i], KC, tw, th, tb, swz);
    if (rc) return rc;
  }
  const int obox = BN < 64 ? BN : 64;
  for (int i = 0; i < g->nvar; ++i) {
    int rc = make_view_map(&p.o_maps[i], g->o[i], obox, tw, th, tb, BN >= 64 ? 3 : (BN == 32 ? 2 : 0));
    if (rc) return rc;
  }
  if (w_is_fwd_packed) {
    EKL_REQUIRE(ekl_tc_dgrad_from_fwd_ok(g) && KC == 64 && BN >= 64, "dgrad from the forward-packed filter: unsupported plan");
    p.b_mn = 1; p.Nm = g->N;
    const int KK = g->KH * g->KW;
    uint64_t dims[2] = {(uint64_t)KK * g->N, (uint64_t)g->Cin};            // [Cout rows][tap][Cin]
    uint64_t strides[1] = {(uint64_t)KK * g->N * 2};
    uint32_t box[2] = {64u, (uint32_t)KC};
    int rc = ekl_make_tmap(&p.w_map, w_packed, 2, dims, strides, box, 3, 2);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)g->ntaps * g->Cin, (uint64_t)g->nvar * g->N};
    uint64_t strides[1] = {(uint64_t)g->ntaps * g->Cin * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)(use2 ? BN / 2 : BN)};       // a CTA of a pair loads half of the filter tile
    int rc = ekl_make_tmap(&p.w_map, w_packed, 2, dims, strides, box, swz, 2);
    if (rc) return rc;
  }
  const int grid = sms;     // persistent: one CTA per SM
  if (use2) {
    if (BN == 256) return launch_tc2<256, 64>(p, grid, st);
    return launch_tc2<128, 64>(p, grid, st);
  }
#define EKL_TC_CASE(bn, kc) if (BN == bn && KC == kc) return launch_tc<bn, kc>(g, p, grid, st);
  EKL_TC_CASE(256, 64) EKL_TC_CASE(128, 64) EKL_TC_CASE(64, 64) EKL_TC_CASE(32, 64) EKL_TC_CASE(16, 64)
  EKL_TC_CASE(256, 32) EKL_TC_CASE(128, 32) EKL_TC_CASE(64, 32) EKL_TC_CASE(32, 32) EKL_TC_CASE(16, 32)
  EKL_TC_CASE(256, 16) EKL_TC_CASE(128, 16) EKL_TC_CASE(64, 16) EKL_TC_CASE(32, 16) EKL_TC_CASE(16, 16)
#undef EKL_TC_CASE
  return ekl_fail(-1, "gather_gemm_tc: no kernel for BN=%d KC=%d", BN, KC);
}
