// tcgen05 weight-gradient kernel.
//
//   dW[co][src(v,t)][ci] += sum_{p in M grid} dY_v[p, co] * A_{map(v,t)}[p + (dh,dw), ci]
//
// GEMM view: D[(t,ci) 128 rows][co BNW cols] accumulated over PIXELS (the GEMM K dimension).  Both operands are
// pixel-major in memory (NHWC), i.e. "MN-major" UMMA operands: a TMA box [64 pixels][CW channels] is a K x MN
// tile whose 128/64/32-byte rows are exactly the canonical SWIZZLE_{128,64,32}B MN-major atoms
// (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>): SBO = 8 pixel rows, LBO = one box.
// The 128 accumulator rows are 128/CW boxes = consecutive (tap, channel-block) pairs sharing one dY operand.
// Split-K over pixel tiles across CTAs; epilogue = tcgen05.ld + red.global.add.f32 straight into the fp32
// gradient buffer in master layout [Cout][KH*KW][Cin] (coalesced along ci).
#include <stdlib.h>

#include "conv_plan.h"
#include "ekl_common.cuh"

namespace {

constexpr int KP = 64;   // pixels per pipeline stage

struct WgParams {
  CUtensorMap a_maps[4];
  CUtensorMap y_maps[EKL_MAX_VAR];
  EklTap taps[EKL_MAX_VAR][EKL_MAX_TAPS];
  float* dw;
  int ntaps, ncb, Cin, Cout, KK, kcrs;
  int ld, off, cout_valid; // master gradient row pitch / channel offset / real output channels
  int tiles_per_var;       // M tiles per variant
  int tb, th, tw, nTh, nTw, ptiles;   // pixel tiling
  int rows_valid;          // tb*th*tw (<= KP)
};

template <int BNW, int CW>
struct WgCfg {
  static constexpr int BPT = 128 / CW;                 // A boxes per M tile
  static constexpr int A_BOX = KP * CW * 2;
  static constexpr int A_BYTES = 128 * KP * 2;
  static constexpr int YW = BNW < 64 ? BNW : 64;       // channels per dY box
  static constexpr int Y_BOX = KP * YW * 2;
  static constexpr int B_BYTES = BNW * KP * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (96 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = BNW < 32 ? 32 : BNW;
  static constexpr uint32_t A_LAYOUT = CW == 64 ? 2u : (CW == 32 ? 4u : 6u);
  static constexpr uint32_t Y_LAYOUT = YW == 64 ? 2u : (YW == 32 ? 4u : 6u);
};

template <int BNW, int CW>
__global__ void __launch_bounds__(192) conv_wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  using C = WgCfg<BNW, CW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x / p.tiles_per_var;
  const int mt = blockIdx.x - v * p.tiles_per_var;
  const int n0 = blockIdx.y * BNW;
  const int npairs = p.ntaps * p.ncb;                  // (tap, channel block) pairs of this variant
  const int pair0 = mt * C::BPT;
  const int nbox = (npairs - pair0) < C::BPT ? (npairs - pair0) : C::BPT;
  // pixel tiles of this split
  const int per = (p.ptiles + gridDim.z - 1) / gridDim.z;
  const int pt0 = blockIdx.z * per;
  const int pt1 = (pt0 + per) < p.ptiles ? (pt0 + per) : p.ptiles;
  const int n_iters = pt1 - pt0;

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

  if (n_iters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t tx = (uint32_t)(nbox * p.rows_valid * CW * 2 + (BNW / C::YW) * p.rows_valid * C::YW * 2);
        for (int it = 0; it < n_iters; ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          int pt = pt0 + it;
          const int twi = pt % p.nTw; pt /= p.nTw;
          const int thi = pt % p.nTh; pt /= p.nTh;
          const int w0 = twi * p.tw, h0 = thi * p.th, b0 = pt * p.tb;
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          mbar_expect_tx(&full[s], tx);
          for (int j = 0; j < nbox; ++j) {
            const int pair = pair0 + j;
            const int t = pair / p.ncb, cb = pair - t * p.ncb;
            const EklTap tap = p.taps[v][t];
            tma_load_4d(&p.a_maps[tap.map], &full[s], sa + j * C::A_BOX, cb * CW, w0 + tap.dw, h0 + tap.dh, b0);
          }
#pragma unroll
          for (int j = 0; j < BNW / C::YW; ++j)
            tma_load_4d(&p.y_maps[v], &full[s], sa + C::A_BYTES + j * C::Y_BOX, n0 + j * C::YW, w0, h0, b0);
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BNW, 1, 1);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const int ksteps = (p.rows_valid + 15) / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = umma_desc(sa + k * 16 * CW * 2, C::A_BOX, 8 * CW * 2, C::A_LAYOUT);
            const uint64_t db = umma_desc(sb + k * 16 * C::YW * 2, C::Y_BOX, 8 * C::YW * 2, C::Y_LAYOUT);
            tc_mma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty[s]);
          if (it == n_iters - 1) tc_commit(tmem_full);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int j = row / CW;                       // box index -> (tap, channel block)
      const bool valid = j < nbox;
      const int pair = pair0 + (valid ? j : 0);
      const int t = pair / p.ncb, cb = pair - t * p.ncb;
      const EklTap tap = p.taps[v][t];
      const int ci = cb * CW + (row % CW);
#pragma unroll 1
      for (int c0 = 0; c0 < BNW; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c0 + i >= BNW) break;                 // BNW = 16: the TMEM allocation is 32 columns, 16 are live
            const int co = n0 + c0 + i;
            const float val = __uint_as_float(r[i]);
            if (co < p.cout_valid)
              for (int s = 0; s < tap.nsrc; ++s)
                atomicAdd(p.dw + EKL_WIDX(p.kcrs, co, tap.src[s], p.off + ci, p.KK, p.ld), val);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// Big-channel variant (Cin % 64 == 0, Cout % 128 == 0, KRSC master layout): the accumulator is transposed,
//   D[co 128 rows][(tap, ci) up to 256 columns],
// so a TMEM lane (= one output channel co) holds runs of 32 CONSECUTIVE ci of one tap: the epilogue moves 16 bytes per
// lane-instruction (ld/st.global.v4 when this CTA is the only writer of its outputs, red.global.add.v4.f32 otherwise)
// instead of one scalar red per element.  (Measured on B200: the scalar-red epilogue of a 128x256 tile costs 25-45 us,
// REDG issues at ~1.3 cycles per lane.)
struct WgCoParams {
  CUtensorMap a_maps[4];
  CUtensorMap y_maps[EKL_MAX_VAR];
  EklTap taps[EKL_MAX_VAR][EKL_MAX_TAPS];
  float* dw;
  int ntaps, ncb, Cin, Cout, KK;
  int ld, off;             // master gradient row pitch / channel offset
  int tiles_per_var;
  int tb, th, tw, nTh, nTw, ptiles;
  int rows_valid;
  int exclusive;           // 1: every output element is written by exactly one CTA -> plain read-modify-write
};

struct WgCoCfg {
  static constexpr int X_BOX = KP * 64 * 2;          // [64 px][64 ci]
  static constexpr int Y_BOX = KP * 64 * 2;          // [64 px][64 co]
  static constexpr int A_BYTES = 2 * Y_BOX;          // M = 128 co
  static constexpr int B_BYTES = 4 * X_BOX;          // N <= 256 (tap, ci) columns
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(192) conv_wgrad_co_kernel(const __grid_constant__ WgCoParams p) {
  using C = WgCoCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x / p.tiles_per_var;
  const int nt = blockIdx.x - v * p.tiles_per_var;
  const int co0 = blockIdx.y * 128;
  const int npairs = p.ntaps * p.ncb;
  const int pair0 = nt * 4;
  const int nbox = (npairs - pair0) < 4 ? (npairs - pair0) : 4;
  const int per = (p.ptiles + gridDim.z - 1) / gridDim.z;
  const int pt0 = blockIdx.z * per;
  const int pt1 = (pt0 + per) < p.ptiles ? (pt0 + per) : p.ptiles;
  const int n_iters = pt1 - pt0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (n_iters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t tx = (uint32_t)((nbox + 2) * p.rows_valid * 64 * 2);
        for (int it = 0; it < n_iters; ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          int pt = pt0 + it;
          const int twi = pt % p.nTw; pt /= p.nTw;
          const int thi = pt % p.nTh; pt /= p.nTh;
          const int w0 = twi * p.tw, h0 = thi * p.th, b0 = pt * p.tb;
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          mbar_expect_tx(&full[s], tx);
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_4d(&p.y_maps[v], &full[s], sa + j * C::Y_BOX, co0 + j * 64, w0, h0, b0);
          for (int j = 0; j < nbox; ++j) {
            const int pair = pair0 + j;
            const int t = pair / p.ncb, cb = pair - t * p.ncb;
            const EklTap tap = p.taps[v][t];
            tma_load_4d(&p.a_maps[tap.map], &full[s], sa + C::A_BYTES + j * C::X_BOX, cb * 64, w0 + tap.dw, h0 + tap.dh, b0);
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = umma_idesc_bf16(128, nbox * 64, 1, 1);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const int ksteps = (p.rows_valid + 15) / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = umma_desc(sa + k * 16 * 128, C::Y_BOX, 8 * 128, 2u);
            const uint64_t db = umma_desc(sb + k * 16 * 128, C::X_BOX, 8 * 128, 2u);
            tc_mma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty[s]);
          if (it == n_iters - 1) tc_commit(tmem_full);
        }
        __syncwarp();
      }
    } else {
      // TMEM lane = output channel; each warp transposes its 32(co) x 32(ci) block through shared memory (the pipeline
      // stages are free: every MMA has retired) so that 8 consecutive lanes cover one 128-byte run of ci:
      // every global access below is 4 output-channel rows x 128 contiguous bytes.
      const int q = warp & 3;
      float* stg = reinterpret_cast<float*>(smem) + q * (32 * 36);
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int rsub = lane >> 3, col = (lane & 7) * 4;
#pragma unroll 1
      for (int c0 = 0; c0 < nbox * 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(stg + lane * 36 + i) =
              make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        __syncwarp();
        const int pair = pair0 + (c0 >> 6);
        const int t = pair / p.ncb, cb = pair - t * p.ncb;
        const EklTap tap = p.taps[v][t];
        const int ci = cb * 64 + (c0 & 63) + col;
        float4 val[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) val[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + rsub) * 36 + col);
        for (int s = 0; s < tap.nsrc; ++s) {
          float* dst = p.dw + ((int64_t)(co0 + q * 32 + rsub) * p.KK + tap.src[s]) * p.ld + p.off + ci;
          const int64_t rstride = (int64_t)4 * p.KK * p.ld;
          if (p.exclusive) {
            float4 o[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) o[it] = *reinterpret_cast<const float4*>(dst + it * rstride);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              o[it].x += val[it].x; o[it].y += val[it].y; o[it].z += val[it].z; o[it].w += val[it].w;
              *reinterpret_cast<float4*>(dst + it * rstride) = o[it];
            }
          } else {
#pragma unroll
            for (int it = 0; it < 8; ++it) red_add_v4(dst + it * rstride, val[it].x, val[it].y, val[it].z, val[it].w);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// Halo variant for the small-channel stride-1 3x3 layers (Cin in {16, 32}: ResBlocks / folded jointConv of stage 3,
// image heads, discriminator stem): huge pixel count, tiny filter.  The generic kernel re-fetches the x tile for each of
// the 9 taps and the dY tile for each (tap, channel-block) group; here a CTA walks 16-row x 8-column pixel tiles and per
// tile loads ONE 18 x 10 halo box of x and ONE dY box.  Accumulator D[co 128 lanes][(tap, ci) 9*Cin columns] stays in
// TMEM over all of the CTA's tiles (split-K over tiles across CTAs), so a tap is an MMA with
//   A = dY tile        (pixel-major = MN-major, M = co; channels beyond Cout are TMA zero fill)
//   B = x halo rows shifted by the tap offset (MN-major, N = Cin; UMMA descriptors swizzle on absolute address bits, so
//       a row-shifted start inside the swizzled box is legal -- verified against the reference in conv_rw.cu's halo mode)
// and the epilogue is the transposed-accumulator one: 16-byte red.global.add per lane into dw[co][tap][ci].
struct WgHaloParams {
  CUtensorMap x_map, y_map;
  EklTap taps[9];
  float* dw;
  int Cin, Cout, nTh, nTw, tiles;
  int ld, off, cout_valid; // master gradient row pitch / channel offset / real output channels
};

template <int CIN>
struct WgHaloCfg {
  static constexpr int ROWB = CIN * 2;
  static constexpr int X_BYTES = ((180 * ROWB + 1023) / 1024) * 1024;
  static constexpr int Y_BOX = 128 * 128;                  // [128 px][64 co]
  static constexpr int STAGE_BYTES = X_BYTES + 2 * Y_BOX;
  static constexpr int STAGES = 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;                    // 9 * CIN <= 288 columns used
  static constexpr uint32_t X_LAYOUT = CIN == 32 ? 4u : 6u; // SW64 / SW32
};

template <int CIN>
__global__ void __launch_bounds__(192, 1) conv_wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  using C = WgHaloCfg<CIN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int ybox = p.Cout > 64 ? 2 : 1;
  const bool has_tiles = cta < p.tiles;

  // the second dY box is never written when Cout <= 64: zero it once so accumulator rows 64..127 stay finite
  if (ybox == 1)
    for (int s = 0; s < C::STAGES; ++s)
      for (int i = threadIdx.x; i < C::Y_BOX / 16; i += 192)
        reinterpret_cast<uint4*>(smem + s * C::STAGE_BYTES + C::X_BYTES + C::Y_BOX)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (has_tiles) {
    if (warp == 0) {
      if (elect_one()) {
        tma_prefetch_desc(&p.x_map);
        tma_prefetch_desc(&p.y_map);
        uint32_t kit = 0;
        for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
          const int s = kit % C::STAGES;
          const uint32_t ph = (kit / C::STAGES) & 1u;
          int t = tile;
          const int twi = t % p.nTw; t /= p.nTw;
          const int thi = t % p.nTh; t /= p.nTh;
          const int w0 = twi * 8, h0 = thi * 16, b0 = t;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sx = smem + s * C::STAGE_BYTES;
          mbar_expect_tx(&full[s], (uint32_t)(180 * C::ROWB + ybox * C::Y_BOX));
          tma_load_4d(&p.x_map, &full[s], sx, 0, w0 - 1, h0 - 1, b0);
          for (int j = 0; j < ybox; ++j) tma_load_4d(&p.y_map, &full[s], sx + C::X_BYTES + j * C::Y_BOX, j * 64, w0, h0, b0);
        }
      }
    } else if (warp == 1) {
      // The three taps of one filter row (dw = -1, 0, +1) are the SAME pixel rows shifted by one row each, so they are
      // issued as ONE MMA with N = 3*CIN whose N-atoms are ROWB bytes apart (LBO = one pixel row: overlapping atoms);
      // the dY operand is then read from shared memory 3 times per k-step instead of 9.
      constexpr uint32_t idesc = umma_idesc_bf16(128, 3 * CIN, 1, 1);
      uint32_t kit = 0;
      for (int tile = cta; tile < p.tiles; tile += grid, ++kit) {
        const int s = kit % C::STAGES;
        const uint32_t ph = (kit / C::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sx = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sy = sx + C::X_BYTES;
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {                       // 16 pixels = image rows 2j, 2j+1 of the tile
            const uint64_t da = umma_desc(sy + j * 16 * 128, C::Y_BOX, 8 * 128, 2u);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              // K rows of B: 8 pixels of halo row (2j + kh), halo column 0.., then the next halo row (SBO = 10 rows)
              const uint64_t db = umma_desc(sx + (uint32_t)((2 * j + kh) * 10) * C::ROWB, C::ROWB, 10 * C::ROWB, C::X_LAYOUT);
              tc_mma_bf16(tmem_base + (uint32_t)(kh * 3 * CIN), da, db, idesc, (kit | (uint32_t)j) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(tmem_full);
      __syncwarp();
    } else {
      // epilogue: lane = output channel; 9*CIN columns = (tap, ci); 32x32 blocks transposed through smem (free by now)
      const int q = warp & 3;
      float* stg = reinterpret_cast<float*>(smem) + q * (32 * 36);
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int rsub = lane >> 3, col = (lane & 7) * 4;
      if (q * 32 < p.Cout) {
#pragma unroll 1
        for (int c0 = 0; c0 < 9 * CIN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(stg + lane * 36 + i) =
                make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          __syncwarp();
          // columns c0+col .. c0+col+3 belong to one tap (CIN is a multiple of 16 >= 16 and col % 4 == 0)
          const int cc = c0 + col;
          const bool in_range = cc < 9 * CIN;                  // CIN = 16: the last 32-column chunk is half used
          const int t = in_range ? cc / CIN : 0, ci = cc - t * CIN;
          const int src = p.taps[t].src[0];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int co = q * 32 + it * 4 + rsub;
            if (in_range && co < p.cout_valid) {
              const float4 v = *reinterpret_cast<const float4*>(stg + (it * 4 + rsub) * 36 + col);
              red_add_v4(p.dw + ((int64_t)co * 9 + src) * p.ld + p.off + ci, v.x, v.y, v.z, v.w);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

int make_map(CUtensorMap* m, const EklView& v, int boxC, int tw, int th, int tb, int swz) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
  uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
  uint32_t box[4] = {(uint32_t)boxC, (uint32_t)tw, (uint32_t)th, (uint32_t)tb};
  return ekl_make_tmap(m, v.base, 4, dims, strides, box, swz, 2);
}

template <int BNW, int CW>
int launch_wg(const EklGather* g, WgParams& p, int splits, cudaStream_t st) {
  using C = WgCfg<BNW, CW>;
  auto kern = conv_wgrad_tc_kernel<BNW, CW>;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(p.tiles_per_var * g->nvar, g->N / BNW, splits);
  kern<<<grid, 192, C::SMEM_BYTES, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int ekl_wgrad_tc_supported(const EklGather* g) {
  auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
  if (g->transposed) return 0;
  if (g->Cin % 16 != 0 || g->N % 16 != 0) return 0;
  if (!pow2(g->mW) || !pow2(g->mH)) return 0;
  for (int i = 0; i < g->n_a; ++i)
    if (g->a[i].f32 || g->a[i].sC != 1) return 0;
  for (int i = 0; i < g->nvar; ++i)
    if (g->o[i].f32 || g->o[i].sC != 1) return 0;
  return 1;
}

int ekl_num_sms();

// big-channel variant: see conv_wgrad_co_kernel
static int wgrad_co(const EklGather* g, float* dw, cudaStream_t st) {
  WgCoParams p;
  memset(&p, 0, sizeof(p));
  memcpy(p.taps, g->taps, sizeof(p.taps));
  p.dw = dw; p.ntaps = g->ntaps; p.Cin = g->Cin; p.Cout = g->N; p.KK = g->KH * g->KW;
  p.ld = g->w_ld; p.off = g->w_off;
  p.ncb = g->Cin / 64;
  p.tiles_per_var = ekl_cdiv(g->ntaps * p.ncb, 4);
  int tw = g->mW < KP ? g->mW : KP;
  int rest = KP / tw;
  int th = g->mH < rest ? g->mH : rest;
  int tb = rest / th;
  p.tw = tw; p.th = th; p.tb = tb;
  p.nTw = ekl_cdiv(g->mW, tw); p.nTh = ekl_cdiv(g->mH, th);
  p.ptiles = p.nTw * p.nTh * ekl_cdiv(g->mB, tb);
  p.rows_valid = tb * th * tw;
  for (int i = 0; i < g->n_a; ++i)
    if (int rc = make_map(&p.a_maps[i], g->a[i], 64, tw, th, tb, 3)) return rc;
  for (int i = 0; i < g->nvar; ++i)
    if (int rc = make_map(&p.y_maps[i], g->o[i], 64, tw, th, tb, 3)) return rc;
  const int base_ctas = p.tiles_per_var * g->nvar * (g->N / 128);
  int splits = (2 * ekl_num_sms()) / base_ctas;
  if (splits > p.ptiles / 8) splits = p.ptiles / 8;
  if (splits < 1) splits = 1;
  if (const char* e = getenv("EKL_WG_SPLITS")) {
    const int v = atoi(e);
    if (v > 0) splits = v < p.ptiles ? v : p.ptiles;
  }
  // exclusive ownership: one split, and no master tap receives more than one (variant, tap) product
  bool multi = false;
  for (int v = 0; v < g->nvar; ++v)
    for (int t = 0; t < g->ntaps; ++t) multi = multi || g->taps[v][t].nsrc != 1;
  p.exclusive = (splits == 1 && g->nvar == 1 && !multi) ? 1 : 0;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_co_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCoCfg::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(p.tiles_per_var * g->nvar, g->N / 128, splits);
  conv_wgrad_co_kernel<<<grid, 192, WgCoCfg::SMEM_BYTES, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

// halo variant: see conv_wgrad_halo_kernel
static int wgrad_halo_ok(const EklGather* g) {
  if (g->w_kcrs || g->nvar != 1 || g->ntaps != 9 || g->n_a != 1) return 0;
  if (!(g->Cin == 16 || g->Cin == 32) || g->N % 16 != 0 || g->N > 128) return 0;
  if (g->mH % 16 != 0 || g->mW % 8 != 0) return 0;
  for (int t = 0; t < 9; ++t) {      // taps in (kh, kw) raster order: tap t reads pixel offset (t/3 - 1, t%3 - 1)
    const EklTap& tp = g->taps[0][t];
    if (tp.nsrc != 1 || tp.dh != t / 3 - 1 || tp.dw != t % 3 - 1) return 0;
  }
  return 1;
}

template <int CIN>
static int launch_wgrad_halo(WgHaloParams& p, int grid, cudaStream_t st) {
  using C = WgHaloCfg<CIN>;
  auto kern = conv_wgrad_halo_kernel<CIN>;
  static bool attr_done = false;
  if (!attr_done) {
    EKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  kern<<<grid, 192, C::SMEM_BYTES, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

static int wgrad_halo(const EklGather* g, float* dw, cudaStream_t st) {
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  memcpy(p.taps, g->taps[0], sizeof(p.taps));
  p.dw = dw; p.Cin = g->Cin; p.Cout = g->N;
  p.ld = g->w_ld; p.off = g->w_off; p.cout_valid = g->w_cout;
  p.nTh = g->mH / 16; p.nTw = g->mW / 8; p.tiles = g->mB * p.nTh * p.nTw;
  {
    const EklView& v = g->a[0];
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
    uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
    uint32_t box[4] = {(uint32_t)g->Cin, 10, 18, 1};
    if (int rc = ekl_make_tmap(&p.x_map, v.base, 4, dims, strides, box, g->Cin == 32 ? 2 : 1, 2)) return rc;
  }
  {
    const EklView& v = g->o[0];
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.dW, (uint64_t)v.dH, (uint64_t)v.dB};
    uint64_t strides[3] = {(uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sB * 2};
    uint32_t box[4] = {64, 8, 16, 1};
    if (int rc = ekl_make_tmap(&p.y_map, v.base, 4, dims, strides, box, 3, 2)) return rc;
  }
  int grid = ekl_num_sms();
  if (grid > p.tiles) grid = p.tiles;
  return g->Cin == 32 ? launch_wgrad_halo<32>(p, grid, st) : launch_wgrad_halo<16>(p, grid, st);
}

// fwd_plan: forward plan whose `o` views hold dY.  dw: master-layout fp32 gradient, ACCUMULATED into.
int ekl_wgrad_tc(const EklGather* g, float* dw, cudaStream_t st) {
  EKL_REQUIRE(ekl_wgrad_tc_supported(g), "wgrad_tc: unsupported shape Cin=%d Cout=%d", g->Cin, g->N);
  {
    static int halo_on = -1;
    if (halo_on < 0) { const char* e = getenv("EKL_DISABLE_WGRAD_HALO"); halo_on = (e && e[0] == '1') ? 0 : 1; }
    if (halo_on && wgrad_halo_ok(g)) return wgrad_halo(g, dw, st);
  }
  {
    static int co_on = -1;
    if (co_on < 0) { const char* e = getenv("EKL_DISABLE_WGRAD_CO"); co_on = (e && e[0] == '1') ? 0 : 1; }
    if (co_on && !g->w_kcrs && g->Cin % 64 == 0 && g->N % 128 == 0 && g->w_cout == g->N) return wgrad_co(g, dw, st);
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  memcpy(p.taps, g->taps, sizeof(p.taps));
  p.dw = dw; p.ntaps = g->ntaps; p.Cin = g->Cin; p.Cout = g->N; p.KK = g->KH * g->KW; p.kcrs = g->w_kcrs;
  p.ld = g->w_ld; p.off = g->w_off; p.cout_valid = g->w_cout;
  const int CW = g->Cin % 64 == 0 ? 64 : (g->Cin % 32 == 0 ? 32 : 16);
  p.ncb = g->Cin / CW;
  const int BPT = 128 / CW;
  p.tiles_per_var = ekl_cdiv(g->ntaps * p.ncb, BPT);
  // pixel tiling: KP = tb*th*tw
  int tw = g->mW < KP ? g->mW : KP;
  int rest = KP / tw;
  int th = g->mH < rest ? g->mH : rest;
  int tb = rest / th;
  p.tw = tw; p.th = th; p.tb = tb;
  p.nTw = ekl_cdiv(g->mW, tw); p.nTh = ekl_cdiv(g->mH, th);
  p.ptiles = p.nTw * p.nTh * ekl_cdiv(g->mB, tb);
  p.rows_valid = tb * th * tw;
  const int BNW = g->N % 256 == 0 ? 256 : (g->N % 128 == 0 ? 128 : (g->N % 64 == 0 ? 64 : (g->N % 32 == 0 ? 32 : 16)));
  const int YW = BNW < 64 ? BNW : 64;
  const int a_swz = CW == 64 ? 3 : (CW == 32 ? 2 : 1);
  const int y_swz = YW == 64 ? 3 : (YW == 32 ? 2 : 1);
  for (int i = 0; i < g->n_a; ++i) {
    int rc = make_map(&p.a_maps[i], g->a[i], CW, tw, th, tb, a_swz);
    if (rc) return rc;
  }
  for (int i = 0; i < g->nvar; ++i) {
    int rc = make_map(&p.y_maps[i], g->o[i], YW, tw, th, tb, y_swz);
    if (rc) return rc;
  }
  // split-K: fill the resident-CTA slots of the machine (2 CTAs per SM) WITHOUT spilling into a second wave, and keep
  // >= 8 pixel tiles per CTA (measured: the red.global epilogue and the CTA prologue dominate shorter main loops)
  const int base_ctas = p.tiles_per_var * g->nvar * (g->N / BNW);
  int splits = (2 * ekl_num_sms()) / base_ctas;
  if (splits > p.ptiles / 8) splits = p.ptiles / 8;
  if (splits < 1) splits = 1;
  if (const char* e = getenv("EKL_WG_SPLITS")) {          // experiment knob
    const int v = atoi(e);
    if (v > 0) splits = v < p.ptiles ? v : p.ptiles;
  }
#define EKL_WG_CASE(bn, cw) if (BNW == bn && CW == cw) return launch_wg<bn, cw>(g, p, splits, st);
  EKL_WG_CASE(256, 64) EKL_WG_CASE(128, 64) EKL_WG_CASE(64, 64) EKL_WG_CASE(32, 64)
  EKL_WG_CASE(256, 32) EKL_WG_CASE(128, 32) EKL_WG_CASE(64, 32) EKL_WG_CASE(32, 32)
  EKL_WG_CASE(256, 16) EKL_WG_CASE(128, 16) EKL_WG_CASE(64, 16) EKL_WG_CASE(32, 16)
  EKL_WG_CASE(16, 64) EKL_WG_CASE(16, 32) EKL_WG_CASE(16, 16)
#undef EKL_WG_CASE
  return ekl_fail(-1, "wgrad_tc: no kernel for BNW=%d CW=%d", BNW, CW);
}
