// Weight packing: fp32 master filters [Cout][KH][KW][Cin] (the channels_last storage of the reference's
// [Cout,Cin,KH,KW] parameters) -> bf16 K-major operands of the gather-GEMM plans (conv_plan.h):
//   forward        Wp[v][co][t][ci] = sum_{s in src(v,t)} W[co][s][ci]
//   data-gradient  Wp[v][ci][t][co] = sum_{s in src(v,t)} W[co][s][ci]      (transposed through a smem tile)
// HBM-bound: 4 B read + 2 B written per (variant, tap) element.
#include "conv_plan.h"
#include "ekl_common.cuh"

namespace {

struct PackParams {
  EklTap taps[EKL_MAX_VAR][EKL_MAX_TAPS];
  const float* w;
  bf16* out;
  int nvar, ntaps, Cout, Cin, KK, kcrs;
  int ld, off, cout_valid;   // master row pitch / channel offset / real output channels (EklGather.w_ld, w_off, w_cout)
};

// non-transposed: one thread per output element, ci fastest (coalesced both sides)
__global__ void pack_fwd_kernel(const __grid_constant__ PackParams p) {
  const int64_t total = (int64_t)p.nvar * p.Cout * p.ntaps * p.Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % p.Cin);
    int64_t r = i / p.Cin;
    const int t = (int)(r % p.ntaps); r /= p.ntaps;
    const int co = (int)(r % p.Cout);
    const int v = (int)(r / p.Cout);
    const EklTap tap = p.taps[v][t];
    float acc = 0.f;
    if (co < p.cout_valid)
      for (int s = 0; s < tap.nsrc; ++s) acc += p.w[EKL_WIDX(p.kcrs, co, tap.src[s], p.off + ci, p.KK, p.ld)];
    p.out[i] = __float2bfloat16(acc);
  }
}

// non-transposed, KRSC master, Cin % 8 == 0: one thread per 8 consecutive ci (2 x float4 in, 1 x uint4 out)
__global__ void pack_fwd_vec8_kernel(const __grid_constant__ PackParams p) {
  const int c8 = p.Cin / 8;
  const int64_t total = (int64_t)p.nvar * p.Cout * p.ntaps * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % c8) * 8;
    int64_t r = i / c8;
    const int t = (int)(r % p.ntaps); r /= p.ntaps;
    const int co = (int)(r % p.Cout);
    const int v = (int)(r / p.Cout);
    const EklTap tap = p.taps[v][t];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < (co < p.cout_valid ? tap.nsrc : 0); ++s) {
      const float4* src = reinterpret_cast<const float4*>(p.w + ((int64_t)co * p.KK + tap.src[s]) * p.ld + p.off + ci);
      const float4 a = src[0], b = src[1];
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    uint4 o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
    *reinterpret_cast<uint4*>(p.out + i * 8) = o;
  }
}

// transposed: a block of 256 threads moves a 64(co) x 64(ci) tile of one (v, t): 16-byte reads along ci (all of a
// thread's loads issued before use), transpose through shared memory, 16-byte writes along co.
// Requires KRSC master layout, Cin % 4 == 0 and Cout % 8 == 0 (the generic kernel below covers the rest).
__global__ void __launch_bounds__(256) pack_dgrad_vec_kernel(const __grid_constant__ PackParams p) {
  __shared__ float tile[64][65];
  const int v = blockIdx.z / p.ntaps, t = blockIdx.z % p.ntaps;
  const EklTap tap = p.taps[v][t];
  const int co0 = blockIdx.y * 64, ci0 = blockIdx.x * 64;
  float4 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + k * 256;          // 0..1023 : (row, 16-byte chunk)
    const int r = idx >> 4, c4 = (idx & 15) * 4;
    const int co = co0 + r, ci = ci0 + c4;
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (co < p.cout_valid && ci < p.Cin)
      for (int s = 0; s < tap.nsrc; ++s) {
        const float4 a = *reinterpret_cast<const float4*>(p.w + ((int64_t)co * p.KK + tap.src[s]) * p.ld + p.off + ci);
        acc[k].x += a.x; acc[k].y += a.y; acc[k].z += a.z; acc[k].w += a.w;
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + k * 256;
    const int r = idx >> 4, c4 = (idx & 15) * 4;
    tile[r][c4] = acc[k].x; tile[r][c4 + 1] = acc[k].y; tile[r][c4 + 2] = acc[k].z; tile[r][c4 + 3] = acc[k].w;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * 256;          // 0..511 : (ci row, 8-co chunk)
    const int r = idx >> 3, c8 = (idx & 7) * 8;
    const int ci = ci0 + r, co = co0 + c8;
    if (ci < p.Cin && co < p.Cout) {
      const uint4 o = make_uint4(pack_bf16x2(tile[c8][r], tile[c8 + 1][r]), pack_bf16x2(tile[c8 + 2][r], tile[c8 + 3][r]),
                                 pack_bf16x2(tile[c8 + 4][r], tile[c8 + 5][r]), pack_bf16x2(tile[c8 + 6][r], tile[c8 + 7][r]));
      *reinterpret_cast<uint4*>(p.out + (((int64_t)v * p.Cin + ci) * p.ntaps + t) * p.Cout + co) = o;
    }
  }
}

// transposed, generic: block (32x8) moves a 64(co) x 32(ci) tile for one (v, t)
__global__ void pack_dgrad_kernel(const __grid_constant__ PackParams p) {
  __shared__ float tile[64][33];
  const int v = blockIdx.z / p.ntaps, t = blockIdx.z % p.ntaps;
  const EklTap tap = p.taps[v][t];
  const int co0 = blockIdx.y * 64, ci0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 64; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    float acc = 0.f;
    if (co < p.cout_valid && ci < p.Cin)
      for (int s = 0; s < tap.nsrc; ++s) acc += p.w[EKL_WIDX(p.kcrs, co, tap.src[s], p.off + ci, p.KK, p.ld)];
    tile[r][threadIdx.x] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + 2 * threadIdx.x;
    if (ci < p.Cin && co + 1 < p.Cout) {
      const uint32_t pk = pack_bf16x2(tile[2 * threadIdx.x][r], tile[2 * threadIdx.x + 1][r]);
      *reinterpret_cast<uint32_t*>(p.out + (((int64_t)v * p.Cin + ci) * p.ntaps + t) * p.Cout + co) = pk;
    } else if (ci < p.Cin && co < p.Cout) {
      p.out[(((int64_t)v * p.Cin + ci) * p.ntaps + t) * p.Cout + co] = __float2bfloat16(tile[2 * threadIdx.x][r]);
    }
  }
}

}  // namespace

// g: forward or data-gradient plan of a conv with master filter [Cout][KH*KW][Cin]
int ekl_pack_weights(const EklGather* g, const float* w_master, void* out, int Cout, int Cin, cudaStream_t st) {
  PackParams p;
  memcpy(p.taps, g->taps, sizeof(p.taps));
  p.w = w_master; p.out = (bf16*)out; p.nvar = g->nvar; p.ntaps = g->ntaps; p.Cout = Cout; p.Cin = Cin;
  p.KK = g->KH * g->KW; p.kcrs = g->w_kcrs;
  p.ld = g->w_ld > 0 ? g->w_ld : Cin; p.off = g->w_off; p.cout_valid = g->w_cout > 0 ? g->w_cout : Cout;
  if (!g->transposed) {
    const int64_t total = (int64_t)p.nvar * Cout * p.ntaps * Cin;
    if (!p.kcrs && Cin % 8 == 0 && p.ld % 4 == 0 && p.off % 4 == 0) {
      int blocks = (int)((total / 8 + 255) / 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      pack_fwd_vec8_kernel<<<blocks, 256, 0, st>>>(p);
    } else {
      int blocks = (int)((total + 255) / 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      pack_fwd_kernel<<<blocks, 256, 0, st>>>(p);
    }
  } else if (!p.kcrs && Cin % 4 == 0 && Cout % 8 == 0 && p.ld % 4 == 0 && p.off % 4 == 0) {
    dim3 grid(ekl_cdiv(Cin, 64), ekl_cdiv(Cout, 64), p.nvar * p.ntaps);
    pack_dgrad_vec_kernel<<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(ekl_cdiv(Cin, 32), ekl_cdiv(Cout, 64), p.nvar * p.ntaps);
    pack_dgrad_kernel<<<grid, dim3(32, 8), 0, st>>>(p);
  }
  EKL_LAUNCH_CHECK();
  return 0;
}
