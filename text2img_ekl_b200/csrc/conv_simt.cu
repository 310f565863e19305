// SIMT gather-conv kernels: (1) the small-channel layers that are HBM-bound and have no tensor-core shape
// (first discriminator conv Cin=3 reading the loader's NCHW fp32 images, image heads Cout=3 + tanh writing
// NCHW fp32), and (2) an independent cross-check of the tcgen05 kernels used by the test-suite
// (ekl_conv.impl = 1).  Same plan structure as conv_tc.cu (conv_plan.h); fp32 accumulation.
#include "conv_plan.h"
#include "ekl_common.cuh"

namespace {

struct SimtParams {
  EklGather g;
  const bf16* w;     // packed [nvar][N][ntaps][Cin]
  int act;           // 0 none, 1 leaky-relu(0.2), 2 tanh
};

__device__ __forceinline__ float load_elem(const EklView& v, int b, int h, int w, int c) {
  if ((unsigned)b >= (unsigned)v.dB || (unsigned)h >= (unsigned)v.dH || (unsigned)w >= (unsigned)v.dW) return 0.f;
  const int64_t off = b * v.sB + h * v.sH + w * v.sW + c * v.sC;
  return v.f32 ? reinterpret_cast<const float*>(v.base)[off] : __bfloat162float(reinterpret_cast<const bf16*>(v.base)[off]);
}

// block = 128 threads: 4 warps, each warp one pixel at a time; lanes stride over output channels.
// A-row of the current (pixel, tap) is staged in smem as fp32 and broadcast to all lanes.
__global__ void __launch_bounds__(128) conv_gather_simt_kernel(const __grid_constant__ SimtParams p) {
  extern __shared__ float arow[];          // [4 warps][Cin]
  const EklGather& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.y;
  const int64_t npix = (int64_t)g.mB * g.mH * g.mW;
  float* my = arow + warp * g.Cin;
  const int Ktot = g.ntaps * g.Cin;
  for (int64_t pix = (int64_t)blockIdx.x * 4 + warp; pix < npix; pix += (int64_t)gridDim.x * 4) {
    const int w_ = (int)(pix % g.mW);
    const int h_ = (int)((pix / g.mW) % g.mH);
    const int b_ = (int)(pix / ((int64_t)g.mW * g.mH));
    for (int n0 = 0; n0 < g.N; n0 += 128) {   // up to 4 output channels per lane per pass
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < g.ntaps; ++t) {
        const EklTap tap = g.taps[v][t];
        const EklView& av = g.a[tap.map];
        __syncwarp();
        for (int c = lane; c < g.Cin; c += 32) my[c] = load_elem(av, b_, h_ + tap.dh, w_ + tap.dw, c);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + lane + 32 * j;
          if (n < g.N) {
            const bf16* wr = p.w + ((int64_t)(v * g.N + n)) * Ktot + (int64_t)t * g.Cin;
            float a = 0.f;
            for (int c = 0; c < g.Cin; ++c) a += my[c] * __bfloat162float(wr[c]);
            acc[j] += a;
          }
        }
      }
      const EklView& ov = g.o[v];
      if (b_ < ov.dB && h_ < ov.dH && w_ < ov.dW) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + lane + 32 * j;
          if (n < g.N) {
            float r = acc[j];
            if (p.act == 1) r = r > 0.f ? r : 0.2f * r;
            else if (p.act == 2) r = tanhf(r);
            const int64_t off = b_ * ov.sB + h_ * ov.sH + w_ * ov.sW + n * ov.sC;
            if (ov.f32) reinterpret_cast<float*>(ov.base)[off] = r;
            else reinterpret_cast<bf16*>(ov.base)[off] = __float2bfloat16(r);
          }
        }
      }
    }
  }
}

// Weight gradient, SIMT: one block per (variant, tap, cout); threads over cin; loop over all pixels.
//   dWm[co][src(kh,kw)][ci] += sum_p dY_v[p, co] * A_{map}[p + (dh,dw), ci]      for every master tap src of (v,t)
// `g` is the FORWARD plan (A = x views, o = y-shaped views holding dy).
struct WgSimtParams {
  EklGather g;
  float* dw;        // master layout [Cout][KH*KW][Cin] fp32, accumulated with atomics
};

__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const __grid_constant__ WgSimtParams p) {
  const EklGather& g = p.g;
  const int co = blockIdx.x, t = blockIdx.y, v = blockIdx.z;
  const EklTap tap = g.taps[v][t];
  const EklView& av = g.a[tap.map];
  const EklView& yv = g.o[v];
  const int64_t npix = (int64_t)g.mB * g.mH * g.mW;
  const int KK = g.KH * g.KW;
  // split threads: ci lanes x pixel slices
  const int nci = g.Cin < 256 ? g.Cin : 256;
  const int slices = 256 / nci > 0 ? 256 / nci : 1;
  const int ci0 = threadIdx.x % nci, sl = threadIdx.x / nci;
  if (sl >= slices) return;
  for (int ci = ci0; ci < g.Cin; ci += nci) {
    float acc = 0.f;
    for (int64_t pix = sl; pix < npix; pix += slices) {
      const int w_ = (int)(pix % g.mW);
      const int h_ = (int)((pix / g.mW) % g.mH);
      const int b_ = (int)(pix / ((int64_t)g.mW * g.mH));
      const float dy = load_elem(yv, b_, h_, w_, co);
      if (dy != 0.f) acc += dy * load_elem(av, b_, h_ + tap.dh, w_ + tap.dw, ci);
    }
    if (co < g.w_cout)
      for (int s = 0; s < tap.nsrc; ++s) atomicAdd(p.dw + EKL_WIDX(g.w_kcrs, co, tap.src[s], g.w_off + ci, KK, g.w_ld), acc);
  }
}

}  // namespace

int ekl_gather_simt(const EklGather* g, const void* w_packed, int act, cudaStream_t st) {
  SimtParams p;
  p.g = *g; p.w = (const bf16*)w_packed; p.act = act;
  const int64_t npix = (int64_t)g->mB * g->mH * g->mW;
  int blocks = (int)((npix + 3) / 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  dim3 grid(blocks, g->nvar);
  const size_t smem = (size_t)4 * g->Cin * sizeof(float);
  EKL_REQUIRE(smem <= 48 * 1024, "gather_simt: Cin %d too large", g->Cin);
  conv_gather_simt_kernel<<<grid, 128, smem, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}

int ekl_wgrad_simt(const EklGather* fwd_plan, float* dw_master, cudaStream_t st) {
  WgSimtParams p;
  p.g = *fwd_plan; p.dw = dw_master;
  dim3 grid(fwd_plan->N, fwd_plan->ntaps, fwd_plan->nvar);
  conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(p);
  EKL_LAUNCH_CHECK();
  return 0;
}
