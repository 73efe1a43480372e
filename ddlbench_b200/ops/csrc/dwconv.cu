// Depthwise 3x3 convolution (MobileNet / MobileNetV2 / NASNet separable blocks) for NHWC bf16:
// forward, data gradient and weight gradient.  One multiply-add per loaded element — the layer is
// pure memory traffic, so the design goal is one coalesced pass: a thread owns 8 channels
// (one 16-byte vector) of one pixel, the 9 taps of those channels live in registers, neighbouring
// threads own neighbouring channel vectors (a warp reads 512 contiguous bytes per tap).
//
//   forward : y[n,oh,ow,c]  = sum_t  x[n, s*oh+kh-p, s*ow+kw-p, c] * w[t][c]
//   dgrad   : dx[n,ih,iw,c] = sum_t  dy[n, (ih+p-kh)/s, (iw+p-kw)/s, c] * w[t][c]   (when divisible)
//   wgrad   : dw[t][c]      = sum_{n,oh,ow} dy[n,oh,ow,c] * x[n, s*oh+kh-p, s*ow+kw-p, c]
//
// Weights are taken TAP-MAJOR, [9][C] bf16 (the caller transposes PyTorch's [C][1][3][3] once per
// step: 9*C elements).  wgrad accumulates into a zeroed fp32 [9][C] workspace: per-thread
// registers over a strip of pixels, a shared-memory tree over the row lanes of the block, then
// one vector atomic per 4 channels.
// (reference: cuDNN grouped-convolution kernels chosen by PyTorch.)
//
// STATUS: default path for depthwise 3x3 layers (DDLB_DWCONV=0: cuDNN grouped convolution); numerics
// validated on hardware (tests/test_kernels_gpu.py::test_dwconv3x3), speed not yet compared to cuDNN.
#include "api.h"
#include "common.cuh"

namespace ddl {

constexpr int kDwThreads = 256;

struct DwDims {
  int N, H, W, C, OH, OW, stride, pad;
};

__device__ __forceinline__ void dw_load8(const __nv_bfloat16* p, float* f) {
  int4 raw = *reinterpret_cast<const int4*>(p);
  unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
}

__global__ void __launch_bounds__(kDwThreads)
dwconv3x3_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wt,
                     __nv_bfloat16* __restrict__ y, DwDims d) {
  pdl_enter();
  const int cv_total = d.C >> 3;
  const int64_t total = (int64_t)d.N * d.OH * d.OW * cv_total;
  const int64_t stride_t = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride_t) {
    const int cv = (int)(i % cv_total);
    const int64_t pix = i / cv_total;
    const int ow = (int)(pix % d.OW);
    const int oh = (int)((pix / d.OW) % d.OH);
    const int n = (int)(pix / ((int64_t)d.OW * d.OH));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * d.stride - d.pad + kh;
      if (ih < 0 || ih >= d.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * d.stride - d.pad + kw;
        if (iw < 0 || iw >= d.W) continue;
        float xv[8], wv[8];
        dw_load8(x + (((int64_t)n * d.H + ih) * d.W + iw) * d.C + cv * 8, xv);
        dw_load8(wt + (int64_t)(kh * 3 + kw) * d.C + cv * 8, wv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv[k], wv[k], acc[k]);
      }
    }
    *reinterpret_cast<bf16x8*>(y + pix * d.C + cv * 8) = pack8(acc);
  }
}

__global__ void __launch_bounds__(kDwThreads)
dwconv3x3_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ wt,
                       __nv_bfloat16* __restrict__ dx, DwDims d) {
  pdl_enter();
  const int cv_total = d.C >> 3;
  const int64_t total = (int64_t)d.N * d.H * d.W * cv_total;
  const int64_t stride_t = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride_t) {
    const int cv = (int)(i % cv_total);
    const int64_t pix = i / cv_total;
    const int iw = (int)(pix % d.W);
    const int ih = (int)((pix / d.W) % d.H);
    const int n = (int)(pix / ((int64_t)d.W * d.H));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = ih + d.pad - kh;
      if (th < 0 || th % d.stride) continue;
      const int oh = th / d.stride;
      if (oh >= d.OH) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = iw + d.pad - kw;
        if (tw < 0 || tw % d.stride) continue;
        const int ow = tw / d.stride;
        if (ow >= d.OW) continue;
        float gv[8], wv[8];
        dw_load8(dy + (((int64_t)n * d.OH + oh) * d.OW + ow) * d.C + cv * 8, gv);
        dw_load8(wt + (int64_t)(kh * 3 + kw) * d.C + cv * 8, wv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(gv[k], wv[k], acc[k]);
      }
    }
    *reinterpret_cast<bf16x8*>(dx + pix * d.C + cv * 8) = pack8(acc);
  }
}

// block = (cvb channel vectors, rl row lanes); grid = (pixel strips, channel slabs)
__global__ void __launch_bounds__(kDwThreads)
dwconv3x3_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                       float* __restrict__ dw, DwDims d, int pix_per_block) {
  __shared__ float red[kDwThreads][9];     // one tap-row per thread for one channel at a time
  pdl_enter();
  const int cv_total = d.C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  if (cv < cv_total) {
    const int64_t npix = (int64_t)d.N * d.OH * d.OW;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
    int64_t p1 = p0 + pix_per_block;
    if (p1 > npix) p1 = npix;
    for (int64_t pix = p0 + threadIdx.y; pix < p1; pix += blockDim.y) {
      const int ow = (int)(pix % d.OW);
      const int oh = (int)((pix / d.OW) % d.OH);
      const int n = (int)(pix / ((int64_t)d.OW * d.OH));
      float gv[8];
      dw_load8(dy + pix * d.C + cv * 8, gv);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * d.stride - d.pad + kh;
        if (ih < 0 || ih >= d.H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow * d.stride - d.pad + kw;
          if (iw < 0 || iw >= d.W) continue;
          float xv[8];
          dw_load8(x + (((int64_t)n * d.H + ih) * d.W + iw) * d.C + cv * 8, xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[kh * 3 + kw][k] = fmaf(gv[k], xv[k], acc[kh * 3 + kw][k]);
        }
      }
    }
  }
  // reduce over the row lanes, one channel of the vector at a time (keeps shared memory small)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int t = 0; t < 9; ++t) red[tid][t] = acc[t][k];
    __syncthreads();
    if (threadIdx.y == 0 && cv < cv_total) {
      float s[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) s[t] = 0.f;
      for (int ly = 0; ly < blockDim.y; ++ly)
#pragma unroll
        for (int t = 0; t < 9; ++t) s[t] += red[ly * blockDim.x + threadIdx.x][t];
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&dw[(int64_t)t * d.C + cv * 8 + k], s[t]);
    }
    __syncthreads();
  }
}

}  // namespace ddl

using namespace ddl;

static inline int dw_blocks(int64_t work, int sms) {
  int64_t b = (work + kDwThreads - 1) / kDwThreads;
  if (b > (int64_t)sms * 16) b = (int64_t)sms * 16;
  return b < 1 ? 1 : (int)b;
}

// x: [N,H,W,C], wt: [9][C], y: [N,OH,OW,C]; mode 0 = forward, 1 = dgrad (dy -> dx), 2 = wgrad
// (dy, x -> fp32 dw[9][C], accumulated).  Returns -1 for unsupported shapes.
extern "C" int ddl_dwconv3x3(int mode, const void* a, const void* b, void* out, int N, int H, int W,
                             int C, int stride, int pad, int sms, cudaStream_t st) {
  if (C % 8 || (stride != 1 && stride != 2) || pad < 0 || pad > 2) return -1;
  DwDims d;
  d.N = N; d.H = H; d.W = W; d.C = C; d.stride = stride; d.pad = pad;
  d.OH = (H + 2 * pad - 3) / stride + 1;
  d.OW = (W + 2 * pad - 3) / stride + 1;
  if (N <= 0 || d.OH <= 0 || d.OW <= 0) return 0;
  cudaError_t e;
  if (mode == 0) {
    e = launch_pdl(dwconv3x3_fwd_kernel, dim3(dw_blocks((int64_t)N * d.OH * d.OW * (C / 8), sms)),
                   dim3(kDwThreads), 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                   (__nv_bfloat16*)out, d);
  } else if (mode == 1) {
    e = launch_pdl(dwconv3x3_dgrad_kernel, dim3(dw_blocks((int64_t)N * H * W * (C / 8), sms)),
                   dim3(kDwThreads), 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                   (__nv_bfloat16*)out, d);
  } else {
    const int cv = C / 8;
    const int cvb = cv < 32 ? cv : 32;
    const int rl = kDwThreads / cvb;
    const int gy = (cv + cvb - 1) / cvb;
    const int64_t npix = (int64_t)N * d.OH * d.OW;
    int64_t gx = (int64_t)sms * 4 / gy;
    if (gx < 1) gx = 1;
    const int64_t max_useful = (npix + rl - 1) / rl;
    if (gx > max_useful) gx = max_useful;
    const int ppb = (int)((npix + gx - 1) / gx);
    gx = (npix + ppb - 1) / ppb;
    e = launch_pdl(dwconv3x3_wgrad_kernel, dim3((unsigned)gx, gy), dim3(cvb, rl), 0, st,
                   (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (float*)out, d, ppb);
  }
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}
