// Max-pool forward/backward for NHWC bf16 activations (ResNet stem 3x3/2, VGG 2x2/2).
//
// Reference: nn.MaxPool2d via ATen/cuDNN (benchmark/imagenet/gpipemodels/resnet/resnet.py:57,
// vgg/vgg.py:23).  Measured on B200 (profiles/launches_*): ATen's NHWC max-pool forward and
// backward cost ~170 us EACH for the 32x64x112x112 stem activation — 5% of the whole training
// step for a 64 MB streaming op.  Here: one thread per 8-channel vector, 16-byte loads, the
// forward records the arg-max position inside the window (1 byte per element) so the backward
// is a gather over the <= 4 windows that cover an input pixel — no atomics, deterministic.
#include "common.cuh"
#include "api.h"

namespace ddl {

__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                   uint8_t* __restrict__ idx, int N, int H, int W, int C, int Ho, int Wo, int k,
                   int s, int p) {
  pdl_enter();
  // blockIdx.y = (n, ho) row of the output; x-dimension = (wo, channel-vector): 32-bit index
  // math only (64-bit div/mod per thread made the first version 4x slower than the copy time)
  const int cv_total = C >> 3;
  const int row_items = Wo * cv_total;
  for (int rowid = blockIdx.y; rowid < N * Ho; rowid += gridDim.y) {
   const int n = rowid / Ho, ho = rowid - n * Ho;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_items; i += gridDim.x * blockDim.x) {
    const int wo = i / cv_total, cv = i - wo * cv_total;
    float best[8];
    uint8_t arg[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      best[c] = -INFINITY;
      arg[c] = 0;
    }
    const int h0 = ho * s - p, w0 = wo * s - p;
    for (int r = 0; r < k; ++r) {
      const int h = h0 + r;
      if (h < 0 || h >= H) continue;
      for (int q = 0; q < k; ++q) {
        const int w = w0 + q;
        if (w < 0 || w >= W) continue;
        int4 raw = ld_stream(x + (((int64_t)n * H + h) * W + w) * C + cv * 8);
        float f[8];
        unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
        const uint8_t pos = (uint8_t)(r * k + q);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (f[c] > best[c] || (f[c] != f[c])) {  // first maximum wins; NaN propagates
            best[c] = f[c];
            arg[c] = pos;
          }
      }
    }
    const int64_t o = (((int64_t)n * Ho + ho) * Wo + wo) * C + cv * 8;
    *reinterpret_cast<bf16x8*>(y + o) = pack8(best);
    if (idx) *reinterpret_cast<uint2*>(idx + o) = *reinterpret_cast<uint2*>(arg);
   }
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo,
                   int k, int s, int p) {
  pdl_enter();
  const int cv_total = C >> 3;
  const int row_items = W * cv_total;
  for (int rowid = blockIdx.y; rowid < N * H; rowid += gridDim.y) {
   const int n = rowid / H, h = rowid - n * H;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_items; i += gridDim.x * blockDim.x) {
    const int w = i / cv_total, cv = i - w * cv_total;
    float g[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) g[c] = 0.f;
    // windows (ho, wo) with ho*s - p <= h <= ho*s - p + k - 1
    int ho_lo = (h + p - k + 1 + s - 1);
    ho_lo = ho_lo > 0 ? ho_lo / s : 0;
    int ho_hi = (h + p) / s;
    if (ho_hi > Ho - 1) ho_hi = Ho - 1;
    int wo_lo = (w + p - k + 1 + s - 1);
    wo_lo = wo_lo > 0 ? wo_lo / s : 0;
    int wo_hi = (w + p) / s;
    if (wo_hi > Wo - 1) wo_hi = Wo - 1;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int r = h - (ho * s - p);
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int q = w - (wo * s - p);
        const uint8_t pos = (uint8_t)(r * k + q);
        const int64_t o = (((int64_t)n * Ho + ho) * Wo + wo) * C + cv * 8;
        uint2 ai = *reinterpret_cast<const uint2*>(idx + o);
        const uint8_t* a = reinterpret_cast<const uint8_t*>(&ai);
        int4 raw = ld_stream(dy + o);
        float f[8];
        unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
#pragma unroll
        for (int c = 0; c < 8; ++c) g[c] += (a[c] == pos) ? f[c] : 0.f;
      }
    }
    *reinterpret_cast<bf16x8*>(dx + (((int64_t)n * H + h) * W + w) * C + cv * 8) = pack8(g);
   }
  }
}

}  // namespace ddl

using namespace ddl;

static inline int pool_blocks(int64_t work, int sms) {
  int64_t b = (work + 255) / 256;
  if (b > (int64_t)sms * 16) b = (int64_t)sms * 16;
  return b < 1 ? 1 : (int)b;
}

extern "C" int ddl_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C,
                               int Ho, int Wo, int k, int s, int p, int sms, cudaStream_t st) {
  if (C % 8 || k < 1 || k > 15) return -1;
  const int row_items = Wo * (C / 8);
  dim3 grid((row_items + 255) / 256, N * Ho > 65535 ? 65535 : N * Ho);
  (void)sms;
  launch_pdl(maxpool_fwd_kernel, dim3(grid), dim3(256), 0, st, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint8_t*)idx, N, H, W, C, Ho, Wo, k, s, p);
  return (int)cudaGetLastError();
}

extern "C" int ddl_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W,
                               int C, int Ho, int Wo, int k, int s, int p, int sms,
                               cudaStream_t st) {
  if (C % 8 || k < 1 || k > 15) return -1;
  const int row_items = W * (C / 8);
  dim3 grid((row_items + 255) / 256, N * H > 65535 ? 65535 : N * H);
  (void)sms;
  launch_pdl(maxpool_bwd_kernel, dim3(grid), dim3(256), 0, st, 
      (const __nv_bfloat16*)dy, (const uint8_t*)idx, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo, k, s,
      p);
  return (int)cudaGetLastError();
}
