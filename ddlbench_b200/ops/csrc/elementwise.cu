// Average pooling (fixed window and adaptive) and dropout for NHWC bf16 activations — the last
// element-wise layers of the model zoo that still went through ATen (VGG's classifier dropout,
// Inception / SqueezeNet / DenseNet average pools).
//
// (reference: nn.Dropout / nn.AvgPool2d / nn.AdaptiveAvgPool2d through ATen,
//  benchmark/imagenet/gpipemodels/vgg/vgg.py:28-46.)
//
// Dropout draws its bits from Philox4x32-10 keyed by (seed, per-call salt) with the counter
// (element index / 8, device step counter): the step counter lives in device memory and is bumped
// once per training step INSIDE the captured graph, so a replayed graph sees fresh masks, and a
// recomputed forward (GPipe checkpointing: same step, same salt) sees the SAME mask.  The mask is
// kept as one bit per element for the backward.
#include "common.cuh"
#include "api.h"

namespace ddl {

// ---------------------------------------------------------------------------------- avg pool
// adaptive == 1: window of output (ho, wo) = [floor(ho*H/Ho), ceil((ho+1)*H/Ho)) (PyTorch's rule)
__device__ __forceinline__ void pool_window(int o, int in, int out, int k, int s, int p, int adaptive,
                                            int& lo, int& hi) {
  if (adaptive) {
    lo = (o * in) / out;
    hi = ((o + 1) * in + out - 1) / out;
  } else {
    lo = o * s - p;
    hi = lo + k;
  }
}

__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                   int W, int C, int Ho, int Wo, int k, int s, int p, int adaptive, int count_pad) {
  pdl_enter();
  const int cv_total = C >> 3;
  const int row_items = Wo * cv_total;
  for (int rowid = blockIdx.y; rowid < N * Ho; rowid += gridDim.y) {
    const int n = rowid / Ho, ho = rowid - n * Ho;
    int h0, h1;
    pool_window(ho, H, Ho, k, s, p, adaptive, h0, h1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_items; i += gridDim.x * blockDim.x) {
      const int wo = i / cv_total, cv = i - wo * cv_total;
      int w0, w1;
      pool_window(wo, W, Wo, k, s, p, adaptive, w0, w1);
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      int cnt = 0;
      for (int h = h0; h < h1; ++h) {
        if (h < 0 || h >= H) continue;
        for (int w = w0; w < w1; ++w) {
          if (w < 0 || w >= W) continue;
          int4 raw = ld_stream(x + (((int64_t)n * H + h) * W + w) * C + cv * 8);
          float f[8];
          unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] += f[c];
          ++cnt;
        }
      }
      // count_include_pad counts the zero padding, but (like ATen) never positions beyond H+p / W+p
      const int hh = min(h1, H + p) - h0, ww = min(w1, W + p) - w0;
      const float inv = 1.f / (float)((count_pad && !adaptive) ? hh * ww : max(cnt, 1));
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] *= inv;
      *reinterpret_cast<bf16x8*>(y + (((int64_t)n * Ho + ho) * Wo + wo) * C + cv * 8) = pack8(acc);
    }
  }
}

__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int H,
                   int W, int C, int Ho, int Wo, int k, int s, int p, int adaptive, int count_pad) {
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
      // gather over the output windows that contain (h, w): deterministic, no atomics
      int ho_lo, ho_hi, wo_lo, wo_hi;
      if (adaptive) {
        ho_lo = 0; ho_hi = Ho - 1; wo_lo = 0; wo_hi = Wo - 1;      // windows may overlap irregularly
      } else {
        ho_lo = h + p - k + 1; ho_lo = ho_lo > 0 ? (ho_lo + s - 1) / s : 0;
        ho_hi = min((h + p) / s, Ho - 1);
        wo_lo = w + p - k + 1; wo_lo = wo_lo > 0 ? (wo_lo + s - 1) / s : 0;
        wo_hi = min((w + p) / s, Wo - 1);
      }
      for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        int h0, h1;
        pool_window(ho, H, Ho, k, s, p, adaptive, h0, h1);
        if (h < h0 || h >= h1) continue;
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          int w0, w1;
          pool_window(wo, W, Wo, k, s, p, adaptive, w0, w1);
          if (w < w0 || w >= w1) continue;
          int cnt;
          if (count_pad && !adaptive) {
            cnt = (min(h1, H + p) - h0) * (min(w1, W + p) - w0);
          } else {
            cnt = (min(h1, H) - max(h0, 0)) * (min(w1, W) - max(w0, 0));
          }
          int4 raw = ld_stream(dy + (((int64_t)n * Ho + ho) * Wo + wo) * C + cv * 8);
          float f[8];
          unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
          const float inv = 1.f / (float)max(cnt, 1);
#pragma unroll
          for (int c = 0; c < 8; ++c) g[c] = fmaf(f[c], inv, g[c]);
        }
      }
      *reinterpret_cast<bf16x8*>(dx + (((int64_t)n * H + h) * W + w) * C + cv * 8) = pack8(g);
    }
  }
}

// ---------------------------------------------------------------------------------- dropout
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t* out) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// y = x * keep / (1 - p); mask bit = keep.  One thread per 8 elements (one 16-byte vector, one mask byte).
__global__ void __launch_bounds__(256)
dropout_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                   uint8_t* __restrict__ mask, int64_t nvec, float p, uint32_t seed_lo, uint32_t seed_hi,
                   uint32_t salt_lo, uint32_t salt_hi, const long long* __restrict__ step) {
  pdl_enter();
  const unsigned long long st = step ? (unsigned long long)*step : 0ull;
  const uint32_t thresh = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);   // drop if r < thresh
  const float scale = 1.f / (1.f - p);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[8];
    philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)st, (uint32_t)(st >> 32) ^ salt_hi,
                  seed_lo ^ salt_lo, seed_hi, r);
    philox4x32_10((uint32_t)v, (uint32_t)(v >> 32) | 0x80000000u, (uint32_t)st, (uint32_t)(st >> 32) ^ salt_hi,
                  seed_lo ^ salt_lo, seed_hi, r + 4);
    int4 raw = ld_stream(x + v * 8);
    float f[8];
    unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
    uint8_t m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const bool keep = r[c] >= thresh;
      m |= (uint8_t)keep << c;
      f[c] = keep ? f[c] * scale : 0.f;
    }
    *reinterpret_cast<bf16x8*>(y + v * 8) = pack8(f);
    mask[v] = m;
  }
}

__global__ void __launch_bounds__(256)
dropout_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ mask,
                   __nv_bfloat16* __restrict__ dx, int64_t nvec, float scale) {
  pdl_enter();
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    int4 raw = ld_stream(dy + v * 8);
    float f[8];
    unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
    const uint8_t m = mask[v];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = ((m >> c) & 1) ? f[c] * scale : 0.f;
    *reinterpret_cast<bf16x8*>(dx + v * 8) = pack8(f);
  }
}

}  // namespace ddl

using namespace ddl;

extern "C" int ddl_avgpool(int backward, const void* in, void* out, int N, int H, int W, int C, int Ho,
                           int Wo, int k, int s, int p, int adaptive, int count_include_pad, int sms,
                           cudaStream_t st) {
  (void)sms;
  if (C % 8) return -1;
  if (!adaptive && (k < 1 || s < 1 || p < 0 || p > k / 2)) return -1;
  if (!backward) {
    const int row_items = Wo * (C / 8);
    dim3 grid((row_items + 255) / 256, N * Ho > 65535 ? 65535 : N * Ho);
    launch_pdl(avgpool_fwd_kernel, grid, dim3(256), 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, N,
               H, W, C, Ho, Wo, k, s, p, adaptive, count_include_pad);
  } else {
    const int row_items = W * (C / 8);
    dim3 grid((row_items + 255) / 256, N * H > 65535 ? 65535 : N * H);
    launch_pdl(avgpool_bwd_kernel, grid, dim3(256), 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, N,
               H, W, C, Ho, Wo, k, s, p, adaptive, count_include_pad);
  }
  return (int)cudaGetLastError();
}

extern "C" int ddl_dropout_fwd(const void* x, void* y, void* mask, int64_t n, float p, uint64_t seed,
                               uint64_t salt, const long long* step, int sms, cudaStream_t st) {
  if (n % 8 || p < 0.f || p >= 1.f) return -1;
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  launch_pdl(dropout_fwd_kernel, dim3((unsigned)(blocks < 1 ? 1 : blocks)), dim3(256), 0, st,
             (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint8_t*)mask, nvec, p, (uint32_t)seed,
             (uint32_t)(seed >> 32), (uint32_t)salt, (uint32_t)(salt >> 32), step);
  return (int)cudaGetLastError();
}

extern "C" int ddl_dropout_bwd(const void* dy, const void* mask, void* dx, int64_t n, float scale, int sms,
                               cudaStream_t st) {
  if (n % 8) return -1;
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  launch_pdl(dropout_bwd_kernel, dim3((unsigned)(blocks < 1 ? 1 : blocks)), dim3(256), 0, st,
             (const __nv_bfloat16*)dy, (const uint8_t*)mask, (__nv_bfloat16*)dx, nvec, scale);
  return (int)cudaGetLastError();
}
