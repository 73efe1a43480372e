// Training-mode BatchNorm for NHWC (channels-last) bf16 activations, fused with the
// ReLU/ReLU6 and residual-add that follow it in every ResNet / MobileNetV2 block.
//
// Reference: the models run BatchNorm2d, ReLU and the residual `out += shortcut(x)` as
// three separate library kernels each way (benchmark/cifar10/pytorchcifargitmodels/
// resnet.py:33-39,61-67; benchmark/imagenet/gpipemodels/resnet/block.py:31-79), i.e.
// 8 full passes over the activation forward.  Here:
//
//   forward : bn_stats (1 read)  -> bn_finalize (C elems) -> bn_apply (read x [+res], write y)
//   backward: bn_bwd_reduce (read dy,y,x [write dz]) -> bn_bwd_finalize (C elems)
//             -> bn_bwd_apply (read dz|dy+y, x; write dx)
//
// The per-channel partial sums live in a small fp32 workspace that the finalize kernels
// re-zero after consuming, so no memset is ever launched and the sequence can be replayed
// from a captured CUDA graph.  When the convolution that produces x is one of our tcgen05
// GEMM kernels, bn_stats is skipped: the GEMM epilogue accumulates the same sums.
//
// Thread mapping: blockDim = (CVb, RL) with CVb = min(C/8, 256) 16-byte channel vectors
// and RL row lanes, so a warp always touches one contiguous span of memory.
#include "common.cuh"
#include "api.h"

namespace ddl {

constexpr int kBnThreads = 256;

struct BnDims {
  int cvb;   // channel vectors per block (blockDim.x)
  int rl;    // row lanes (blockDim.y)
  int gy;    // channel chunks (gridDim.y)
  int gx;    // row blocks (gridDim.x)
  int rows_per_block;
};

static inline BnDims bn_dims(int64_t M, int C, int sms, int waves) {
  BnDims d;
  const int cv = C / 8;
  d.cvb = cv < kBnThreads ? cv : kBnThreads;
  d.rl = kBnThreads / d.cvb;
  if (d.rl < 1) d.rl = 1;
  d.gy = (cv + d.cvb - 1) / d.cvb;
  int64_t target = (int64_t)sms * waves / d.gy;
  if (target < 1) target = 1;
  int64_t max_useful = (M + d.rl - 1) / d.rl;  // at least one row per lane
  if (target > max_useful) target = max_useful;
  if (target < 1) target = 1;
  d.gx = (int)target;
  d.rows_per_block = (int)((M + d.gx - 1) / d.gx);
  return d;
}

// act: 0 none, 1 relu, 2 relu6
__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
__device__ __forceinline__ float act_mask(float y, int act) {
  if (act == 1) return y > 0.f ? 1.f : 0.f;
  if (act == 2) return (y > 0.f && y < 6.f) ? 1.f : 0.f;
  return 1.f;
}

// --------------------------------------------------------------------------- forward
// sums: [2][C] fp32 (sum, sumsq), must be zero on entry
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums, int64_t M, int C,
                int rows_per_block) {
  __shared__ float red[16][kBnThreads];
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  if (cv < cv_total) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      int4 raw = ld_stream(x + r * C + (int64_t)cv * 8);
      float f[8];
      unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += f[k];
        q[k] = fmaf(f[k], f[k], q[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[k][tid] = s[k];
    red[8 + k][tid] = q[k];
  }
  __syncthreads();
  const int nthreads = blockDim.x * blockDim.y;
  const int nout = blockDim.x * 16;
  for (int o = tid; o < nout; o += nthreads) {
    const int tx = o % blockDim.x, k = o / blockDim.x;
    const int c_vec = blockIdx.y * blockDim.x + tx;
    if (c_vec >= cv_total) continue;
    float acc = 0.f;
    for (int ty = 0; ty < blockDim.y; ++ty) acc += red[k][ty * blockDim.x + tx];
    const int ch = c_vec * 8 + (k & 7);
    atomicAdd(&sums[(k >> 3) * C + ch], acc);
  }
}

// mean/invstd from the sums; running-stat update; re-zero the sums.
// stat_out: [2][C] (mean, invstd)
__global__ void bn_finalize_kernel(float* __restrict__ sums, float* __restrict__ stat_out,
                                   float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t M, int C, float eps,
                                   float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv_m = 1.f / (float)M;
  const float mean = sums[c] * inv_m;
  float var = fmaf(-mean, mean, sums[C + c] * inv_m);
  var = fmaxf(var, 0.f);
  stat_out[c] = mean;
  stat_out[C + c] = rsqrtf(var + eps);
  if (running_mean) {
    const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
    running_mean[c] = fmaf(momentum, mean - running_mean[c], running_mean[c]);
    running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
  }
  sums[c] = 0.f;
  sums[C + c] = 0.f;
}

// y = act((x - mean) * invstd * gamma + beta [+ res])
template <typename PT>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                __nv_bfloat16* __restrict__ y, const float* __restrict__ stat,
                const PT* __restrict__ gamma, const PT* __restrict__ beta, int64_t M, int C,
                int rows_per_block, int act) {
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= cv_total) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cv * 8 + k;
    const float g = to_f(gamma[c]), b = to_f(beta[c]);
    sc[k] = g * stat[C + c];
    sh[k] = fmaf(-stat[c], sc[k], b);
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const int64_t off = r * C + (int64_t)cv * 8;
    int4 raw = ld_stream(x + off);
    float f[8];
    unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
    if (res) {
      int4 rr = ld_stream(res + off);
      float g[8];
      unpack8(*reinterpret_cast<bf16x8*>(&rr), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], sh[k]) + g[k], act);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], sh[k]), act);
    }
    bf16x8 o = pack8(f);
    *reinterpret_cast<bf16x8*>(y + off) = o;  // keep in L2: the next conv reads it
  }
}

// -------------------------------------------------------------------------- backward
// sums[0][c] += sum dz ; sums[1][c] += sum dz * xhat ; optionally materialise dz
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                     const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ dz_out,
                     const float* __restrict__ stat, float* __restrict__ sums, int64_t M, int C,
                     int rows_per_block, int act) {
  __shared__ float red[16][kBnThreads];
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float s[8], q[8], mean[8], istd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = mean[k] = istd[k] = 0.f;
  if (cv < cv_total) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mean[k] = stat[cv * 8 + k];
      istd[k] = stat[C + cv * 8 + k];
    }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const int64_t off = r * C + (int64_t)cv * 8;
      int4 a = ld_stream(dy + off);
      int4 b = *reinterpret_cast<const int4*>(x + off);  // re-read by bwd_apply: keep cached
      float d[8], xv[8];
      unpack8(*reinterpret_cast<bf16x8*>(&a), d);
      unpack8(*reinterpret_cast<bf16x8*>(&b), xv);
      if (dy2) {  // gradient fan-in (block output feeds the next conv AND the next skip)
        int4 a2 = ld_stream(dy2 + off);
        float d2[8];
        unpack8(*reinterpret_cast<bf16x8*>(&a2), d2);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] += d2[k];
      }
      if (act) {
        int4 c4 = ld_stream(y + off);
        float yv[8];
        unpack8(*reinterpret_cast<bf16x8*>(&c4), yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] *= act_mask(yv[k], act);
      }
      if (dz_out) *reinterpret_cast<bf16x8*>(dz_out + off) = pack8(d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += d[k];
        q[k] = fmaf(d[k], (xv[k] - mean[k]) * istd[k], q[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[k][tid] = s[k];
    red[8 + k][tid] = q[k];
  }
  __syncthreads();
  const int nthreads = blockDim.x * blockDim.y;
  const int nout = blockDim.x * 16;
  for (int o = tid; o < nout; o += nthreads) {
    const int tx = o % blockDim.x, k = o / blockDim.x;
    const int c_vec = blockIdx.y * blockDim.x + tx;
    if (c_vec >= cv_total) continue;
    float acc = 0.f;
    for (int ty = 0; ty < blockDim.y; ++ty) acc += red[k][ty * blockDim.x + tx];
    atomicAdd(&sums[(k >> 3) * C + c_vec * 8 + (k & 7)], acc);
  }
}

// coef: [3][C] = {gamma*invstd, sum_dz/M, sum_dz_xhat/M}; writes dgamma/dbeta; re-zeroes sums
template <typename PT, typename GT>
__global__ void bn_bwd_finalize_kernel(float* __restrict__ sums, const float* __restrict__ stat,
                                       const PT* __restrict__ gamma, float* __restrict__ coef,
                                       GT* __restrict__ dgamma, GT* __restrict__ dbeta,
                                       int accumulate, int64_t M, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sdz = sums[c], sdzx = sums[C + c];
  const float inv_m = 1.f / (float)M;
  coef[c] = to_f(gamma[c]) * stat[C + c];
  coef[C + c] = sdz * inv_m;
  coef[2 * C + c] = sdzx * inv_m;
  if (accumulate) {
    dgamma[c] = from_f<GT>(to_f(dgamma[c]) + sdzx);
    dbeta[c] = from_f<GT>(to_f(dbeta[c]) + sdz);
  } else {
    dgamma[c] = from_f<GT>(sdzx);
    dbeta[c] = from_f<GT>(sdz);
  }
  sums[c] = 0.f;
  sums[C + c] = 0.f;
}

// dx = coef0 * (dz - coef1 - xhat * coef2); dz is read from dz_in, or rebuilt from dy & y
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz_in, const __nv_bfloat16* __restrict__ dy,
                    const __nv_bfloat16* __restrict__ dy2, const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                    __nv_bfloat16* __restrict__ dx, const float* __restrict__ stat,
                    const float* __restrict__ coef, int64_t M, int C, int rows_per_block,
                    int act) {
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= cv_total) return;
  float mean[8], istd[8], c0[8], c1[8], c2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cv * 8 + k;
    mean[k] = stat[c];
    istd[k] = stat[C + c];
    c0[k] = coef[c];
    c1[k] = coef[C + c];
    c2[k] = coef[2 * C + c];
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const int64_t off = r * C + (int64_t)cv * 8;
    float d[8], xv[8];
    if (dz_in) {
      int4 a = ld_stream(dz_in + off);
      unpack8(*reinterpret_cast<bf16x8*>(&a), d);
    } else {
      int4 a = ld_stream(dy + off);
      unpack8(*reinterpret_cast<bf16x8*>(&a), d);
      if (dy2) {
        int4 a2 = ld_stream(dy2 + off);
        float d2[8];
        unpack8(*reinterpret_cast<bf16x8*>(&a2), d2);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] += d2[k];
      }
      if (act) {
        int4 c4 = ld_stream(y + off);
        float yv[8];
        unpack8(*reinterpret_cast<bf16x8*>(&c4), yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] *= act_mask(yv[k], act);
      }
    }
    int4 b = ld_stream(x + off);
    unpack8(*reinterpret_cast<bf16x8*>(&b), xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xhat = (xv[k] - mean[k]) * istd[k];
      d[k] = c0[k] * (d[k] - c1[k] - xhat * c2[k]);
    }
    *reinterpret_cast<bf16x8*>(dx + off) = pack8(d);
  }
}

// ------------------------------------------------------------------ bias + activation (VGG)
// y = act(x + bias[c]) for NHWC bf16; backward mask is recomputed from y.
template <typename PT>
__global__ void __launch_bounds__(256)
bias_act_kernel(const __nv_bfloat16* __restrict__ x, const PT* __restrict__ bias,
                __nv_bfloat16* __restrict__ y, int64_t nvec, int cv_total, int act) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % cv_total);
    int4 raw = ld_stream(x + v * 8);
    float f[8];
    unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = act_fwd(f[k] + (bias ? to_f(bias[cv * 8 + k]) : 0.f), act);
    *reinterpret_cast<bf16x8*>(y + v * 8) = pack8(f);
  }
}

// dz = dy * mask(y) [written in place of dy allowed], plus optional second addend (fan-out sum)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
               const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dz, int64_t nvec,
               int act) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    int4 a = ld_stream(dy + v * 8);
    float d[8];
    unpack8(*reinterpret_cast<bf16x8*>(&a), d);
    if (dy2) {
      int4 a2 = ld_stream(dy2 + v * 8);
      float d2[8];
      unpack8(*reinterpret_cast<bf16x8*>(&a2), d2);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] += d2[k];
    }
    if (act) {
      int4 c4 = ld_stream(y + v * 8);
      float yv[8];
      unpack8(*reinterpret_cast<bf16x8*>(&c4), yv);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] *= act_mask(yv[k], act);
    }
    *reinterpret_cast<bf16x8*>(dz + v * 8) = pack8(d);
  }
}

// per-channel sum over rows of an NHWC bf16 tensor (bias gradient), sums zeroed on entry.
__global__ void __launch_bounds__(kBnThreads)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums, int64_t M, int C,
              int rows_per_block) {
  __shared__ float red[8][kBnThreads];
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  if (cv < cv_total) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      int4 raw = ld_stream(x + r * C + (int64_t)cv * 8);
      float f[8];
      unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[k][tid] = s[k];
  __syncthreads();
  const int nthreads = blockDim.x * blockDim.y;
  const int nout = blockDim.x * 8;
  for (int o = tid; o < nout; o += nthreads) {
    const int tx = o % blockDim.x, k = o / blockDim.x;
    const int c_vec = blockIdx.y * blockDim.x + tx;
    if (c_vec >= cv_total) continue;
    float acc = 0.f;
    for (int ty = 0; ty < blockDim.y; ++ty) acc += red[k][ty * blockDim.x + tx];
    atomicAdd(&sums[c_vec * 8 + k], acc);
  }
}

// out[c] = (GT) sums[c] (or +=); re-zero sums
template <typename GT>
__global__ void colsum_finalize_kernel(float* __restrict__ sums, GT* __restrict__ out,
                                       int accumulate, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float v = sums[c];
  out[c] = accumulate ? from_f<GT>(to_f(out[c]) + v) : from_f<GT>(v);
  sums[c] = 0.f;
}

}  // namespace ddl

using namespace ddl;
typedef __nv_bfloat16 bf16;

extern "C" int ddl_bn_stats(const void* x, float* sums, int64_t M, int C, int sms,
                            cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 4);
  bn_stats_kernel<<<dim3(d.gx, d.gy), dim3(d.cvb, d.rl), 0, st>>>((const bf16*)x, sums, M, C,
                                                                 d.rows_per_block);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_finalize(float* sums, float* stat, float* running_mean, float* running_var,
                               int64_t M, int C, float eps, float momentum, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, stat, running_mean, running_var, M, C,
                                                     eps, momentum);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_apply(const void* x, const void* res, void* y, const float* stat,
                            const void* gamma, const void* beta, int param_is_bf16, int64_t M,
                            int C, int act, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 8);
  if (param_is_bf16)
    bn_apply_kernel<bf16><<<dim3(d.gx, d.gy), dim3(d.cvb, d.rl), 0, st>>>(
        (const bf16*)x, (const bf16*)res, (bf16*)y, stat, (const bf16*)gamma, (const bf16*)beta, M,
        C, d.rows_per_block, act);
  else
    bn_apply_kernel<float><<<dim3(d.gx, d.gy), dim3(d.cvb, d.rl), 0, st>>>(
        (const bf16*)x, (const bf16*)res, (bf16*)y, stat, (const float*)gamma, (const float*)beta,
        M, C, d.rows_per_block, act);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_bwd_reduce(const void* dy, const void* dy2, const void* y, const void* x,
                                 void* dz_out,
                                 const float* stat, float* sums, int64_t M, int C, int act,
                                 int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 4);
  bn_bwd_reduce_kernel<<<dim3(d.gx, d.gy), dim3(d.cvb, d.rl), 0, st>>>(
      (const bf16*)dy, (const bf16*)dy2, (const bf16*)y, (const bf16*)x, (bf16*)dz_out, stat, sums,
      M, C, d.rows_per_block, act);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_bwd_finalize(float* sums, const float* stat, const void* gamma,
                                   int param_is_bf16, float* coef, void* dgamma, void* dbeta,
                                   int grad_is_bf16, int accumulate, int64_t M, int C,
                                   cudaStream_t st) {
  const int g = (C + 127) / 128;
#define L(PT, GT)                                                                           \
  bn_bwd_finalize_kernel<PT, GT><<<g, 128, 0, st>>>(sums, stat, (const PT*)gamma, coef,       \
                                                     (GT*)dgamma, (GT*)dbeta, accumulate, M, C)
  if (param_is_bf16 && grad_is_bf16) L(bf16, bf16);
  else if (param_is_bf16) L(bf16, float);
  else if (grad_is_bf16) L(float, bf16);
  else L(float, float);
#undef L
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_bwd_apply(const void* dz_in, const void* dy, const void* dy2, const void* y,
                                const void* x,
                                void* dx, const float* stat, const float* coef, int64_t M, int C,
                                int act, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 8);
  bn_bwd_apply_kernel<<<dim3(d.gx, d.gy), dim3(d.cvb, d.rl), 0, st>>>(
      (const bf16*)dz_in, (const bf16*)dy, (const bf16*)dy2, (const bf16*)y, (const bf16*)x,
      (bf16*)dx, stat, coef, M, C, d.rows_per_block, act);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bias_act(const void* x, const void* bias, int bias_is_bf16, void* y,
                            int64_t M, int C, int act, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  const int64_t nvec = M * (C / 8);
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  if (bias_is_bf16)
    bias_act_kernel<bf16><<<(int)blocks, 256, 0, st>>>((const bf16*)x, (const bf16*)bias, (bf16*)y,
                                                       nvec, C / 8, act);
  else
    bias_act_kernel<float><<<(int)blocks, 256, 0, st>>>((const bf16*)x, (const float*)bias,
                                                        (bf16*)y, nvec, C / 8, act);
  return (int)cudaGetLastError();
}

extern "C" int ddl_act_bwd(const void* dy, const void* dy2, const void* y, void* dz, int64_t n,
                           int act, int sms, cudaStream_t st) {
  if (n % 8) return -1;
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  act_bwd_kernel<<<(int)blocks, 256, 0, st>>>((const bf16*)dy, (const bf16*)dy2, (const bf16*)y,
                                              (bf16*)dz, nvec, act);
  return (int)cudaGetLastError();
}

extern "C" int ddl_colsum(const void* x, float* sums, void* out, int out_is_bf16, int accumulate,
                          int64_t M, int C, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 4);
  colsum_kernel<<<dim3(d.gx, d.gy), dim3(d.cvb, d.rl), 0, st>>>((const bf16*)x, sums, M, C,
                                                               d.rows_per_block);
  if (out_is_bf16)
    colsum_finalize_kernel<bf16><<<(C + 127) / 128, 128, 0, st>>>(sums, (bf16*)out, accumulate, C);
  else
    colsum_finalize_kernel<float><<<(C + 127) / 128, 128, 0, st>>>(sums, (float*)out, accumulate, C);
  return (int)cudaGetLastError();
}
