// Training-mode BatchNorm for NHWC (channels-last) bf16 activations, fused with the
// ReLU/ReLU6 and residual-add that follow it in every ResNet / MobileNetV2 block.
//
// Reference: the models run BatchNorm2d, ReLU and the residual `out += shortcut(x)` as
// three separate library kernels each way (benchmark/cifar10/pytorchcifargitmodels/
// resnet.py:33-39,61-67; benchmark/imagenet/gpipemodels/resnet/block.py:31-79), i.e.
// 8 full passes over the activation forward.  Here:
//
//   forward : bn_stats  (1 read; the LAST CTA to finish turns the partial sums into
//                        mean/invstd, updates the running statistics and re-zeroes the
//                        workspace — no separate finalize launch, no memset)
//             bn_apply  (read x [+res], write y = act(BN(x) [+ res]))
//   backward: bn_bwd_reduce (read dy[,dy2],y,x [write dz]; last CTA produces the per-channel
//                        coefficients and writes dgamma/dbeta into the flat gradient buffer)
//             bn_bwd_apply  (read dz|dy+y, x; write dx)
//
// Partial sums are accumulated with fp32 atomics into R REPLICATED copies of the [2][C]
// workspace (CTA b uses replica b % R): same-address L2 atomics from hundreds of CTAs
// serialise (measured: the single-copy version spent 58 us/launch, 32% of the training step,
// in that serialisation), replication divides the contention by R.
// When the convolution that produces x is one of our tcgen05 GEMM kernels, bn_stats is
// skipped: the GEMM epilogue accumulates the same sums and bn_finalize closes them.
//
// Thread mapping: blockDim = (CVb, RL) with CVb = min(C/8, 256) 16-byte channel vectors
// and RL row lanes, so a warp always touches one contiguous span of memory.
#include "common.cuh"
#include "api.h"

namespace ddl {

constexpr int kBnThreads = 256;
constexpr int kRep = DDL_BN_REPLICAS;  // replicas of the partial-sum workspace

struct BnDims {
  int cvb;   // channel vectors per block (blockDim.x)
  int rl;    // row lanes (blockDim.y)
  int gy;    // channel chunks (gridDim.y)
  int gx;    // row blocks (gridDim.x)
  int rows_per_block;
};

static inline BnDims bn_dims(int64_t M, int C, int sms, int waves, int min_rows_per_lane) {
  BnDims d;
  const int cv = C / 8;
  d.cvb = cv < kBnThreads ? cv : kBnThreads;
  d.rl = kBnThreads / d.cvb;
  if (d.rl < 1) d.rl = 1;
  d.gy = (cv + d.cvb - 1) / d.cvb;
  int64_t target = (int64_t)sms * waves / d.gy;
  if (target < 1) target = 1;
  int64_t max_useful = (M + (int64_t)d.rl * min_rows_per_lane - 1) / ((int64_t)d.rl * min_rows_per_lane);
  if (target > max_useful) target = max_useful;
  if (target < 1) target = 1;
  d.gx = (int)target;
  d.rows_per_block = (int)((M + d.gx - 1) / d.gx);
  d.gx = (int)((M + d.rows_per_block - 1) / d.rows_per_block);
  return d;
}

// Mapping for the REDUCING kernels (statistics, backward sums, bias gradient): a block is one
// warp of channel vectors wide (<= 256 channels) and 8 row lanes deep, the grid is
// (row splits, channel slabs).  The number of partial-sum atomics is gx*gy*cvb*NV/4; with the
// full-width mapping above a 2048-channel layer issued 1.6 M scalar atomics for 6 MB of data
// (ncu: 31 us, 0.4 TB/s), this one 75 K vector atomics.
static inline BnDims bn_dims_reduce(int64_t M, int C, int sms, int waves, int min_rows_per_lane) {
  BnDims d;
  const int cv = C / 8;
  d.cvb = cv < 32 ? cv : 32;
  d.rl = kBnThreads / d.cvb;
  d.gy = (cv + d.cvb - 1) / d.cvb;
  int64_t target = (int64_t)sms * waves / d.gy;
  if (target < 1) target = 1;
  int64_t max_useful = (M + (int64_t)d.rl * min_rows_per_lane - 1) / ((int64_t)d.rl * min_rows_per_lane);
  if (target > max_useful) target = max_useful;
  if (target < 1) target = 1;
  d.gx = (int)target;
  d.rows_per_block = (int)((M + d.gx - 1) / d.gx);
  d.gx = (int)((M + d.rows_per_block - 1) / d.rows_per_block);
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

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) {
  int4 raw = ld_stream(p);
  unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
}

// The activation derivative as ONE BIT per element (bit k of byte [row][c/8] = act'(y[c+k]) != 0),
// written by bn_apply and read by the two backward kernels instead of the 16x larger y: the
// backward is bandwidth-bound on the big layers (ncu: 5.5 TB/s), so bytes are time.  The bit is
// derived from the ROUNDED bf16 output so it equals act_mask(y) exactly.
__device__ __forceinline__ uint8_t act_bits(const bf16x8& p, int act) {
  float q[8];
  unpack8(p, q);
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) m |= (act_mask(q[k], act) != 0.f ? 1u : 0u) << k;
  return (uint8_t)m;
}
constexpr int kActBitmask = 0x100;  // flag in `act`: the `y` argument points to such a bitmask

// block-level reduction of NV (8 or 16) per-thread values over the row lanes, then one
// atomicAdd per (value, channel) into replica `rep` of the workspace.
template <int NV>
__device__ __forceinline__ void block_reduce_to_workspace(const float* vals, float (*red)[kBnThreads],
                                                          float* sums, int C, int cv_total,
                                                          int rep) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < NV; ++k) red[k][tid] = vals[k];
  __syncthreads();
  const int nthreads = blockDim.x * blockDim.y;
  const int nout = blockDim.x * (NV / 4);
  float* dst = sums + (size_t)rep * 2 * C;
  // one 16-byte vector atomic (red.global.add.v4.f32, sm_90+) per 4 adjacent channels:
  // a quarter of the L2 atomic transactions of the scalar form
  for (int o = tid; o < nout; o += nthreads) {
    const int tx = o % blockDim.x, k0 = (o / blockDim.x) * 4;
    const int c_vec = blockIdx.y * blockDim.x + tx;
    if (c_vec >= cv_total) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ty = 0; ty < blockDim.y; ++ty) {
      const int t = ty * blockDim.x + tx;
      acc.x += red[k0][t];
      acc.y += red[k0 + 1][t];
      acc.z += red[k0 + 2][t];
      acc.w += red[k0 + 3][t];
    }
    atomicAdd(reinterpret_cast<float4*>(&dst[(k0 >> 3) * C + c_vec * 8 + (k0 & 7)]), acc);
  }
}

// true in every thread of the LAST CTA (of this channel chunk) to get here
__device__ __forceinline__ bool last_cta_of_chunk(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int t = atomicAdd(&counter[blockIdx.y], 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) counter[blockIdx.y] = 0u;  // self-reset for the next launch / graph replay
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// sum the R replicas of channel c (both rows) and zero them
__device__ __forceinline__ void gather_and_clear(float* sums, int C, int c, float& s0, float& s1) {
  s0 = 0.f;
  s1 = 0.f;
#pragma unroll
  for (int r = 0; r < kRep; ++r) {
    float* p = sums + (size_t)r * 2 * C;
    s0 += __ldcg(p + c);
    s1 += __ldcg(p + C + c);
    p[c] = 0.f;
    p[C + c] = 0.f;
  }
}

__device__ __forceinline__ void finalize_stats(float s, float q, int c, int C, int64_t M, float eps,
                                               float momentum, float* stat_out, float* rmean,
                                               float* rvar) {
  const float inv_m = 1.f / (float)M;
  const float mean = s * inv_m;
  float var = fmaf(-mean, mean, q * inv_m);
  var = fmaxf(var, 0.f);
  stat_out[c] = mean;
  stat_out[C + c] = rsqrtf(var + eps);
  if (rmean) {
    const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
    rmean[c] = fmaf(momentum, mean - rmean[c], rmean[c]);
    rvar[c] = fmaf(momentum, unbiased - rvar[c], rvar[c]);
  }
}

// --------------------------------------------------------------------------- forward
// sums: [R][2][C] fp32 (sum, sumsq), zero on entry, zero again on exit.
__global__ void __launch_bounds__(kBnThreads, 4)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums,
                unsigned int* __restrict__ counter, float* __restrict__ stat_out,
                float* __restrict__ running_mean, float* __restrict__ running_var, int64_t M,
                int C, int rows_per_block, float eps, float momentum) {
  pdl_enter();
  __shared__ float red[16][kBnThreads];
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  if (cv < cv_total) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    const __nv_bfloat16* base = x + (int64_t)cv * 8;
    const int step = blockDim.y;
    int64_t r = r0 + threadIdx.y;
    for (; r + 3 * step < r1; r += 4 * step) {  // 4 independent 16-byte loads in flight
      float f0[8], f1[8], f2[8], f3[8];
      load8(base + r * C, f0);
      load8(base + (r + step) * C, f1);
      load8(base + (r + 2 * step) * C, f2);
      load8(base + (r + 3 * step) * C, f3);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] += (f0[k] + f1[k]) + (f2[k] + f3[k]);
        v[8 + k] = fmaf(f0[k], f0[k], v[8 + k]);
        v[8 + k] = fmaf(f1[k], f1[k], v[8 + k]);
        v[8 + k] = fmaf(f2[k], f2[k], v[8 + k]);
        v[8 + k] = fmaf(f3[k], f3[k], v[8 + k]);
      }
    }
    for (; r < r1; r += step) {
      float f[8];
      load8(base + r * C, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] += f[k];
        v[8 + k] = fmaf(f[k], f[k], v[8 + k]);
      }
    }
  }
  block_reduce_to_workspace<16>(v, red, sums, C, cv_total, blockIdx.x % kRep);
  if (last_cta_of_chunk(counter)) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * blockDim.x * 8;
    const int nch = min(blockDim.x * 8, C - c0);
    for (int i = tid; i < nch; i += blockDim.x * blockDim.y) {
      float s, q;
      gather_and_clear(sums, C, c0 + i, s, q);
      finalize_stats(s, q, c0 + i, C, M, eps, momentum, stat_out, running_mean, running_var);
    }
  }
}

// Stand-alone finalize for statistics produced by the GEMM epilogue.
__global__ void bn_finalize_kernel(float* __restrict__ sums, float* __restrict__ stat_out,
                                   float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t M, int C, float eps,
                                   float momentum) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s, q;
  gather_and_clear(sums, C, c, s, q);
  finalize_stats(s, q, c, C, M, eps, momentum, stat_out, running_mean, running_var);
}

// y = act((x - mean) * invstd * gamma + beta [+ res])
//
// `sums` != null ("apply-side finalize"): (mean, invstd) are not read from `stat` but computed here
// from the R replicated per-channel sum / sum-of-squares buffers the producing convolution's
// epilogue accumulated (sums[r][0][c], sums[r][1][c]) — every CTA needs 16 L2 loads per channel it
// owns.  The convolution then ends with fire-and-forget reductions (no fence, no ticket, no
// last-CTA finalize tail on the critical path, no finalize launch).  The row-block-0 CTAs write
// (mean, invstd) to `stat_out` for the backward pass and update the running statistics; the LAST
// CTA of a channel block to have read the sums (ticket taken right after its loads, checked after
// its row loop: the atomic's round trip hides behind the streaming work) re-zeroes them.
template <typename PT>
__global__ void __launch_bounds__(kBnThreads, 4)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                __nv_bfloat16* __restrict__ y, const float* __restrict__ stat,
                const PT* __restrict__ gamma, const PT* __restrict__ beta,
                uint8_t* __restrict__ mask_out, int64_t M, int C, int rows_per_block, int act,
                float* __restrict__ sums, float* __restrict__ stat_out,
                float* __restrict__ running_mean, float* __restrict__ running_var,
                unsigned int* __restrict__ ticket, float eps, float momentum) {
  pdl_enter();
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= cv_total) return;        // (never taken in sums mode: the host requires full blocks)
  float sc[8], sh[8];
  unsigned int arrived = 0;
  if (sums) {
    // one pass over the CTA's channels (blockDim.x * 8 of them): 16 L2 loads per channel per CTA,
    // coefficients parked in shared memory for the row lanes
    __shared__ float s_sc[kBnThreads * 8], s_sh[kBnThreads * 8];
    const float inv_m = 1.f / (float)M;
    const int nch = blockDim.x * 8, ch0 = blockIdx.y * nch;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int i = tid; i < nch; i += nthr) {
      const int c = ch0 + i;
      float s_ = 0.f, q_ = 0.f;
#pragma unroll
      for (int r = 0; r < DDL_BN_REPLICAS; ++r) {
        s_ += __ldcg(sums + (size_t)(2 * r) * C + c);
        q_ += __ldcg(sums + (size_t)(2 * r + 1) * C + c);
      }
      const float mean = s_ * inv_m;
      const float var = fmaxf(fmaf(-mean, mean, q_ * inv_m), 0.f);
      const float istd = rsqrtf(var + eps);
      if (blockIdx.x == 0) {
        stat_out[c] = mean;
        stat_out[C + c] = istd;
        if (running_mean) {
          const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
          running_mean[c] = fmaf(momentum, mean - running_mean[c], running_mean[c]);
          running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
        }
      }
      const float g = to_f(gamma[c]), b = to_f(beta[c]);
      s_sc[i] = g * istd;
      s_sh[i] = fmaf(-mean, g * istd, b);
    }
    __syncthreads();                 // the CTA has read its sums
    if (tid == 0) {
      __threadfence();
      arrived = atomicAdd(ticket + blockIdx.y, 1u);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = s_sc[threadIdx.x * 8 + k];
      sh[k] = s_sh[threadIdx.x * 8 + k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cv * 8 + k;
      const float g = to_f(gamma[c]), b = to_f(beta[c]);
      sc[k] = g * stat[C + c];
      sh[k] = fmaf(-stat[c], sc[k], b);
    }
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  const int step = blockDim.y;
  int64_t r = r0 + threadIdx.y;
  for (; r + step < r1; r += 2 * step) {  // two rows in flight
    const int64_t o0 = r * C + (int64_t)cv * 8, o1 = (r + step) * C + (int64_t)cv * 8;
    float f0[8], f1[8];
    load8(x + o0, f0);
    load8(x + o1, f1);
    if (res) {
      float g0[8], g1[8];
      load8(res + o0, g0);
      load8(res + o1, g1);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f0[k] = act_fwd(fmaf(f0[k], sc[k], sh[k]) + g0[k], act);
        f1[k] = act_fwd(fmaf(f1[k], sc[k], sh[k]) + g1[k], act);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f0[k] = act_fwd(fmaf(f0[k], sc[k], sh[k]), act);
        f1[k] = act_fwd(fmaf(f1[k], sc[k], sh[k]), act);
      }
    }
    const bf16x8 p0 = pack8(f0), p1 = pack8(f1);
    *reinterpret_cast<bf16x8*>(y + o0) = p0;  // stays in L2: the next conv reads it
    *reinterpret_cast<bf16x8*>(y + o1) = p1;
    if (mask_out) {
      mask_out[r * cv_total + cv] = act_bits(p0, act);
      mask_out[(r + step) * cv_total + cv] = act_bits(p1, act);
    }
  }
  for (; r < r1; r += step) {
    const int64_t off = r * C + (int64_t)cv * 8;
    float f[8];
    load8(x + off, f);
    if (res) {
      float g[8];
      load8(res + off, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], sh[k]) + g[k], act);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], sh[k]), act);
    }
    const bf16x8 p = pack8(f);
    *reinterpret_cast<bf16x8*>(y + off) = p;
    if (mask_out) mask_out[r * cv_total + cv] = act_bits(p, act);
  }
  if (sums) {
    __shared__ int s_last;
    if (threadIdx.x == 0 && threadIdx.y == 0) s_last = (arrived == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
      // every CTA of this channel block has read the sums: leave them zero for the next forward
      const int ch0 = blockIdx.y * blockDim.x * 8, nch = blockDim.x * 8;
      const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
      for (int i = tid; i < 2 * DDL_BN_REPLICAS * nch; i += nthr)
        sums[(size_t)(i / nch) * C + ch0 + (i % nch)] = 0.f;
      if (tid == 0) ticket[blockIdx.y] = 0u;
    }
  }
}

// -------------------------------------------------------------------------- backward
// sums[0][c] += sum dz ; sums[1][c] += sum dz * xhat ; optionally materialise dz.
// Last CTA: coef[3][C] = {gamma*invstd, sum_dz/M, sum_dz_xhat/M}, dgamma/dbeta written.
template <typename PT, typename GT>
__global__ void __launch_bounds__(kBnThreads, 4)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                     const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                     __nv_bfloat16* __restrict__ dz_out, const float* __restrict__ stat,
                     float* __restrict__ sums, unsigned int* __restrict__ counter,
                     const PT* __restrict__ gamma, float* __restrict__ coef,
                     GT* __restrict__ dgamma, GT* __restrict__ dbeta, int accumulate, int64_t M,
                     int C, int rows_per_block, int act_flags) {
  pdl_enter();
  __shared__ float red[16][kBnThreads];
  const bool bitmask = (act_flags & kActBitmask) != 0;
  const int act = act_flags & 0xff;
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  float v[16], mean[8];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  if (cv < cv_total) {
#pragma unroll
    for (int k = 0; k < 8; ++k) mean[k] = stat[cv * 8 + k];   // invstd is applied once, at finalize
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    const int step = blockDim.y;
    // one row per trip: its 3-4 independent 16-byte loads are issued back to back; occupancy
    // (4 CTAs/SM at <= 64 registers) supplies the rest of the memory-level parallelism
    for (int64_t r = r0 + threadIdx.y; r < r1; r += step) {
      const int64_t o0 = r * C + (int64_t)cv * 8;
      int4 rd = ld_stream(dy + o0);
      int4 rx = *reinterpret_cast<const int4*>(x + o0);  // re-read by bn_bwd_apply: cacheable
      int4 r2, ry;
      unsigned mb = 0;
      if (dy2) r2 = ld_stream(dy2 + o0);
      if (bitmask) mb = reinterpret_cast<const uint8_t*>(y)[r * cv_total + cv];
      else if (act) ry = ld_stream(y + o0);
      float d0[8], x0[8];
      unpack8(*reinterpret_cast<bf16x8*>(&rd), d0);
      if (dy2) {  // gradient fan-in (block output feeds the next conv AND the next skip)
        float e0[8];
        unpack8(*reinterpret_cast<bf16x8*>(&r2), e0);
#pragma unroll
        for (int k = 0; k < 8; ++k) d0[k] += e0[k];
      }
      if (bitmask) {
#pragma unroll
        for (int k = 0; k < 8; ++k) d0[k] = (mb >> k) & 1u ? d0[k] : 0.f;
      } else if (act) {
        float y0[8];
        unpack8(*reinterpret_cast<bf16x8*>(&ry), y0);
#pragma unroll
        for (int k = 0; k < 8; ++k) d0[k] *= act_mask(y0[k], act);
      }
      if (dz_out) *reinterpret_cast<bf16x8*>(dz_out + o0) = pack8(d0);
      unpack8(*reinterpret_cast<bf16x8*>(&rx), x0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] += d0[k];
        v[8 + k] = fmaf(d0[k], x0[k] - mean[k], v[8 + k]);
      }
    }
  }
  block_reduce_to_workspace<16>(v, red, sums, C, cv_total, blockIdx.x % kRep);
  if (last_cta_of_chunk(counter)) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * blockDim.x * 8;
    const int nch = min(blockDim.x * 8, C - c0);
    const float inv_m = 1.f / (float)M;
    for (int i = tid; i < nch; i += blockDim.x * blockDim.y) {
      const int c = c0 + i;
      float sdz, sdzx;
      gather_and_clear(sums, C, c, sdz, sdzx);
      sdzx *= stat[C + c];                       // sum dz * xhat
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
    }
  }
}

// dx = coef0 * (dz - coef1 - xhat * coef2); dz is read from dz_in, or rebuilt from dy & y.
// kMinBlocks = 4 (default, validated): 64 registers, 52 B of spills; kMinBlocks = 3
// (DDLB_BN_BWD_APPLY_BLOCKS=3, for A/B): 85 registers, no spills, 25 % fewer resident warps.
template <int kMinBlocks>
__global__ void __launch_bounds__(kBnThreads, kMinBlocks)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz_in, const __nv_bfloat16* __restrict__ dy,
                    const __nv_bfloat16* __restrict__ dy2, const __nv_bfloat16* __restrict__ y,
                    const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ dx,
                    const float* __restrict__ stat, const float* __restrict__ coef, int64_t M, int C,
                    int rows_per_block, int act_flags) {
  pdl_enter();
  const bool bitmask = (act_flags & kActBitmask) != 0;
  const int act = act_flags & 0xff;
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= cv_total) return;
  // dx = c0*(d - c1 - (x-mean)*istd*c2) = A*d + Bx*x + K  (3 coefficients per channel)
  float A[8], Bx[8], K[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cv * 8 + k;
    const float mean = stat[c], istd = stat[C + c];
    const float c0 = coef[c], c1 = coef[C + c], c2 = coef[2 * C + c];
    A[k] = c0;
    Bx[k] = -c0 * istd * c2;
    K[k] = c0 * (mean * istd * c2 - c1);
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  const int step = blockDim.y;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += 2 * step) {
    const bool two = (r + step) < r1;
    const int64_t o0 = r * C + (int64_t)cv * 8;
    const int64_t o1 = two ? (r + step) * C + (int64_t)cv * 8 : o0;
    float d0[8], d1[8], x0[8], x1[8];
    if (dz_in) {
      load8(dz_in + o0, d0);
      load8(dz_in + o1, d1);
    } else {
      load8(dy + o0, d0);
      load8(dy + o1, d1);
      if (dy2) {
        float e0[8], e1[8];
        load8(dy2 + o0, e0);
        load8(dy2 + o1, e1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          d0[k] += e0[k];
          d1[k] += e1[k];
        }
      }
      if (bitmask) {
        const uint8_t* mk = reinterpret_cast<const uint8_t*>(y);
        const unsigned m0 = mk[r * cv_total + cv];
        const unsigned m1 = mk[(two ? r + step : r) * cv_total + cv];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          d0[k] = (m0 >> k) & 1u ? d0[k] : 0.f;
          d1[k] = (m1 >> k) & 1u ? d1[k] : 0.f;
        }
      } else if (act) {
        float y0[8], y1[8];
        load8(y + o0, y0);
        load8(y + o1, y1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          d0[k] *= act_mask(y0[k], act);
          d1[k] *= act_mask(y1[k], act);
        }
      }
    }
    load8(x + o0, x0);
    load8(x + o1, x1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      d0[k] = fmaf(A[k], d0[k], fmaf(Bx[k], x0[k], K[k]));
      d1[k] = fmaf(A[k], d1[k], fmaf(Bx[k], x1[k], K[k]));
    }
    *reinterpret_cast<bf16x8*>(dx + o0) = pack8(d0);
    if (two) *reinterpret_cast<bf16x8*>(dx + o1) = pack8(d1);
  }
}

// ------------------------------------------------------------------ bias + activation (VGG)
// y = act(x + bias[c]) for NHWC bf16; backward mask is recomputed from y.
template <typename PT>
__global__ void __launch_bounds__(256)
bias_act_kernel(const __nv_bfloat16* __restrict__ x, const PT* __restrict__ bias,
                __nv_bfloat16* __restrict__ y, int64_t nvec, int cv_total, int act) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % cv_total);
    float f[8];
    load8(x + v * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = act_fwd(f[k] + (bias ? to_f(bias[cv * 8 + k]) : 0.f), act);
    *reinterpret_cast<bf16x8*>(y + v * 8) = pack8(f);
  }
}

// dz = (dy [+ dy2]) * mask(y)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
               const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dz, int64_t nvec,
               int act) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float d[8];
    load8(dy + v * 8, d);
    if (dy2) {
      float d2[8];
      load8(dy2 + v * 8, d2);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] += d2[k];
    }
    if (act) {
      float yv[8];
      load8(y + v * 8, yv);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] *= act_mask(yv[k], act);
    }
    *reinterpret_cast<bf16x8*>(dz + v * 8) = pack8(d);
  }
}

// per-channel sum over rows of an NHWC bf16 tensor (bias gradient); last CTA writes `out`.
template <typename GT>
__global__ void __launch_bounds__(kBnThreads, 4)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums,
              unsigned int* __restrict__ counter, GT* __restrict__ out, int accumulate, int64_t M,
              int C, int rows_per_block) {
  pdl_enter();
  __shared__ float red[16][kBnThreads];
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0.f;
  if (cv < cv_total) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      float f[8];
      load8(x + r * C + (int64_t)cv * 8, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += f[k];
    }
  }
  block_reduce_to_workspace<8>(v, red, sums, C, cv_total, blockIdx.x % kRep);
  if (last_cta_of_chunk(counter)) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * blockDim.x * 8;
    const int nch = min(blockDim.x * 8, C - c0);
    for (int i = tid; i < nch; i += blockDim.x * blockDim.y) {
      float s, unused;
      gather_and_clear(sums, C, c0 + i, s, unused);
      out[c0 + i] = accumulate ? from_f<GT>(to_f(out[c0 + i]) + s) : from_f<GT>(s);
    }
  }
}

}  // namespace ddl

using namespace ddl;
typedef __nv_bfloat16 bf16;

extern "C" int ddl_bn_replicas() { return kRep; }

extern "C" int ddl_bn_stats(const void* x, float* sums, unsigned int* counter, float* stat,
                            float* running_mean, float* running_var, int64_t M, int C, float eps,
                            float momentum, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims_reduce(M, C, sms, 4, 4);
  launch_pdl(bn_stats_kernel, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st, 
      (const bf16*)x, sums, counter, stat, running_mean, running_var, M, C, d.rows_per_block, eps,
      momentum);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_finalize(float* sums, float* stat, float* running_mean, float* running_var,
                               int64_t M, int C, float eps, float momentum, cudaStream_t st) {
  launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, st, sums, stat, running_mean, running_var, M, C,
                                                     eps, momentum);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_apply(const void* x, const void* res, void* y, const float* stat,
                            const void* gamma, const void* beta, int param_is_bf16, void* mask_out,
                            int64_t M, int C, int act, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 8, 2);
  if (param_is_bf16)
    launch_pdl(bn_apply_kernel<bf16>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st, 
        (const bf16*)x, (const bf16*)res, (bf16*)y, stat, (const bf16*)gamma, (const bf16*)beta,
        (uint8_t*)mask_out, M, C, d.rows_per_block, act, (float*)nullptr, (float*)nullptr,
        (float*)nullptr, (float*)nullptr, (unsigned int*)nullptr, 0.f, 0.f);
  else
    launch_pdl(bn_apply_kernel<float>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st, 
        (const bf16*)x, (const bf16*)res, (bf16*)y, stat, (const float*)gamma, (const float*)beta,
        (uint8_t*)mask_out, M, C, d.rows_per_block, act, (float*)nullptr, (float*)nullptr,
        (float*)nullptr, (float*)nullptr, (unsigned int*)nullptr, 0.f, 0.f);
  return (int)cudaGetLastError();
}

// Apply-side finalize: statistics from the replicated sums (see bn_apply_kernel).  `ticket`: >= 16
// zero counters (self-resetting).  Returns -1 when the shape does not fit (caller finalizes separately).
extern "C" int ddl_bn_apply_sums(const void* x, const void* res, void* y, float* sums, float* stat_out,
                                 float* running_mean, float* running_var, unsigned int* ticket,
                                 const void* gamma, const void* beta, int param_is_bf16,
                                 void* mask_out, int64_t M, int C, float eps, float momentum, int act,
                                 int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 8, 2);
  if ((C / 8) % d.cvb || d.gy > 16) return -1;       // full channel blocks only
  if (param_is_bf16)
    launch_pdl(bn_apply_kernel<bf16>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st,
        (const bf16*)x, (const bf16*)res, (bf16*)y, (const float*)nullptr, (const bf16*)gamma,
        (const bf16*)beta, (uint8_t*)mask_out, M, C, d.rows_per_block, act, sums, stat_out,
        running_mean, running_var, ticket, eps, momentum);
  else
    launch_pdl(bn_apply_kernel<float>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st,
        (const bf16*)x, (const bf16*)res, (bf16*)y, (const float*)nullptr, (const float*)gamma,
        (const float*)beta, (uint8_t*)mask_out, M, C, d.rows_per_block, act, sums, stat_out,
        running_mean, running_var, ticket, eps, momentum);
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_bwd_reduce(const void* dy, const void* dy2, const void* y, const void* x,
                                 void* dz_out, const float* stat, float* sums,
                                 unsigned int* counter, const void* gamma, int param_is_bf16,
                                 float* coef, void* dgamma, void* dbeta, int grad_is_bf16,
                                 int accumulate, int64_t M, int C, int act, int sms,
                                 cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims_reduce(M, C, sms, 4, 4);
#define L(PT, GT)                                                                              \
  launch_pdl((bn_bwd_reduce_kernel<PT, GT>), dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st,                  \
      (const bf16*)dy, (const bf16*)dy2, (const bf16*)y, (const bf16*)x, (bf16*)dz_out, stat,   \
      sums, counter, (const PT*)gamma, coef, (GT*)dgamma, (GT*)dbeta, accumulate, M, C,         \
      d.rows_per_block, act)
  if (param_is_bf16 && grad_is_bf16) L(bf16, bf16);
  else if (param_is_bf16) L(bf16, float);
  else if (grad_is_bf16) L(float, bf16);
  else L(float, float);
#undef L
  return (int)cudaGetLastError();
}

extern "C" int ddl_bn_bwd_apply(const void* dz_in, const void* dy, const void* dy2, const void* y,
                                const void* x, void* dx, const float* stat, const float* coef,
                                int64_t M, int C, int act, int sms, cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims(M, C, sms, 8, 2);
  static const int min_blocks = [] {
    const char* e = std::getenv("DDLB_BN_BWD_APPLY_BLOCKS");
    return (e && e[0] == '4') ? 4 : 3;   // 3 blocks/SM: 85 registers, no spills (measured faster)
  }();
  if (min_blocks == 3)
    launch_pdl(bn_bwd_apply_kernel<3>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st,
        (const bf16*)dz_in, (const bf16*)dy, (const bf16*)dy2, (const bf16*)y, (const bf16*)x,
        (bf16*)dx, stat, coef, M, C, d.rows_per_block, act);
  else
    launch_pdl(bn_bwd_apply_kernel<4>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st,
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
    launch_pdl(bias_act_kernel<bf16>, dim3((int)blocks), dim3(256), 0, st, (const bf16*)x, (const bf16*)bias, (bf16*)y,
                                                       nvec, C / 8, act);
  else
    launch_pdl(bias_act_kernel<float>, dim3((int)blocks), dim3(256), 0, st, (const bf16*)x, (const float*)bias,
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
  launch_pdl(act_bwd_kernel, dim3((int)blocks), dim3(256), 0, st, (const bf16*)dy, (const bf16*)dy2, (const bf16*)y,
                                              (bf16*)dz, nvec, act);
  return (int)cudaGetLastError();
}

extern "C" int ddl_colsum(const void* x, float* sums, unsigned int* counter, void* out,
                          int out_is_bf16, int accumulate, int64_t M, int C, int sms,
                          cudaStream_t st) {
  if (C % 8) return -1;
  BnDims d = bn_dims_reduce(M, C, sms, 4, 4);
  if (out_is_bf16)
    launch_pdl(colsum_kernel<bf16>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st, 
        (const bf16*)x, sums, counter, (bf16*)out, accumulate, M, C, d.rows_per_block);
  else
    launch_pdl(colsum_kernel<float>, dim3(dim3(d.gx, d.gy)), dim3(dim3(d.cvb, d.rl)), 0, st, 
        (const bf16*)x, sums, counter, (float*)out, accumulate, M, C, d.rows_per_block);
  return (int)cudaGetLastError();
}
