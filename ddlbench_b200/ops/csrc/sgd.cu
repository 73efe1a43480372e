// Fused multi-tensor SGD(momentum, weight-decay) on FLAT buffers.
//
// The reference runs torch.optim.SGD as an unfused per-tensor loop (161 tensors
// for ResNet-50; benchmark/imagenet/imagenet_pytorch.py:125-127,157) plus a
// separate zero_grad memset per tensor (:155), and PipeDream adds three more
// parameter-sized copies per minibatch for weight stashing
// (pipedream-fork/runtime/optimizer.py:110-116,157-159).
//
// Here all parameters of a stage live in one flat fp32 master buffer with a flat
// fp32 momentum buffer; gradients arrive in one flat (bf16 or fp32) buffer; one
// launch updates everything and writes the new model weights (bf16 or fp32)
// straight into the *version slot* the next forward will read (w_out) — weight
// stashing becomes "which pointer do I pass", not a copy.  Optionally a second
// destination (w_out2) receives the same weights (e.g. a transposed/peer copy).
//
// hyper = device float[4] {lr, momentum, weight_decay, grad_scale}: living on the
// device lets a captured CUDA graph see LR-schedule changes without re-capture.
#include "common.cuh"
#include "api.h"

namespace ddl {

template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    bf16x8 v = *reinterpret_cast<const bf16x8*>(p);
    unpack8(v, f);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    *reinterpret_cast<bf16x8*>(p) = pack8(f);
  }
};

template <typename GradT, typename WT>
__global__ void __launch_bounds__(256)
sgd_flat_kernel(float* __restrict__ master, float* __restrict__ mom,
                const GradT* __restrict__ grad, WT* __restrict__ w_out,
                WT* __restrict__ w_out2, const float* __restrict__ hyper, int64_t n) {
  pdl_enter();
  const float lr = hyper[0], mu = hyper[1], wd = hyper[2], gs = hyper[3];
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int64_t i = v << 3;
    float p[8], m[8], g[8];
    Vec8<float>::load(master + i, p);
    Vec8<float>::load(mom + i, m);
    Vec8<GradT>::load(grad + i, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gk = fmaf(wd, p[k], g[k] * gs);
      m[k] = fmaf(mu, m[k], gk);
      p[k] = fmaf(-lr, m[k], p[k]);
    }
    Vec8<float>::store(master + i, p);
    Vec8<float>::store(mom + i, m);
    Vec8<WT>::store(w_out + i, p);
    if (w_out2) Vec8<WT>::store(w_out2 + i, p);
  }
  // scalar tail (n not a multiple of 8)
  const int64_t tail0 = nvec << 3;
  for (int64_t i = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float p = master[i];
    float gk = fmaf(wd, p, to_f<GradT>(grad[i]) * gs);
    float m = fmaf(mu, mom[i], gk);
    p = fmaf(-lr, m, p);
    master[i] = p;
    mom[i] = m;
    w_out[i] = from_f<WT>(p);
    if (w_out2) w_out2[i] = from_f<WT>(p);
  }
}

// sum of squares of a flat buffer -> out[0] (atomic); used for the gradient-norm
// helper that replaces the reference's THCudaHalfTensor_normall ctypes hack
// (pipedream-fork/profiler/translation/seq2seq/utils.py:209-219).
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ x, float* out, int64_t n) {
  pdl_enter();
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = to_f<T>(x[i]);
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

static inline int grid_for(int64_t nvec, int sms) {
  int64_t blocks = (nvec + 255) / 256;
  int64_t cap = (int64_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace ddl

extern "C" int ddl_sgd_flat(void* master, void* mom, const void* grad, int grad_is_bf16,
                            void* w_out, void* w_out2, int w_is_bf16, const float* hyper,
                            int64_t n, int sms, cudaStream_t stream) {
  using namespace ddl;
  if (n <= 0) return 0;
  int grid = grid_for((n + 7) >> 3, sms);
  float* M = (float*)master;
  float* V = (float*)mom;
  if (grad_is_bf16 && w_is_bf16)
    launch_pdl((sgd_flat_kernel<__nv_bfloat16, __nv_bfloat16>), dim3(grid), dim3(256), 0, stream, 
        M, V, (const __nv_bfloat16*)grad, (__nv_bfloat16*)w_out, (__nv_bfloat16*)w_out2, hyper, n);
  else if (grad_is_bf16 && !w_is_bf16)
    launch_pdl((sgd_flat_kernel<__nv_bfloat16, float>), dim3(grid), dim3(256), 0, stream, 
        M, V, (const __nv_bfloat16*)grad, (float*)w_out, (float*)w_out2, hyper, n);
  else if (!grad_is_bf16 && w_is_bf16)
    launch_pdl((sgd_flat_kernel<float, __nv_bfloat16>), dim3(grid), dim3(256), 0, stream, 
        M, V, (const float*)grad, (__nv_bfloat16*)w_out, (__nv_bfloat16*)w_out2, hyper, n);
  else
    launch_pdl((sgd_flat_kernel<float, float>), dim3(grid), dim3(256), 0, stream, 
        M, V, (const float*)grad, (float*)w_out, (float*)w_out2, hyper, n);
  return (int)cudaGetLastError();
}

extern "C" int ddl_sumsq(const void* x, int is_bf16, float* out, int64_t n, int sms,
                         cudaStream_t stream) {
  using namespace ddl;
  if (n <= 0) return 0;
  int grid = grid_for(n, sms);
  if (is_bf16)
    launch_pdl(sumsq_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, out, n);
  else
    launch_pdl(sumsq_kernel<float>, dim3(grid), dim3(256), 0, stream, (const float*)x, out, n);
  return (int)cudaGetLastError();
}

// run-time switch for programmatic dependent launch (see common.cuh); returns the previous value
extern "C" int ddl_set_pdl(int on) {
  const int prev = ddl::pdl_flag();
  ddl::pdl_flag() = on ? 1 : 0;
  return prev;
}
