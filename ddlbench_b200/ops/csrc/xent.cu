// Fused softmax-cross-entropy: forward loss, backward dlogits, top-1 / top-5 hit
// counts — ONE launch.
//
// Reference call sites: F.cross_entropy (benchmark/imagenet/imagenet_pytorch.py:153),
// nn.CrossEntropyLoss + a separate accuracy() top-k pass
// (pipedream-fork/runtime/image_classification/main_with_runtime.py:150,639-653) and the
// label-smoothing loss of the GNMT harness
// (pipedream-fork/runtime/translation/seq2seq/train/smoothing.py:7-34).  In the reference
// that is log_softmax + nll_loss forward, two backward kernels, topk + eq + sum.
//
// One CTA per row; the row is cached in shared memory as fp32 so global memory is read
// once and written once.  The mean over rows is produced by the last-arriving CTA in a
// fixed order (deterministic, and self-resetting so a captured CUDA graph can replay it).
#include "common.cuh"
#include "api.h"

namespace ddl {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

template <int BLOCK>
__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect sm reuse
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < BLOCK / 32) ? sm[l] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;  // valid in every thread of every warp
}

// row_stats: [rows][3] = {loss, top1 hit, top5 hit}; out: [3] = {mean loss, #top1, #top5}
template <typename T, int BLOCK, bool CACHE>
__global__ void __launch_bounds__(BLOCK)
xent_fwd_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target,
                    T* __restrict__ dlogits, float* __restrict__ row_stats,
                    float* __restrict__ out, unsigned int* __restrict__ ticket, int rows,
                    int C, float smoothing, float grad_scale) {
  pdl_enter();
  extern __shared__ float row[];
  __shared__ float red[BLOCK / 32];
  __shared__ bool is_last;
  const int r = blockIdx.x;
  const T* x = logits + (int64_t)r * C;
  const int t = (int)target[r];

  float mx = -INFINITY;
  for (int i = threadIdx.x; i < C; i += BLOCK) {
    float v = ldf<T>(x + i);
    if (CACHE) row[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_reduce<BLOCK>(mx, red, true);
  const float xt = CACHE ? row[min(max(t, 0), C - 1)] : ldf<T>(x + min(max(t, 0), C - 1));
  // (row[] written by other threads: the __syncthreads inside block_reduce ordered it)
  float se = 0.f, sx = 0.f, gt = 0.f;
  for (int i = threadIdx.x; i < C; i += BLOCK) {
    float v = CACHE ? row[i] : ldf<T>(x + i);
    se += __expf(v - mx);
    sx += v;
    gt += (v > xt) ? 1.f : 0.f;
  }
  se = block_reduce<BLOCK>(se, red, false);
  sx = block_reduce<BLOCK>(sx, red, false);
  gt = block_reduce<BLOCK>(gt, red, false);
  const float lse = mx + __logf(se);
  const float inv_se = 1.f / se;
  const float eps = smoothing;
  const float on = 1.f - eps, off = eps / (float)C;
  if (dlogits) {
    T* d = dlogits + (int64_t)r * C;
    for (int i = threadIdx.x; i < C; i += BLOCK) {
      float v = CACHE ? row[i] : ldf<T>(x + i);
      float p = __expf(v - mx) * inv_se;
      float g = p - off - ((i == t) ? on : 0.f);
      stf<T>(d + i, g * grad_scale);
    }
  }
  if (threadIdx.x == 0) {
    // loss = (1-eps) * (lse - x_t) + eps * (lse - mean(x))
    float loss = on * (lse - xt) + eps * (lse - sx / (float)C);
    row_stats[r * 3 + 0] = loss;
    row_stats[r * 3 + 1] = (gt < 0.5f) ? 1.f : 0.f;
    row_stats[r * 3 + 2] = (gt < 4.5f) ? 1.f : 0.f;
    __threadfence();
    unsigned int prev = atomicAdd(ticket, 1u);
    is_last = (prev == (unsigned int)rows - 1u);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < rows; i += BLOCK) {
      a += row_stats[i * 3 + 0];
      b += row_stats[i * 3 + 1];
      c += row_stats[i * 3 + 2];
    }
    a = block_reduce<BLOCK>(a, red, false);
    b = block_reduce<BLOCK>(b, red, false);
    c = block_reduce<BLOCK>(c, red, false);
    if (threadIdx.x == 0) {
      out[0] = a / (float)rows;
      out[1] = b;
      out[2] = c;
      *ticket = 0u;  // self-reset for the next launch / graph replay
    }
  }
}

}  // namespace ddl

extern "C" int ddl_xent(const void* logits, int is_bf16, const int64_t* target, void* dlogits,
                        float* row_stats, float* out, unsigned int* ticket, int rows, int C,
                        float smoothing, float grad_scale, cudaStream_t stream) {
  using namespace ddl;
  if (rows <= 0) return 0;
  constexpr int BLOCK = 256;
  const bool cache = (size_t)C * sizeof(float) <= 48 * 1024;
  const size_t smem = cache ? (size_t)C * sizeof(float) : 0;
#define LAUNCH(T, CACHE)                                                                      \
  launch_pdl((xent_fwd_bwd_kernel<T, BLOCK, CACHE>), dim3(rows), dim3(BLOCK), smem, stream,                         \
      (const T*)logits, target, (T*)dlogits, row_stats, out, ticket, rows, C, smoothing,      \
      grad_scale)
  if (is_bf16) {
    if (cache) LAUNCH(__nv_bfloat16, true); else LAUNCH(__nv_bfloat16, false);
  } else {
    if (cache) LAUNCH(float, true); else LAUNCH(float, false);
  }
#undef LAUNCH
  return (int)cudaGetLastError();
}
