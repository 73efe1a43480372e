// Shared device helpers for the ddlbench_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#define DDL_CHECK_CUDA(expr)                                                     \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return (int)_e;                                       \
  } while (0)

// Replicas of the per-channel partial-sum workspace ([R][2][C] fp32): spreads same-address
// L2 atomics from many CTAs over R copies (bn.cu, gemm_tcgen05.cu epilogue).
#define DDL_BN_REPLICAS 8

namespace ddl {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 x bf16 packed in one 16-byte vector
struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// streaming 16-byte global load/store (bypass L1 allocation for single-use data)
__device__ __forceinline__ int4 ld_stream(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w));
}

// system-scope acquire/release on flags living in peer (NVLink-mapped) memory
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A training step at batch 32 is ~250 short kernels
// in one stream/graph; with ordinary stream order each boundary costs the full launch +
// scheduling latency AFTER the previous grid has drained.  Every kernel here
//   * calls pdl_trigger() first: the next grid in the stream may be scheduled as soon as all
//     CTAs of this one are resident (its launch latency and prologue overlap our tail), and
//   * calls pdl_wait() before its first access to global memory: that blocks until the
//     preceding grid has completed and its writes are visible.
// Because every participating kernel waits before touching memory and cannot complete before
// its own wait returns, completion order is still the stream order (transitively).  Kernels
// launched without the attribute (cuDNN, NCCL) are ordinary full dependencies either way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}

}  // namespace ddl

// Host side: launch with the programmatic-stream-serialization attribute (DDLB_PDL=0 turns it
// off; works under stream capture, the edge becomes a programmatic graph dependency).
#include <cstdlib>
namespace ddl {
// One flag for the whole extension (an inline function's static local is a single object across
// translation units); initialised from DDLB_PDL, switchable at run time (ddl_set_pdl) so a caller
// can retry a failed graph capture without programmatic edges.
inline int& pdl_flag() {
  static int on = [] {
    const char* e = std::getenv("DDLB_PDL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on;
}
inline bool pdl_enabled() { return pdl_flag() != 0; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace ddl
