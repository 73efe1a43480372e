// Weight-gradient GEMM on the tcgen05 tensor cores: both operands MN-major.
//
//     D[M, N] += A^T B        A = dY [K = pixels, M = Cout],  B = X [K = pixels, N = Cin]
//
// (the weight gradient of a 1x1 convolution / linear layer in NHWC: the reduction runs over the
// pixel index, which is the SLOW dimension of both operands.)  Nothing is transposed in memory:
// a TMA box of 64 pixel rows x 64 channels lands in shared memory as 64 rows of 128 bytes, and
// the MMA consumes it through an MN-MAJOR shared-memory descriptor (instruction-descriptor bits
// a_major = b_major = 1): canonical SWIZZLE_128B MN-major layout ((8,8,m),(8,k)) — 64 contiguous
// channels per row, rows 128 B apart, 8-row groups SBO = 1024 B apart, 64-channel atoms LBO apart
// (here one 8 KB TMA box each).  One k-block = 64 pixels = 4 MMAs of K = 16 (the descriptor start
// address advances by 16 rows = 2048 B per MMA).
//
// K (= batch*H*W, 10^3..10^5) is far larger than M*N/(128*128) tiles can cover, so the grid is
// (tiles_m * tiles_n) x splits and every CTA adds its fp32 partial tile into D with 16-byte
// vector atomics straight from the TMEM read-out (no shared-memory staging, no second pass).
//
// STATUS: the default weight-gradient path (ops.conv_wgrad_into / ops.wgrad_gemm_into); validated on
// hardware (tests/igemm_check.py, every ResNet-50 / VGG shape against fp32 autograd).  Split-K
// partial tiles are written with plain stores into fp32 planes and summed by wgrad_reduce_kernel;
// the vector-atomic mode described above is the legacy entry point (ddl_gemm_bf16_nt_mn).  The
// hardware confirmed descriptor variant 0: LBO = stride between 64-element MN atoms,
// SBO = stride between 8-row K groups follows the canonical Major-MN SWIZZLE_128B layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units documented in CUTLASS
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>); `desc_variant`
// (env DDLB_WGRAD_DESC_VARIANT=1) swaps the two — it fails on hardware (max error 21 on the first
// shape, profiles/r2/optin_validation_first_run.log), kept only as the record of that bring-up.
#include <cuda.h>
#include <cstdio>

#include "api.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

constexpr int kMnBox = 64;                       // channels per TMA box (128 B) and pixels per k-block
constexpr int kMnBoxBytes = kMnBox * 128;        // 8 KB

// MN-major SWIZZLE_128B descriptor: lbo = byte stride between 64-element MN atoms, sbo = byte
// stride between 8-row K groups.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo,
                                                            uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}


// Where the finished gradient goes.
//   out == null            : legacy mode, fp32 vector atomics into D (round-1 interface).
//   out != null, 1 k-split : the tile is written straight to `out` (bf16 or fp32, same indexing
//                            as D; `accumulate` adds to what is there — micro-batches after the first).
//   out != null, S k-splits: split s writes its partial tile with PLAIN stores into plane s of the
//                            fp32 workspace D (planes `plane_stride` floats apart, output indexing);
//                            wgrad_reduce_kernel then sums the planes into `out`.  No atomics: fp32
//                            vector atomics at this volume cost several times the GEMM itself
//                            (profiles/r2/conv_vs_cudnn_v3_atomics_splitk.json).
__device__ unsigned long long* g_wgrad_trace_dev = nullptr;
__device__ __forceinline__ unsigned long long wg_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug stamps (tools/wgrad_trace.py): [cta][16]: 0 entry, 1 prologue done, 2 accumulator ready,
// 3 epilogue done, 4.. : producer issued k-block 0, 4, 8, 12, ...
#define WG_STAMP(k)                                                                             \
  do {                                                                                          \
    if (g_wgrad_trace_dev)                                                                      \
      g_wgrad_trace_dev[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (k)] = wg_timer(); \
  } while (0)

struct WgradOut {
  void* out;
  int out_bf16;
  int accumulate;
  long long plane_stride;
};

__device__ __forceinline__ void wgrad_store32(const WgradOut& o, size_t idx, const float* v) {
  if (o.out_bf16) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(o.out) + idx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = v[8 * j + i];
      if (o.accumulate) {
        float g[8];
        unpack8(*reinterpret_cast<const bf16x8*>(dst + 8 * j), g);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += g[i];
      }
      *reinterpret_cast<bf16x8*>(dst + 8 * j) = pack8(f);
    }
  } else {
    float* dst = reinterpret_cast<float*>(o.out) + idx;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      if (o.accumulate) {
        const float4 g = *reinterpret_cast<const float4*>(dst + 4 * j);
        t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
      }
      *reinterpret_cast<float4*>(dst + 4 * j) = t;
    }
  }
}

// Staged tile (shared memory, row stride BLOCK_N + 4 floats) -> row-major global memory, whole rows
// per warp.  MODE 0: bf16 store, 1: bf16 accumulate, 2: fp32 store, 3: fp32 accumulate, 4: fp32
// plane (st.global.cg).  `dst` indexed in elements of its type; `base` = element offset of the tile.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <int BLOCK_N, int MODE>
__device__ __forceinline__ void wgrad_copy_rows(uint32_t st_u32, int wq, int lane, void* dst, size_t base,
                                                int ldd) {
  constexpr int LDS = BLOCK_N + 4;
  constexpr int LPR = BLOCK_N / 4;            // lanes per row (16-byte pieces)
  constexpr int RPI = 32 / LPR;               // rows per warp instruction
  constexpr int ITERS = 32 / RPI;             // this warp's rows: wq + 4 * k
  const int sub = lane / LPR, c = (lane % LPR) * 4;
#pragma unroll
  for (int i0 = 0; i0 < ITERS; i0 += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = wq + 4 * ((i0 + u) * RPI + sub);
      v[u] = lds_f4(st_u32 + (uint32_t)((r * LDS + c) * 4));
    }
    if (MODE == 1 || MODE == 3) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = wq + 4 * ((i0 + u) * RPI + sub);
        const size_t idx = base + (size_t)r * ldd + c;
        if (MODE == 1) {
          const uint2 g = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dst) + idx);
          const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.x));
          const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.y));
          v[u].x += g0.x; v[u].y += g0.y; v[u].z += g1.x; v[u].w += g1.y;
        } else {
          const float4 g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dst) + idx);
          v[u].x += g.x; v[u].y += g.y; v[u].z += g.z; v[u].w += g.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = wq + 4 * ((i0 + u) * RPI + sub);
      const size_t idx = base + (size_t)r * ldd + c;
      if (MODE == 0 || MODE == 1) {
        uint2 pk;
        *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(v[u].x, v[u].y);
        *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(v[u].z, v[u].w);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + idx) = pk;
      } else if (MODE == 4) {
        __stcg(reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + idx), v[u]);
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + idx) = v[u];
      }
    }
  }
}

// Epilogue of both weight-gradient kernels (128 threads).  `row0` / `col0`: global coordinates of the
// tile inside D; `ncols`: valid columns from col0; `stage`: the (idle by now) operand ring.
//
// TMEM hands every thread one ROW of the tile, but D is row-major: written straight from the
// registers each warp store would touch 32 different rows (16 bytes each) — measured 8-11 us per
// CTA of partial-sector traffic, more than the whole main loop of the 14x14 / 7x7 layers
// (profiles/r2/traces/wgrad_trace_v1.log).  So the tile is transposed through shared memory
// (row stride BLOCK_N + 4 floats: conflict-free 16-byte accesses both ways) and every warp then
// writes whole 512-byte rows.
template <int BLOCK_N>
__device__ __forceinline__ void wgrad_epilogue(uint32_t tmem_base, int warp, int lane, float* stage,
                                               float* D, int ldd, int row0, int nrows, int col0,
                                               int ncols, const WgradOut& o, int split, int splits) {
  constexpr int LDS = BLOCK_N + 4;
  const int wq = warp & 3;                       // TMEM lane quadrant == this warp's index here
  {
    const uint32_t mine = smem_u32(stage) + (uint32_t)((wq * 32 + lane) * LDS * 4);
#pragma unroll
    for (int cb = 0; cb < BLOCK_N / 32; ++cb) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cb * 32), r);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + (cb * 32 + 4 * j) * 4),
                     "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                     : "memory");
    }
  }
  if (threadIdx.x == 64) WG_STAMP(13);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 64) WG_STAMP(14);
  const bool vec_ok = (ldd & 3) == 0 && (col0 & 3) == 0;
  const bool direct = o.out != nullptr && splits == 1;
  float* plane = (o.out != nullptr) ? D + (size_t)split * o.plane_stride : D;
  // full tiles (every real layer): branch-free, 8-deep unrolled copy loops — one warp per SM
  // sub-partition here, so the shared-memory loads of 8 rows are issued before the first store
  // (the generic loop below is a serial ld -> st chain: 4.3 us per tile, traces/wgrad_trace_v3.log)
  if (o.out != nullptr && vec_ok && ncols >= BLOCK_N && row0 + BLOCK_M <= nrows) {
    const int mode = !direct ? 4 : (o.out_bf16 ? (o.accumulate ? 1 : 0) : (o.accumulate ? 3 : 2));
    const uint32_t st_u32 = smem_u32(stage);
    const size_t base = (size_t)row0 * ldd + col0;
    switch (mode) {
      case 0: wgrad_copy_rows<BLOCK_N, 0>(st_u32, wq, lane, o.out, base, ldd); break;
      case 1: wgrad_copy_rows<BLOCK_N, 1>(st_u32, wq, lane, o.out, base, ldd); break;
      case 2: wgrad_copy_rows<BLOCK_N, 2>(st_u32, wq, lane, o.out, base, ldd); break;
      case 3: wgrad_copy_rows<BLOCK_N, 3>(st_u32, wq, lane, o.out, base, ldd); break;
      default: wgrad_copy_rows<BLOCK_N, 4>(st_u32, wq, lane, plane, base, ldd); break;
    }
    return;
  }
  for (int r = wq; r < BLOCK_M; r += 4) {
    const int grow = row0 + r;
    if (grow >= nrows) break;
#pragma unroll
    for (int c = lane * 4; c < BLOCK_N; c += 128) {
      if (c >= ncols) continue;
      const float4 v4 = *reinterpret_cast<const float4*>(stage + (size_t)r * LDS + c);
      float v[4] = {v4.x, v4.y, v4.z, v4.w};
      const size_t idx = (size_t)grow * ldd + col0 + c;
      const bool full = vec_ok && c + 4 <= ncols;
      const int cnt = min(4, ncols - c);
      if (direct) {
        if (o.out_bf16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(o.out) + idx;
          if (full) {
            if (o.accumulate) {
              const uint2 g = *reinterpret_cast<const uint2*>(dst);
              const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.x));
              const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.y));
              v[0] += g0.x; v[1] += g0.y; v[2] += g1.x; v[3] += g1.y;
            }
            uint2 pk;
            *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
            *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(dst) = pk;
          } else {
            for (int j = 0; j < cnt; ++j) {
              float f = v[j];
              if (o.accumulate) f += __bfloat162float(dst[j]);
              dst[j] = __float2bfloat16_rn(f);
            }
          }
        } else {
          float* dst = reinterpret_cast<float*>(o.out) + idx;
          if (full) {
            float4 t = make_float4(v[0], v[1], v[2], v[3]);
            if (o.accumulate) {
              const float4 g = *reinterpret_cast<const float4*>(dst);
              t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
            }
            *reinterpret_cast<float4*>(dst) = t;
          } else {
            for (int j = 0; j < cnt; ++j) dst[j] = o.accumulate ? dst[j] + v[j] : v[j];
          }
        }
      } else if (o.out != nullptr) {
        // partial tile -> this split's plane
        float* dst = plane + idx;
        if (full) __stcg(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
        else
          for (int j = 0; j < cnt; ++j) dst[j] = v[j];
      } else if (full) {
        atomicAdd(reinterpret_cast<float4*>(D + idx), make_float4(v[0], v[1], v[2], v[3]));
      } else {
        for (int j = 0; j < cnt; ++j) atomicAdd(D + idx + j, v[j]);
      }
    }
  }
}

// out[i] (+)= sum over `planes` of ws[p * stride + i]: the second half of a split-K weight gradient.
// Block = 32 element groups (float4) x 8 plane lanes: the planes of one element are summed by 8
// threads in parallel (up to ~300 planes for the 64-channel layers) and combined in shared memory.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int planes, long long stride, long long n,
                    void* __restrict__ out, int out_bf16, int accumulate) {
  __shared__ float4 part[8][32];
  pdl_enter();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long i0 = ((long long)blockIdx.x * 32 + tx) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i0 + 4 <= n) {
    int p = ty;
#pragma unroll 4
    for (; p < planes; p += 8) {
      const float4 u = __ldcg(reinterpret_cast<const float4*>(ws + (size_t)p * stride + i0));
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
    }
  } else if (i0 < n) {
    for (int p = ty; p < planes; p += 8) {
      const float* src = ws + (size_t)p * stride + i0;
      a.x += __ldcg(src);
      if (i0 + 1 < n) a.y += __ldcg(src + 1);
      if (i0 + 2 < n) a.z += __ldcg(src + 2);
    }
  }
  part[ty][tx] = a;
  __syncthreads();
  if (ty != 0 || i0 >= n) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const float4 u = part[k][tx];
    a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
  }
  const float v[4] = {a.x, a.y, a.z, a.w};
  const int cnt = (int)((n - i0) < 4 ? (n - i0) : 4);
  if (out_bf16) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(out) + i0;
    if (cnt == 4) {
      float f[4] = {v[0], v[1], v[2], v[3]};
      if (accumulate) {
        const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dst));
        const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dst + 2));
        f[0] += g0.x; f[1] += g0.y; f[2] += g1.x; f[3] += g1.y;
      }
      *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(f[0], f[1]);
      *reinterpret_cast<__nv_bfloat162*>(dst + 2) = __floats2bfloat162_rn(f[2], f[3]);
    } else {
      for (int k = 0; k < cnt; ++k) {
        float f = v[k];
        if (accumulate) f += __bfloat162float(dst[k]);
        dst[k] = __float2bfloat16_rn(f);
      }
    }
  } else {
    float* dst = reinterpret_cast<float*>(out) + i0;
    for (int k = 0; k < cnt; ++k) dst[k] = accumulate ? dst[k] + v[k] : v[k];
  }
}

static inline int launch_wgrad_reduce(const float* ws, int planes, long long stride, long long n,
                                      void* out, int out_bf16, int accumulate, cudaStream_t st) {
  const long long groups = (n + 3) / 4;
  const int blocks = (int)((groups + 31) / 32);
  cudaError_t err = launch_pdl(wgrad_reduce_kernel, dim3(blocks), dim3(256), 0, st, ws, planes, stride, n,
                               out, out_bf16, accumulate);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

struct WgradParams {
  float* D;          // [M, N] fp32, accumulated into
  int M, N, K, ldd;
  int kb_per_split;  // k-blocks (of 64 rows) per CTA
  int desc_variant;  // 0: (LBO = MN atom stride, SBO = K group stride); 1: swapped
  WgradOut o;
};

template <int BLOCK_N, int kStages>
struct WgradSmem {
  static constexpr int kABytes = (BLOCK_M / kMnBox) * kMnBoxBytes;   // 16 KB
  static constexpr int kBBytes = (BLOCK_N / kMnBox) * kMnBoxBytes;   // 8/16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kBarOffset = kRingBytes;
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * kStages + 1) * 8;
  static constexpr int kTotal = kTmemPtrOffset + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kRingBytes >= BLOCK_M * (BLOCK_N + 4) * 4, "the epilogue stages the fp32 tile in the ring");
};

template <int BLOCK_N, int kStages>
__global__ void __launch_bounds__(kThreads, 2)
gemm_bf16_mn_kernel(const __grid_constant__ CUtensorMap map_a,
                    const __grid_constant__ CUtensorMap map_b, WgradParams wp) {
  using L = WgradSmem<BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for the staging accesses)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (wp.N + BLOCK_N - 1) / BLOCK_N;
  const int tile = blockIdx.x;
  const int tile_m = tile / tiles_n, tile_n = tile % tiles_n;
  const int total_kb = (wp.K + kMnBox - 1) / kMnBox;
  const int kb0 = blockIdx.y * wp.kb_per_split;
  int kb1 = kb0 + wp.kb_per_split;
  if (kb1 > total_kb) kb1 = total_kb;
  const int num_k_blocks = kb1 - kb0;            // >= 1 by construction of the grid

  pdl_trigger();
  if (threadIdx.x == 64) WG_STAMP(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  if (threadIdx.x == 64) WG_STAMP(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int i = 0; i < num_k_blocks; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        const int krow = (kb0 + i) * kMnBox;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if ((i & 3) == 0 && i < 36) WG_STAMP(4 + (i >> 2));
        uint8_t* a_dst = smem + s * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
#pragma unroll
        for (int j = 0; j < BLOCK_M / kMnBox; ++j)
          tma_load_2d(a_dst + j * kMnBoxBytes, &map_a, &full_bar[s],
                      tile_m * BLOCK_M + j * kMnBox, krow);
#pragma unroll
        for (int j = 0; j < BLOCK_N / kMnBox; ++j)
          tma_load_2d(b_dst + j * kMnBoxBytes, &map_b, &full_bar[s],
                      tile_n * BLOCK_N + j * kMnBox, krow);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 1, 1);   // both MN-major
    const uint32_t mn_stride = kMnBoxBytes, kgrp_stride = 1024;
    const uint32_t lbo = wp.desc_variant ? kgrp_stride : mn_stride;
    const uint32_t sbo = wp.desc_variant ? mn_stride : kgrp_stride;
    for (int i = 0; i < num_k_blocks; ++i) {
      const int s = i % kStages;
      const uint32_t ph = (i / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
        for (int k = 0; k < kMnBox / UMMA_K; ++k) {
          // 16 pixel rows per MMA: +2048 B inside every 8 KB box
          const uint64_t adesc = make_mnmajor_sw128_desc(a_addr + k * (UMMA_K * 128), lbo, sbo);
          const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr + k * (UMMA_K * 128), lbo, sbo);
          umma_f16(tmem_base, adesc, bdesc, idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (i == num_k_blocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 thr)
    mbar_wait(tmem_full_bar, 0);       // every MMA has completed: the ring is idle and becomes the staging tile
    tc_fence_after();
    if (threadIdx.x == 64) WG_STAMP(2);
    const int col0 = tile_n * BLOCK_N;
    wgrad_epilogue<BLOCK_N>(tmem_base, warp, lane, reinterpret_cast<float*>(smem), wp.D, wp.ldd,
                            tile_m * BLOCK_M, wp.M, col0, wp.N - col0, wp.o, (int)blockIdx.y,
                            (int)gridDim.y);
    if (threadIdx.x == 64) WG_STAMP(3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

template <int BLOCK_N, int kStages>
static int launch_mn(const void* A, const void* B, float* D, int M, int N, int K, int lda, int ldb,
                     int ldd, int splits, int variant, cudaStream_t st, WgradOut o = WgradOut{},
                     int* splits_used = nullptr) {
  using L = WgradSmem<BLOCK_N, kStages>;
  CUtensorMap ma, mb;
  int rc;
  // A: [K rows, M cols], box = 64 columns x 64 rows
  if ((rc = make_map_2d(&ma, A, K, M, lda, kMnBox, kMnBox))) return rc;
  if ((rc = make_map_2d(&mb, B, K, N, ldb, kMnBox, kMnBox))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_mn_kernel<BLOCK_N, kStages>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int total_kb = (K + kMnBox - 1) / kMnBox;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;            // no empty CTAs
  if (splits_used) *splits_used = splits;
  WgradParams wp{D, M, N, K, ldd, per, variant, o};
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  cudaError_t err = launch_pdl((gemm_bf16_mn_kernel<BLOCK_N, kStages>), dim3(tiles, splits),
                               dim3(kThreads), (size_t)L::kDynamic, st, ma, mb, wp);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

// D (fp32 [M, N], leading dimension ldd) += A^T B.  `splits` <= 0 picks enough k-splits to put
// about two CTAs on every SM.  Only fp32 accumulation targets are supported (out_is_bf16 and
// workspace are accepted for interface stability and must be 0 / null).
extern "C" int ddl_gemm_bf16_nt_mn(const void* A, const void* B, void* D, int M, int N, int K,
                                   int lda, int ldb, int ldd, int out_is_bf16, int splits,
                                   float* workspace, int sms, cudaStream_t st) {
  using namespace ddl::gemm;
  (void)workspace;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (out_is_bf16) return -1;
  if ((lda % 8) || (ldb % 8) || (M % 8) || (N % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(D)) & 15)
    return -1;   // TMA sources and the 16-byte vector reductions into D need 16-byte alignment
  static const int variant = [] {
    const char* e = std::getenv("DDLB_WGRAD_DESC_VARIANT");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  const int bn = N <= 64 ? 64 : 128;
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + bn - 1) / bn);
  if (splits <= 0) splits = (2 * sms + tiles - 1) / tiles;
  if (bn == 64) return launch_mn<64, 3>(A, B, (float*)D, M, N, K, lda, ldb, ldd, splits, variant, st);
  return launch_mn<128, 3>(A, B, (float*)D, M, N, K, lda, ldb, ldd, splits, variant, st);
}

// =====================================================================================
// Weight gradient of a KH x KW convolution (stride 1 or 2), same machinery:
//
//     dW[co][kh][kw][ci] += sum_{n,oh,ow} dY[n, oh, ow, co] * X[n, s*oh + kh - ph, s*ow + kw - pw, ci]
//
// One grid slice (blockIdx.z) per filter tap; the reduction runs over boxes of 64 output pixels
// (TW x TH x TN): the dY box and the X box SHIFTED by the tap offset are both 4-D TMA loads
// (the X tensor map carries the element strides of a strided convolution and zero-fills the
// halo); both tiles are consumed MN-major.  Split-K over blockIdx.y, fp32 vector reductions into
// the KRSC gradient matrix [Cout][KH*KW*Cin] at column offset tap*Cin.
namespace ddl {
namespace gemm {

struct ConvWgradParams {
  float* D;                    // [Cout][KH*KW*Cin] fp32, accumulated into
  int Cout, Cin, ldd;          // ldd = KH*KW*Cin
  int KW, pad_h, pad_w, stride;
  int TW, TH, TN, tiles_w, tiles_h, total_tiles;
  int tiles_per_split;
  WgradOut o;
};

template <int BLOCK_N, int kStages>
__global__ void __launch_bounds__(kThreads, 2)
conv_wgrad_bf16_kernel(const __grid_constant__ CUtensorMap map_dy,
                       const __grid_constant__ CUtensorMap map_x, ConvWgradParams wp) {
  using L = WgradSmem<BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for the staging accesses)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (wp.Cin + BLOCK_N - 1) / BLOCK_N;
  const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
  const int tap = blockIdx.z;
  const int kh = tap / wp.KW, kw = tap % wp.KW;
  const int t0 = blockIdx.y * wp.tiles_per_split;
  int t1 = t0 + wp.tiles_per_split;
  if (t1 > wp.total_tiles) t1 = wp.total_tiles;
  const int num_k_blocks = t1 - t0;

  pdl_trigger();
  if (threadIdx.x == 64) WG_STAMP(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  if (threadIdx.x == 64) WG_STAMP(1);

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < num_k_blocks; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        const int pt = t0 + i;
        const int w0 = (pt % wp.tiles_w) * wp.TW;
        const int h0 = ((pt / wp.tiles_w) % wp.tiles_h) * wp.TH;
        const int n0 = (pt / (wp.tiles_w * wp.tiles_h)) * wp.TN;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if ((i & 3) == 0 && i < 36) WG_STAMP(4 + (i >> 2));
        uint8_t* a_dst = smem + s * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
#pragma unroll
        for (int j = 0; j < BLOCK_M / kMnBox; ++j)
          tma_load_4d(a_dst + j * kMnBoxBytes, &map_dy, &full_bar[s],
                      tile_m * BLOCK_M + j * kMnBox, w0, h0, n0);
#pragma unroll
        for (int j = 0; j < BLOCK_N / kMnBox; ++j)
          tma_load_4d(b_dst + j * kMnBoxBytes, &map_x, &full_bar[s],
                      tile_n * BLOCK_N + j * kMnBox, w0 * wp.stride + kw - wp.pad_w,
                      h0 * wp.stride + kh - wp.pad_h, n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 1, 1);
    for (int i = 0; i < num_k_blocks; ++i) {
      const int s = i % kStages;
      const uint32_t ph = (i / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
        for (int k = 0; k < kMnBox / UMMA_K; ++k) {
          const uint64_t adesc = make_mnmajor_sw128_desc(a_addr + k * (UMMA_K * 128), kMnBoxBytes, 1024);
          const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr + k * (UMMA_K * 128), kMnBoxBytes, 1024);
          umma_f16(tmem_base, adesc, bdesc, idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (i == num_k_blocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) WG_STAMP(2);
    if (threadIdx.x == 96) WG_STAMP(15);
    const int ci0 = tile_n * BLOCK_N;
    wgrad_epilogue<BLOCK_N>(tmem_base, warp, lane, reinterpret_cast<float*>(smem), wp.D, wp.ldd,
                            tile_m * BLOCK_M, wp.Cout, tap * wp.Cin + ci0, wp.Cin - ci0, wp.o,
                            (int)blockIdx.y, (int)gridDim.y);
    if (threadIdx.x == 64) WG_STAMP(3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

template <int BLOCK_N, int kStages>
static int launch_conv_wgrad(const void* dy, const void* x, float* D, int NB, int H, int W, int OH,
                             int OW, int Cin, int Cout, int KH, int KW, int pad_h, int pad_w,
                             int stride, int sms, cudaStream_t st, WgradOut o = WgradOut{},
                             int max_planes = 1 << 30, int* splits_used = nullptr) {
  using L = WgradSmem<BLOCK_N, kStages>;
  ConvWgradParams wp{};
  wp.o = o;
  wp.D = D; wp.Cout = Cout; wp.Cin = Cin; wp.ldd = KH * KW * Cin;
  wp.KW = KW; wp.pad_h = pad_h; wp.pad_w = pad_w; wp.stride = stride;
  choose_pixel_box(kMnBox, NB, OH, OW, &wp.TW, &wp.TH, &wp.TN);
  wp.tiles_w = (OW + wp.TW - 1) / wp.TW;
  wp.tiles_h = (OH + wp.TH - 1) / wp.TH;
  wp.total_tiles = wp.tiles_w * wp.tiles_h * ((NB + wp.TN - 1) / wp.TN);
  CUtensorMap mdy, mx;
  int rc;
  if ((rc = make_map_nhwc(&mdy, dy, NB, OH, OW, Cout, kMnBox, wp.TW, wp.TH, wp.TN))) return rc;
  if ((rc = make_map_nhwc(&mx, x, NB, H, W, Cin, kMnBox, wp.TW, wp.TH, wp.TN, stride))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_bf16_kernel<BLOCK_N, kStages>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((Cout + BLOCK_M - 1) / BLOCK_M) * ((Cin + BLOCK_N - 1) / BLOCK_N);
  const int taps = KH * KW;
  // enough (tile, tap) units to cover 2/3 of the SMs: no split, the tile goes straight to the
  // gradient buffer (no planes, no reduce launch); otherwise ~1.5 CTAs per SM
  int splits = tiles * taps >= (2 * sms) / 3 ? 1 : (3 * sms / 2 + tiles * taps - 1) / (tiles * taps);
  if (splits > wp.total_tiles / 8) splits = wp.total_tiles / 8;   // a split is worth >= 8 k-blocks
  if (splits < 1) splits = 1;
  if (splits > max_planes) splits = max_planes;
  wp.tiles_per_split = (wp.total_tiles + splits - 1) / splits;
  splits = (wp.total_tiles + wp.tiles_per_split - 1) / wp.tiles_per_split;
  if (splits_used) *splits_used = splits;
  cudaError_t err = launch_pdl((conv_wgrad_bf16_kernel<BLOCK_N, kStages>), dim3(tiles, splits, taps),
                               dim3(kThreads), (size_t)L::kDynamic, st, mdy, mx, wp);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

// dy: [NB, OH, OW, Cout], x: [NB, H, W, Cin] bf16 NHWC; D: fp32 [Cout][KH][KW][Cin], accumulated.
extern "C" int ddl_conv_wgrad_bf16_nhwc(const void* dy, const void* x, float* D, int NB, int H, int W,
                                        int Cin, int Cout, int KH, int KW, int pad_h, int pad_w,
                                        int stride, int sms, cudaStream_t st) {
  using namespace ddl::gemm;
  if (NB <= 0 || H <= 0 || W <= 0) return 0;
  if (Cin % 64 || Cout % 64 || KH < 1 || KW < 1 || KH > 7 || KW > 7) return -1;
  if (stride != 1 && stride != 2) return -1;
  const int OH = (H + 2 * pad_h - KH) / stride + 1, OW = (W + 2 * pad_w - KW) / stride + 1;
  if (OH <= 0 || OW <= 0) return -1;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
       reinterpret_cast<uintptr_t>(D)) & 15)
    return -1;
  if (Cin <= 64)
    return launch_conv_wgrad<64, 3>(dy, x, D, NB, H, W, OH, OW, Cin, Cout, KH, KW, pad_h, pad_w, stride, sms, st);
  return launch_conv_wgrad<128, 3>(dy, x, D, NB, H, W, OH, OW, Cin, Cout, KH, KW, pad_h, pad_w, stride, sms, st);
}

// ---- weight gradients straight into the gradient buffer: one launch when a single k-split fills
// the machine, otherwise the splits write fp32 planes into `ws` (`ws_floats` available; plane =
// M*ldd floats, output indexing) and a small reduce kernel sums them into `out` (bf16 or fp32;
// `accumulate`: add to it).  No atomics, nothing to zero.
extern "C" int ddl_wgrad_gemm(const void* A, const void* B, float* ws, long long ws_floats, void* out,
                              int out_bf16, int accumulate, int M, int N, int K, int lda, int ldb,
                              int ldd, int sms, cudaStream_t st) {
  using namespace ddl::gemm;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (M % 8) || (N % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(out)) & 15)
    return -1;
  const int bn = N <= 64 ? 64 : 128;
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + bn - 1) / bn);
  const long long plane = (long long)M * ldd;
  int splits = tiles >= (2 * sms) / 3 ? 1 : (3 * sms / 2 + tiles - 1) / tiles;
  {
    const int total_kb = (K + kMnBox - 1) / kMnBox;
    if (splits > total_kb / 8) splits = total_kb / 8;      // a split is worth >= 8 k-blocks
  }
  const long long max_planes = ws ? ws_floats / plane : 0;
  if (splits > max_planes) splits = (int)max_planes;
  if (splits < 1) splits = 1;
  WgradOut o{out, out_bf16, accumulate, plane};
  int used = 1;
  const int rc = bn == 64 ? launch_mn<64, 3>(A, B, ws, M, N, K, lda, ldb, ldd, splits, 0, st, o, &used)
                          : launch_mn<128, 3>(A, B, ws, M, N, K, lda, ldb, ldd, splits, 0, st, o, &used);
  if (rc || used == 1) return rc;
  return launch_wgrad_reduce(ws, used, plane, plane, out, out_bf16, accumulate, st);
}

extern "C" int ddl_wgrad_conv(const void* dy, const void* x, float* ws, long long ws_floats, void* out,
                              int out_bf16, int accumulate, int NB, int H, int W, int Cin, int Cout,
                              int KH, int KW, int pad_h, int pad_w, int stride, int sms,
                              cudaStream_t st) {
  using namespace ddl::gemm;
  if (NB <= 0 || H <= 0 || W <= 0) return 0;
  if (Cin % 8 || Cout % 8 || KH < 1 || KW < 1 || KH > 7 || KW > 7) return -1;
  if (stride != 1 && stride != 2) return -1;
  const int OH = (H + 2 * pad_h - KH) / stride + 1, OW = (W + 2 * pad_w - KW) / stride + 1;
  if (OH <= 0 || OW <= 0) return -1;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
       reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(out)) & 15)
    return -1;
  const long long plane = (long long)Cout * KH * KW * Cin;
  long long max_planes = ws ? ws_floats / plane : 0;
  if (max_planes < 1) max_planes = 1;
  WgradOut o{out, out_bf16, accumulate, plane};
  int used = 1;
  const int mp = (int)(max_planes > (1 << 20) ? (1 << 20) : max_planes);
  const int rc = Cin <= 64
      ? launch_conv_wgrad<64, 3>(dy, x, ws, NB, H, W, OH, OW, Cin, Cout, KH, KW, pad_h, pad_w, stride, sms, st, o, mp, &used)
      : launch_conv_wgrad<128, 3>(dy, x, ws, NB, H, W, OH, OW, Cin, Cout, KH, KW, pad_h, pad_w, stride, sms, st, o, mp, &used);
  if (rc || used == 1) return rc;
  return launch_wgrad_reduce(ws, used, plane, plane, out, out_bf16, accumulate, st);
}

// debug: per-CTA stage timestamps of the following weight-gradient launches ([ctas][16] uint64; null: off)
extern "C" void ddl_wgrad_set_trace(unsigned long long* buf) {
  cudaMemcpyToSymbol(ddl::gemm::g_wgrad_trace_dev, &buf, sizeof(buf));
}
