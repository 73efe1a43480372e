// TMA-fed, TMEM-accumulating bf16 GEMM on the Blackwell 5th-gen tensor cores (tcgen05),
// hand-written for sm_100a with inline PTX.  This is the implicit-GEMM engine behind every
// 1x1 convolution and fully-connected layer of the model zoo in NHWC layout:
//
//     D[M, N] = A[M, K] * B[N, K]^T        M = batch*H*W pixels, N = Cout, K = Cin
//
// (reference: those layers run as cuDNN/cuBLAS library kernels chosen by PyTorch,
//  SURVEY.md §2.9(a); BatchNorm statistics are a separate pass over the output.)
//
// Structure (one 128 x BLOCK_N output tile per CTA, two CTAs co-resident per SM so one
// CTA's epilogue overlaps the other's main loop):
//   warp 0   : TMA producer  — cp.async.bulk.tensor (128B-swizzled boxes) into a 3-stage
//              shared-memory ring, completion signalled on mbarriers
//   warp 1   : allocates TMEM, then ONE elected lane issues tcgen05.mma (M=128, N=BLOCK_N,
//              K=16 per instruction, fp32 accumulator in TMEM) and tcgen05.commit to
//              release ring slots / publish the accumulator
//   warps 2-5: epilogue — tcgen05.ld TMEM->registers, optional bias+ReLU, convert to bf16,
//              stage the tile in (re-used) swizzled shared memory, accumulate the
//              per-output-channel sum / sum-of-squares needed by the following BatchNorm
//              (fused statistics: the BN never re-reads the activation), TMA-store the tile.
//
// Every spin loop is bounded: a protocol bug traps instead of hanging the GPU.
#include <cuda.h>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "api.h"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

struct EpilogueParams {
  float* col_sums;            // [R][2][N] replicated partial sums, or null
  const __nv_bfloat16* bias;  // [N] or null
  int relu;
  int M, N, K;
  // Optional fused BatchNorm finalize: the LAST CTA of the grid turns the partial sums into
  // (mean, invstd), updates the running statistics and re-zeroes the workspace, so the
  // BatchNorm that follows a 1x1 convolution needs no statistics kernel at all.
  unsigned int* ticket;       // grid-wide arrival counter (self-resetting), or null
  float* stat_out;            // [2][N] mean, invstd
  float* running_mean;        // [N] or null
  float* running_var;         // [N] or null
  float eps, momentum;
};

__device__ __forceinline__ void bn_finalize_channel(const EpilogueParams& ep, int c) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int r = 0; r < DDL_BN_REPLICAS; ++r) {
    float* p = ep.col_sums + (size_t)r * 2 * ep.N;
    s += __ldcg(p + c);
    q += __ldcg(p + ep.N + c);
    p[c] = 0.f;
    p[ep.N + c] = 0.f;
  }
  const float inv_m = 1.f / (float)ep.M;
  const float mean = s * inv_m;
  const float var = fmaxf(fmaf(-mean, mean, q * inv_m), 0.f);
  ep.stat_out[c] = mean;
  ep.stat_out[ep.N + c] = rsqrtf(var + ep.eps);
  if (ep.running_mean) {
    const float unbiased = ep.M > 1 ? var * ((float)ep.M / (float)(ep.M - 1)) : var;
    ep.running_mean[c] = fmaf(ep.momentum, mean - ep.running_mean[c], ep.running_mean[c]);
    ep.running_var[c] = fmaf(ep.momentum, unbiased - ep.running_var[c], ep.running_var[c]);
  }
}

template <int BLOCK_N, int kStages>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;       // 16 KB
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;       // 8/16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = BLOCK_M * BLOCK_N * 2;     // staged output tile (aliases stage 0..)
  static constexpr int kRingBytes = kStages * kStageBytes;
  static_assert(kOutBytes <= kRingBytes, "output staging must fit in the ring");
  static constexpr int kBarOffset = kRingBytes;               // full[kStages], empty[kStages], tmem_full
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * kStages + 1) * 8;
  static constexpr int kBiasOffset = kTmemPtrOffset + 16;
  static constexpr int kTotal = kBiasOffset + BLOCK_N * 4;
  static constexpr int kDynamic = kTotal + 1024;              // slack for 1024-B alignment
};

// kStages is chosen from K on the host: the 1x1 convolutions of the first ResNet stages have
// K = 64..256 (1-4 k-blocks), where a short ring frees shared memory for more co-resident CTAs
// (ncu on the first version: 2 CTAs/SM, 16% warps active, one long load->MMA->epilogue chain
// per CTA; the kernel is latency-bound, so more CTAs in flight is what hides it).
template <int BLOCK_N, int kStages>
__global__ void __launch_bounds__(kThreads, (kStages <= 2) ? 4 : 2)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a,
                    const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_d, EpilogueParams ep) {
  using L = SmemLayout<BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);
  float* bias_s = reinterpret_cast<float*>(smem + L::kBiasOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (ep.N + BLOCK_N - 1) / BLOCK_N;
  const int tile_m = blockIdx.x / tiles_n;
  const int tile_n = blockIdx.x % tiles_n;
  const int num_k_blocks = (ep.K + BLOCK_K - 1) / BLOCK_K;

  pdl_trigger();   // the next kernel's launch + prologue may overlap this grid's tail
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
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
  // barrier init, TMEM allocation and descriptor prefetch above overlap the previous grid;
  // from here on we read its output
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
        tma_load_2d(a_dst, &map_a, &full_bar[s], kb * BLOCK_K, tile_m * BLOCK_M);
        tma_load_2d(b_dst, &map_b, &full_bar[s], kb * BLOCK_K, tile_n * BLOCK_N);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
        const uint64_t bdesc = make_kmajor_sw128_desc(b_addr);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the >>4 field
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                   (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                       // slot free when these MMAs retire
        if (kb == num_k_blocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 thr)
    const int et = threadIdx.x - 64;       // 0..127
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // row of the tile held by this thread
    for (int i = et; i < BLOCK_N; i += 128) {
      const int col = tile_n * BLOCK_N + i;
      bias_s[i] = (ep.bias && col < ep.N) ? __bfloat162float(ep.bias[col]) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // stage the bf16 tile: BLOCK_N/64 regions of [128 rows][128 B], 16-B chunks XOR-swizzled
    // with (row & 7) — the SWIZZLE_128B pattern the D tensor map expects.
#pragma unroll
    for (int cb = 0; cb < BLOCK_N / 32; ++cb) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb * 32), r);
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = __uint_as_float(r[i]) + bias_s[cb * 32 + i];
        f[i] = ep.relu ? fmaxf(v, 0.f) : v;
      }
      uint8_t* region = smem + (cb / 2) * (BLOCK_M * 128);
      const int chunk0 = (cb & 1) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bf16x8 p = pack8(f + j * 8);
        const int chunk = (chunk0 + j) ^ (row & 7);
        *reinterpret_cast<bf16x8*>(region + row * 128 + chunk * 16) = p;
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (et == 0) {
#pragma unroll
      for (int h = 0; h < BLOCK_N / 64; ++h)
        tma_store_2d(&map_d, smem + h * (BLOCK_M * 128), tile_n * BLOCK_N + h * 64,
                     tile_m * BLOCK_M);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (ep.col_sums) {
      // per-channel sum / sum-of-squares of the (bf16-rounded) tile, rows beyond M are zero
      constexpr int kPairs = BLOCK_N / 2;
      constexpr int kGroups = 128 / kPairs;
      constexpr int kRowsPer = BLOCK_M / kGroups;
      const int pair = et % kPairs, grp = et / kPairs;
      const int c = pair * 2;
      const uint8_t* region = smem + (c / 64) * (BLOCK_M * 128);
      const int cc = c % 64;
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < kRowsPer; ++rr) {
        const int r_ = grp * kRowsPer + rr;
        const int chunk = (cc >> 3) ^ (r_ & 7);
        const __nv_bfloat162 v =
            *reinterpret_cast<const __nv_bfloat162*>(region + r_ * 128 + chunk * 16 + (cc & 7) * 2);
        const float2 fv = __bfloat1622float2(v);
        s0 += fv.x; s1 += fv.y;
        q0 = fmaf(fv.x, fv.x, q0); q1 = fmaf(fv.y, fv.y, q1);
      }
      const int col = tile_n * BLOCK_N + c;
      // replicated workspace [R][2][N]: M-tile t adds into replica t % R (see bn.cu)
      float* cs = ep.col_sums + (size_t)(tile_m % DDL_BN_REPLICAS) * 2 * ep.N;
      if (col < ep.N) {
        atomicAdd(&cs[col], s0);
        atomicAdd(&cs[ep.N + col], q0);
      }
      if (col + 1 < ep.N) {
        atomicAdd(&cs[col + 1], s1);
        atomicAdd(&cs[ep.N + col + 1], q1);
      }
    }
    if (ep.col_sums && ep.ticket) {
      // last CTA to arrive finalizes the statistics of all N channels
      __shared__ bool is_last;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const unsigned int t = atomicAdd(ep.ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ep.ticket = 0u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (is_last) {
        __threadfence();
        for (int c = et; c < ep.N; c += 128) bn_finalize_channel(ep, c);
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

// ------------------------------------------------------------------------------ host side
template <int BLOCK_N, int kStages>
static int launch_tn(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                     int ldd, float* col_sums, const void* bias, int relu, const DdlBnFinalize* fin,
                     cudaStream_t st) {
  using L = SmemLayout<BLOCK_N, kStages>;
  CUtensorMap ma, mb, md;
  int rc;
  if ((rc = make_map_2d(&ma, A, M, K, lda, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mb, B, N, K, ldb, BLOCK_K, BLOCK_N))) return rc;
  if ((rc = make_map_2d(&md, D, M, N, ldd, 64, BLOCK_M))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, kStages>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  EpilogueParams ep{col_sums, (const __nv_bfloat16*)bias, relu, M, N, K,
                    nullptr, nullptr, nullptr, nullptr, 0.f, 0.f};
  if (fin && col_sums) {
    ep.ticket = fin->ticket;
    ep.stat_out = fin->stat_out;
    ep.running_mean = fin->running_mean;
    ep.running_var = fin->running_var;
    ep.eps = fin->eps;
    ep.momentum = fin->momentum;
  }
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  cudaError_t err = launch_pdl((gemm_bf16_tn_kernel<BLOCK_N, kStages>), dim3(tiles), dim3(kThreads),
                               (size_t)L::kDynamic, st, ma, mb, md, ep);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

extern "C" int ddl_gemm_bf16_tn(const void* A, const void* B, void* D, int M, int N, int K,
                                int lda, int ldb, int ldd, float* col_sums, const void* bias,
                                int relu, const DdlBnFinalize* fin, int sms, cudaStream_t st) {
  using namespace ddl::gemm;
  (void)sms;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldd % 8) || (N % 8) || (K % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(D)) & 15)
    return -1;
#define DDL_GEMM_ARGS A, B, D, M, N, K, lda, ldb, ldd, col_sums, bias, relu, fin, st
  const int kb = (K + BLOCK_K - 1) / BLOCK_K;
  if (N <= 64) {
    if (kb <= 1) return launch_tn<64, 1>(DDL_GEMM_ARGS);
    if (kb <= 2) return launch_tn<64, 2>(DDL_GEMM_ARGS);
    return launch_tn<64, 4>(DDL_GEMM_ARGS);
  }
  if (kb <= 1) return launch_tn<128, 1>(DDL_GEMM_ARGS);
  if (kb <= 2) return launch_tn<128, 2>(DDL_GEMM_ARGS);
  return launch_tn<128, 3>(DDL_GEMM_ARGS);
#undef DDL_GEMM_ARGS
}

extern "C" int ddl_gemm_selftest_info(char* buf, int len) {
  using namespace ddl::gemm;
  snprintf(buf, len,
           "tcgen05 bf16 TN GEMM: BLOCK_M=%d BLOCK_K=%d stages=1..4 smem(N=128,3st)=%d B smem(N=128,1st)=%d B",
           BLOCK_M, BLOCK_K, SmemLayout<128, 3>::kDynamic, SmemLayout<128, 1>::kDynamic);
  return 0;
}
