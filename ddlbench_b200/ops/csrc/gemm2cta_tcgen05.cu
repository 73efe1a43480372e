// CTA-pair GEMM on the tcgen05 tensor cores:  D[M,N] = A[M,K] * B[N,K]^T (+ bias, ReLU), bf16 in,
// fp32 accumulate, bf16 out — the large-GEMM core (Linear layers / 1x1 convolutions with a long
// reduction, the cuBLAS comparison of profiles/gemm_tcgen05_vs_cublas.json).
//
// (reference: every GEMM of the reference is a cuBLAS call made by PyTorch, e.g. the classifier of
//  benchmark/imagenet/gpipemodels/vgg/vgg.py:28-46)
//
// Structure — what the single-CTA kernels (igemm_tcgen05.cu) cannot do:
//   * a CLUSTER OF TWO CTAs (one per SM of a TPC) computes one 256 x BLOCK_N tile with
//     `tcgen05.mma.cta_group::2`: each CTA stages its own 128 rows of A and HALF of the B tile; the
//     tensor cores of both SMs read both halves, so every B byte is fetched from L2 once per PAIR
//     (the per-SM L2->SM bandwidth, ~42 B/clk, is what bounds a 128x128 single-CTA tile);
//   * both CTAs' TMA loads complete on the LEADER's mbarrier (cp.async.bulk.tensor ... .cta_group::2),
//     the leader's single elected thread issues every MMA, and its tcgen05.commit is MULTICAST to
//     the barriers of both CTAs (ring slots free in both, accumulator-ready seen by both epilogues);
//   * persistent: 74 clusters walk the tile list (grouped so that a wave shares A rows / B columns
//     in L2); TMEM holds TWO 256-column accumulators per CTA (all 512 columns), so the 8 epilogue
//     warps of each CTA drain tile i (tcgen05.ld -> bias/ReLU -> bf16 -> swizzled staging -> TMA
//     store, per warp, double-buffered) while the tensor cores already accumulate tile i+1;
//   * the peer's epilogue warps hand an accumulator back by arriving REMOTELY (mapa +
//     mbarrier.arrive.shared::cluster) on the leader's barrier.
//
// Every spin is bounded (mbar_wait traps after ~2 s) — a protocol bug cannot hang the GPU.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "api.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

constexpr int k2Threads = 320;           // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address

struct Gemm2Params {
  int M, N, K;
  int tiles_m, tiles_n;       // pair tiles: 256 rows x BLOCK_N columns
  int group_m;                // m-tiles per raster group
  const __nv_bfloat16* bias;  // [N] or null
  int relu;
};

template <int BLOCK_N, int kStages>
struct G2Smem {
  static constexpr int kABytes = 128 * BLOCK_K * 2;                 // this CTA's 128 rows of A
  static constexpr int kBBytes = (BLOCK_N / 2) * BLOCK_K * 2;       // this CTA's half of B
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kOutOffset = kRingBytes;                     // 8 warps x 2 buffers x 4 KB
  static constexpr int kOutBytes = 8 * 2 * 4096;
  static constexpr int kBarOffset = kOutOffset + kOutBytes;         // full[], empty[], tfull[2], tempty[2]
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * kStages + 4) * 8;
  static constexpr int kTotal = kTmemPtrOffset + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kDynamic <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's barrier (same offset as `bar` here)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 0xFFFF) == 0xFFFF && clock64() - t0 > kSpinCycles) __trap();
  }
}
__device__ __forceinline__ void tmem_ld_32x32_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// tile t of the persistent walk -> (m-tile, n-tile): groups of `group_m` m-tiles, n-major inside a
// group, so that the ~74 tiles in flight cover few distinct A row blocks and B column blocks
__device__ __forceinline__ void tile_coords(const Gemm2Params& p, int t, int& tm, int& tn) {
  const int per_group = p.group_m * p.tiles_n;
  const int g = t / per_group;
  const int first = g * p.group_m;
  const int gsz = min(p.group_m, p.tiles_m - first);
  const int r = t - g * per_group;
  tm = first + r % gsz;
  tn = r / gsz;
}

template <int BLOCK_N, int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm2cta_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_d,
                     const __grid_constant__ Gemm2Params p) {
  using L = G2Smem<BLOCK_N, kStages>;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;          // two accumulators (512 columns at N = 256)
  constexpr int kHalfN = BLOCK_N / 2;                   // B rows staged per CTA == columns per epilogue half
  constexpr int kChunks = kHalfN / 64;                  // 64-column store chunks per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs (the MMA and the multicast commit address the peer by offset)
  // (offset arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for the staging accesses)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;           // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]  (only the leader's are waited on)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);          // the leader's producer (expect_tx for BOTH CTAs' bytes)
      mbar_init(&empty_bar[s], 1);         // multicast commit
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 16);         // 8 epilogue warps x 2 CTAs
    mbar_init(&tempty_bar[1], 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = cluster_id; t < n_tiles; t += n_clusters) {
        int tm, tn;
        tile_coords(p, t, tm, tn);
        const int row_a = tm * 256 + (int)rank * 128;
        const int row_b = tn * BLOCK_N + (int)rank * kHalfN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * L::kStageBytes;
          uint8_t* b_dst = a_dst + L::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * L::kStageBytes);
          tma_load_2d_pair(a_dst, &map_a, &full_bar[s], kb * BLOCK_K, row_a);
          tma_load_2d_pair(b_dst, &map_b, &full_bar[s], kb * BLOCK_K, row_b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);
      uint32_t it = 0, iter = 0;
      for (int t = cluster_id; t < n_tiles; t += n_clusters, ++iter) {
        const uint32_t acc = iter & 1, acc_ph = (iter >> 1) & 1;
        mbar_wait_cluster(&tempty_bar[acc], acc_ph ^ 1);   // both epilogues have drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
            const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
            const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + L::kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                            (kb > 0 || k) ? 1u : 0u);
            umma_commit_pair(&empty_bar[s]);
            if (kb == k_blocks - 1) umma_commit_pair(&tfull_bar[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                   // which BLOCK_N/2 columns
    uint8_t* stage = smem + L::kOutOffset + ew * 8192;
    uint32_t iter = 0, nstore = 0;
    for (int t = cluster_id; t < n_tiles; t += n_clusters, ++iter) {
      int tm, tn;
      tile_coords(p, t, tm, tn);
      const uint32_t acc = iter & 1, acc_ph = (iter >> 1) & 1;
      const int row0 = tm * 256 + (int)rank * 128 + quad * 32;      // first row held by this warp
      const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N + half * kHalfN;
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        const int col0 = tn * BLOCK_N + half * kHalfN + ch * 64;
        float f[64];
        {
          uint32_t r[64];
          tmem_ld_32x32_nw(tacc + ch * 64, r);
          tmem_ld_32x32_nw(tacc + ch * 64 + 32, r + 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = __uint_as_float(r[i]);
        }
        if (ch == kChunks - 1) {
          // accumulator fully read: hand it back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tempty_bar[acc], 0);
        }
        if (p.bias) {
          const int c_lo = col0 + lane, c_hi = col0 + 32 + lane;
          const float b_lo = c_lo < p.N ? __bfloat162float(p.bias[c_lo]) : 0.f;
          const float b_hi = c_hi < p.N ? __bfloat162float(p.bias[c_hi]) : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            f[i] += __shfl_sync(0xffffffffu, b_lo, i);
            f[32 + i] += __shfl_sync(0xffffffffu, b_hi, i);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        // this warp's staging buffer `nstore & 1` was last read by the store before the previous one
        uint8_t* buf = stage + (nstore & 1) * 4096;
        if (lane == 0 && nstore >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bf16x8 pk = pack8(f + j * 8);
          *reinterpret_cast<bf16x8*>(buf + lane * 128 + ((j ^ (lane & 7)) * 16)) = pk;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && row0 < p.M && col0 < p.N) {
          tma_store_2d(&map_d, buf, col0, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else if (lane == 0) {
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // keep the group count in step
        }
        ++nstore;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  // nobody leaves while the peer may still read this CTA's shared memory, arrive on its barriers
  // or while MMAs that write this CTA's TMEM are in flight (the epilogue waited for all of them)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

template <int BLOCK_N, int kStages>
static int launch_g2(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                     int ldd, const void* bias, int relu, int sms, cudaStream_t st) {
  using L = G2Smem<BLOCK_N, kStages>;
  CUtensorMap ma, mb, md;
  int rc;
  if ((rc = make_map_2d(&ma, A, M, K, lda, BLOCK_K, 128))) return rc;
  if ((rc = make_map_2d(&mb, B, N, K, ldb, BLOCK_K, BLOCK_N / 2))) return rc;
  if ((rc = make_map_2d(&md, D, M, N, ldd, 64, 32))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2cta_bf16_kernel<BLOCK_N, kStages>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = (M + 255) / 256;
  p.tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  p.group_m = 8;
  if (const char* e = getenv("DDLB_G2_GROUP")) p.group_m = atoi(e) > 0 ? atoi(e) : 8;
  p.bias = (const __nv_bfloat16*)bias;
  p.relu = relu;
  int clusters = sms / 2;
  const int n_tiles = p.tiles_m * p.tiles_n;
  if (clusters > n_tiles) clusters = n_tiles;
  if (clusters < 1) clusters = 1;
  gemm2cta_bf16_kernel<BLOCK_N, kStages><<<dim3(2 * clusters), dim3(k2Threads), L::kDynamic, st>>>(ma, mb, md, p);
  return (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

using namespace ddl::gemm;

// D[M,N] (bf16, ld ldd) = A[M,K] (ld lda) * B[N,K]^T (ld ldb) (+ bias[N]) (ReLU).  Returns -1 when
// the shape is not worth / not possible on the pair kernel (caller falls back to igemm_gemm).
extern "C" int ddl_gemm2cta(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                            int ldd, const void* bias, int relu, int sms, cudaStream_t st) {
  if (M < 1 || N < 1 || K < 1) return -1;
  if ((lda % 8) || (ldb % 8) || (ldd % 8) || (K % 8)) return -1;          // 16-byte TMA strides
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)D) & 15) return -1;
  if (N > 128) return launch_g2<256, 5>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, sms, st);
  return launch_g2<128, 6>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, sms, st);
}
