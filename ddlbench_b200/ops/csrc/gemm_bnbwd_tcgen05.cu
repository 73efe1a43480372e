// Data gradient of a 1x1 convolution whose A operand — the gradient w.r.t. the convolution's
// output, "dc" — is produced ON THE WAY IN by the backward of the BatchNorm that follows the
// convolution (the mirror image of gemm_bnfold_tcgen05.cu):
//
//     dz = (dy [+ dy2]) * act'(y)                  act' from the forward's 1-bit mask
//     dc = A[k] * dz + Bx[k] * c + K[k]            per-channel coefficients of the BN backward
//     dx[M, N] = dc[M, K] * Wt[N, K]^T             N = Cin, K = Cout, Wt = W^T (K-major)
//
// The element-wise bn_bwd_apply pass (read dy / mask / c, write dc: 0.84 ms of the 4.64 ms
// ResNet-50 step, the hottest kernel) does not run.  Per k-block the TMA warp loads the raw tiles
// of dy (and dy2, the second gradient of a fan-out) and c; the four epilogue warps — idle during
// the main loop — rebuild dz and dc in shared memory (thread = one 128-byte row of the
// SWIZZLE_128B tile), hand the dc tile to the MMA warp through a second mbarrier, and the CTAs of
// the first N tile TMA-store dc (the weight-gradient kernel reads it) and, when the layer has a
// residual input, dz (the gradient of the skip connection).
//
// A, Bx, K come from the statistics and the three sums bn_bwd_reduce leaves in `coef`
// (c0 = gamma*invstd, c1 = mean(dz), c2 = mean(dz*xhat)):
//     A = c0,  Bx = -c0*invstd*c2,  K = c0*(mean*invstd*c2 - c1)          (as bn_bwd_apply_kernel)
//
// STATUS: validated on hardware at the start of round 2 (profiles/r2/optin_validation_first_run.log:
// dx / dc / dz / dgamma / dbeta against the unfused path on four shapes); round-1 family
// (DDLB_IGEMM=0 + DDLB_BN_BWD_FOLD=1), kept as an A/B baseline.
#include <cuda.h>
#include <cstdio>

#include "api.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

constexpr int kBwdFoldMaxK = 2048;       // coefficients for the CTA's whole K range live in smem

struct BwdFoldParams {
  int M, N, K;
  const float* stat;          // [2][K] mean, invstd
  const float* coef;          // [3][K] c0, c1, c2 (bn_bwd_reduce)
  int act;                    // 0 none, 1 relu, 2 relu6 (only selects "mask or not")
  const uint8_t* mask;        // [M][K/8] activation-derivative bits, or null when act == 0
  int store_dz;               // 1: materialise dz (gradient of the residual input)
};

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BLOCK_N, int kStages, bool HAS_DY2>
struct BwdFoldSmem {
  static constexpr int kTileBytes = BLOCK_M * BLOCK_K * 2;              // 16 KB
  static constexpr int kDyOff = 0;                                      // dy, becomes dz
  static constexpr int kDy2Off = kTileBytes;                            // dy2 (optional)
  static constexpr int kCOff = kTileBytes * (HAS_DY2 ? 2 : 1);          // c, becomes dc = A operand
  static constexpr int kBOff = kCOff + kTileBytes;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kBOff + kBBytes;
  static constexpr int kOutBytes = BLOCK_M * BLOCK_N * 2;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static_assert(kOutBytes <= kRingBytes, "output staging must fit in the ring");
  static_assert(kStageBytes % 1024 == 0, "every tile must stay 1024-byte aligned");
  static constexpr int kBarOffset = kRingBytes;                          // full, xform, empty, tmem_full
  static constexpr int kTmemPtrOffset = kBarOffset + (3 * kStages + 1) * 8;
  static constexpr int kCoefOffset = kTmemPtrOffset + 16;
  static constexpr int kTotal = kCoefOffset + 3 * kBwdFoldMaxK * 4;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kDynamic <= 227 * 1024, "shared memory budget");
};

template <int BLOCK_N, int kStages, bool HAS_DY2>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bnbwd_tn_kernel(const __grid_constant__ CUtensorMap map_dy,
                     const __grid_constant__ CUtensorMap map_dy2,
                     const __grid_constant__ CUtensorMap map_c,
                     const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_dx,
                     const __grid_constant__ CUtensorMap map_dc,
                     const __grid_constant__ CUtensorMap map_dz, BwdFoldParams fp) {
  using L = BwdFoldSmem<BLOCK_N, kStages, HAS_DY2>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* xform_bar = full_bar + kStages;
  uint64_t* empty_bar = xform_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);
  float* a_s = reinterpret_cast<float*>(smem + L::kCoefOffset);
  float* bx_s = a_s + kBwdFoldMaxK;
  float* k_s = bx_s + kBwdFoldMaxK;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (fp.N + BLOCK_N - 1) / BLOCK_N;
  const int tile_m = blockIdx.x / tiles_n;
  const int tile_n = blockIdx.x % tiles_n;
  const int num_k_blocks = fp.K / BLOCK_K;
  const bool storer = tile_n == 0;         // one CTA per M tile materialises dc (and dz)
  const bool store_dz = storer && fp.store_dz;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    if (HAS_DY2) tma_prefetch_desc(&map_dy2);
    tma_prefetch_desc(&map_c);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_dx);
    tma_prefetch_desc(&map_dc);
    tma_prefetch_desc(&map_dz);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xform_bar[s], 1);
      // a slot is free again when the MMAs that read it retired AND (in a storing CTA) the TMA
      // stores have finished reading it
      mbar_init(&empty_bar[s], storer ? 2 : 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * L::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
        tma_load_2d(st + L::kDyOff, &map_dy, &full_bar[s], kb * BLOCK_K, tile_m * BLOCK_M);
        if (HAS_DY2)
          tma_load_2d(st + L::kDy2Off, &map_dy2, &full_bar[s], kb * BLOCK_K, tile_m * BLOCK_M);
        tma_load_2d(st + L::kCOff, &map_c, &full_bar[s], kb * BLOCK_K, tile_m * BLOCK_M);
        tma_load_2d(st + L::kBOff, &map_b, &full_bar[s], kb * BLOCK_K, tile_n * BLOCK_N);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&xform_bar[s], ph);               // the dc tile, not the raw c tile
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + s * L::kStageBytes);
        const uint64_t adesc = make_kmajor_sw128_desc(st + L::kCOff);
        const uint64_t bdesc = make_kmajor_sw128_desc(st + L::kBOff);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                   (kb | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (kb == num_k_blocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ transform, then epilogue
    const int et = threadIdx.x - 64;       // 0..127
    const int quad = warp & 3;
    const int row = quad * 32 + lane;      // tile row owned by this thread (TMEM-quadrant order)
    const int grow = tile_m * BLOCK_M + row;
    const bool row_ok = grow < fp.M;
    for (int k = et; k < fp.K; k += 128) {
      const float mean = fp.stat[k], istd = fp.stat[fp.K + k];
      const float c0 = fp.coef[k], c1 = fp.coef[fp.K + k], c2 = fp.coef[2 * fp.K + k];
      a_s[k] = c0;
      bx_s[k] = -c0 * istd * c2;
      k_s[k] = c0 * (mean * istd * c2 - c1);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      uint8_t* st = smem + s * L::kStageBytes;
      uint8_t* dy_tile = st + L::kDyOff;
      const uint8_t* dy2_tile = st + L::kDy2Off;
      uint8_t* c_tile = st + L::kCOff;
      // 64 derivative bits of this row for this k-block (K % 64 == 0 is required by the host)
      unsigned long long bits = ~0ull;
      if (fp.act != 0)
        bits = row_ok ? *reinterpret_cast<const unsigned long long*>(
                            fp.mask + (size_t)grow * (fp.K >> 3) + kb * 8)
                      : 0ull;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int p = c ^ (row & 7);                         // physical 16-byte chunk (SWIZZLE_128B)
        bf16x8* dptr = reinterpret_cast<bf16x8*>(dy_tile + row * 128 + p * 16);
        bf16x8* cptr = reinterpret_cast<bf16x8*>(c_tile + row * 128 + p * 16);
        float d[8], x[8];
        unpack8(*dptr, d);
        if (HAS_DY2) {
          float e[8];
          unpack8(*reinterpret_cast<const bf16x8*>(dy2_tile + row * 128 + p * 16), e);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] += e[j];
        }
        const unsigned m = (unsigned)(bits >> (8 * c)) & 0xffu;
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = ((m >> j) & 1u) ? d[j] : 0.f;
        if (fp.store_dz) {
          // the unfused path hands a bf16 dz to the apply kernel: keep the same rounding point
          // (in every CTA of the M tile, so all of them feed the same dc to the tensor cores)
          const bf16x8 packed = pack8(d);
          *dptr = packed;
          unpack8(packed, d);
        }
        unpack8(*cptr, x);
        const int k0 = kb * BLOCK_K + c * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fmaf(a_s[k0 + j], d[j], fmaf(bx_s[k0 + j], x[j], k_s[k0 + j]));
          d[j] = row_ok ? v : 0.f;            // rows beyond M must feed zeros to the MMA
        }
        *cptr = pack8(d);
      }
      // generic-proxy writes -> visible to the async proxy (tcgen05.mma operand reads, TMA stores)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        if (storer) {
          tma_store_2d(&map_dc, c_tile, kb * BLOCK_K, tile_m * BLOCK_M);
          if (store_dz) tma_store_2d(&map_dz, dy_tile, kb * BLOCK_K, tile_m * BLOCK_M);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        mbar_arrive1(&xform_bar[s]);                         // MMA may consume the dc tile
        if (storer) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive1(&empty_bar[s]);                       // ... and the stores are done reading
        }
      }
    }

    // ---------------------------------------------------------------- epilogue: dx tile
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int cb = 0; cb < BLOCK_N / 32; ++cb) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb * 32), r);
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
      uint8_t* region = smem + (cb / 2) * (BLOCK_M * 128);
      const int chunk0 = (cb & 1) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bf16x8 pk = pack8(f + j * 8);
        const int chunk = (chunk0 + j) ^ (row & 7);
        *reinterpret_cast<bf16x8*>(region + row * 128 + chunk * 16) = pk;
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (et == 0) {
#pragma unroll
      for (int h = 0; h < BLOCK_N / 64; ++h)
        tma_store_2d(&map_dx, smem + h * (BLOCK_M * 128), tile_n * BLOCK_N + h * 64,
                     tile_m * BLOCK_M);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

template <int BLOCK_N, int kStages, bool HAS_DY2>
static int launch_bwd_fold(const void* dy, const void* dy2, const void* c, const void* bt, void* dx,
                           void* dc, void* dz, int ld, int ldb, int ldd, BwdFoldParams fp,
                           cudaStream_t st) {
  using L = BwdFoldSmem<BLOCK_N, kStages, HAS_DY2>;
  CUtensorMap mdy, mdy2, mc, mb, mdx, mdc, mdz;
  int rc;
  if ((rc = make_map_2d(&mdy, dy, fp.M, fp.K, ld, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mdy2, HAS_DY2 ? dy2 : dy, fp.M, fp.K, ld, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mc, c, fp.M, fp.K, ld, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mb, bt, fp.N, fp.K, ldb, BLOCK_K, BLOCK_N))) return rc;
  if ((rc = make_map_2d(&mdx, dx, fp.M, fp.N, ldd, 64, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mdc, dc, fp.M, fp.K, ld, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mdz, fp.store_dz ? dz : dc, fp.M, fp.K, ld, BLOCK_K, BLOCK_M))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bnbwd_tn_kernel<BLOCK_N, kStages, HAS_DY2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((fp.M + BLOCK_M - 1) / BLOCK_M) * ((fp.N + BLOCK_N - 1) / BLOCK_N);
  cudaError_t err = launch_pdl((gemm_bnbwd_tn_kernel<BLOCK_N, kStages, HAS_DY2>), dim3(tiles),
                               dim3(kThreads), (size_t)L::kDynamic, st, mdy, mdy2, mc, mb, mdx, mdc,
                               mdz, fp);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

// dy, dy2 (or null), c, dc (written), dz (written, or null): [M, K] bf16 with leading dimension ld;
// bt = W^T [N, K]; dx [M, N] (written).  Returns -1 for shapes the kernel declines.
extern "C" int ddl_gemm_bf16_tn_bnbwd(const void* dy, const void* dy2, const void* c,
                                      const void* mask, const void* bt, void* dx, void* dc, void* dz,
                                      int M, int N, int K, int ld, int ldb, int ldd,
                                      const float* stat, const float* coef, int act, int sms,
                                      cudaStream_t st) {
  using namespace ddl::gemm;
  (void)sms;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % BLOCK_K || K > kBwdFoldMaxK || (N % 8) || (ld % 8) || (ldb % 8) || (ldd % 8)) return -1;
  if (act != 0 && mask == nullptr) return -1;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dy2) |
       reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(bt) |
       reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dc) |
       reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(mask)) & 15)
    return -1;
  BwdFoldParams fp{M, N, K, stat, coef, act, (const uint8_t*)mask, dz != nullptr ? 1 : 0};
  const bool narrow = N <= 64;
  if (dy2) {
    if (narrow) return launch_bwd_fold<64, 2, true>(dy, dy2, c, bt, dx, dc, dz, ld, ldb, ldd, fp, st);
    return launch_bwd_fold<128, 2, true>(dy, dy2, c, bt, dx, dc, dz, ld, ldb, ldd, fp, st);
  }
  if (narrow) return launch_bwd_fold<64, 3, false>(dy, dy2, c, bt, dx, dc, dz, ld, ldb, ldd, fp, st);
  return launch_bwd_fold<128, 3, false>(dy, dy2, c, bt, dx, dc, dz, ld, ldb, ldd, fp, st);
}
