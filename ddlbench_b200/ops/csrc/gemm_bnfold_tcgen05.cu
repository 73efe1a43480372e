// 1x1-convolution GEMM whose A operand is produced ON THE WAY IN by the preceding BatchNorm:
//
//     Y[M, K] = act( C[M, K] * scale[k] + shift[k]  [+ R[M, K]] )          (BN + residual + ReLU)
//     D[M, N] = Y[M, K] * B[N, K]^T                                        (the next convolution)
//
// The element-wise BatchNorm pass (read C, write Y: 0.58 ms of the 4.64 ms ResNet-50 step) never
// runs as a kernel of its own.  Per k-block the TMA warp loads the RAW tile of C (and of the
// residual R) into the shared-memory ring; the four warps that will later run the epilogue —
// idle during the main loop of gemm_tcgen05.cu — rewrite the tile in place
// (thread = one 128-byte row of the SWIZZLE_128B tile, per-channel scale/shift from shared memory),
// emit the 1-bit activation mask, and hand the tile to the MMA warp through a second mbarrier; the
// CTAs of the first N tile also TMA-store the transformed tile to Y, which the backward pass and
// any later consumer (skip connection) read.  Everything else (tcgen05.mma into TMEM, epilogue with
// bias/ReLU, fused BatchNorm sums of D, TMA store) is the pipeline of gemm_tcgen05.cu.
//
// scale = gamma * invstd, shift = beta - mean * scale are derived per CTA from the finalized
// statistics, so no extra launch prepares them.
//
// STATUS: validated on hardware at the start of round 2 (profiles/r2/optin_validation_first_run.log:
// y / mask / GEMM output against the unfused path on four shapes); belongs to the round-1 one-tile
// GEMM family (DDLB_IGEMM=0 + DDLB_BN_FOLD=1) and is kept as an A/B baseline — the persistent
// implicit-GEMM family that is the default today has no operand-transform stage.
#include <cuda.h>
#include <cstdio>

#include "api.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

constexpr int kFoldMaxK = 2048;          // scale/shift for the CTA's whole K range live in smem

struct FoldParams {
  float* col_sums;
  const __nv_bfloat16* bias;
  int relu;
  int M, N, K;
  const float* stat;          // [2][K] mean, invstd of the folded BatchNorm
  const void* gamma;          // [K]
  const void* beta;           // [K]
  int param_is_bf16;
  int act;                    // 0 none, 1 relu, 2 relu6
  uint8_t* mask;              // [M][K/8] or null
  int store_y;                // 1: this launch must materialise Y (always, in practice)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BLOCK_N, int kStages, bool HAS_RES>
struct FoldSmem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;                 // 16 KB (C tile, becomes Y)
  static constexpr int kRBytes = HAS_RES ? kABytes : 0;                 // residual tile
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kRBytes + kBBytes;
  static constexpr int kOutBytes = BLOCK_M * BLOCK_N * 2;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static_assert(kOutBytes <= kRingBytes, "output staging must fit in the ring");
  static constexpr int kBarOffset = kRingBytes;                          // full, xform, empty, tmem_full
  static constexpr int kTmemPtrOffset = kBarOffset + (3 * kStages + 1) * 8;
  static constexpr int kBiasOffset = kTmemPtrOffset + 16;
  static constexpr int kScaleOffset = kBiasOffset + BLOCK_N * 4;
  static constexpr int kTotal = kScaleOffset + 2 * kFoldMaxK * 4;
  static constexpr int kDynamic = kTotal + 1024;
};

template <int BLOCK_N, int kStages, bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bnfold_tn_kernel(const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_r,
                      const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_d,
                      const __grid_constant__ CUtensorMap map_y, FoldParams fp) {
  using L = FoldSmem<BLOCK_N, kStages, HAS_RES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* xform_bar = full_bar + kStages;
  uint64_t* empty_bar = xform_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);
  float* bias_s = reinterpret_cast<float*>(smem + L::kBiasOffset);
  float* scale_s = reinterpret_cast<float*>(smem + L::kScaleOffset);
  float* shift_s = scale_s + kFoldMaxK;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (fp.N + BLOCK_N - 1) / BLOCK_N;
  const int tile_m = blockIdx.x / tiles_n;
  const int tile_n = blockIdx.x % tiles_n;
  const int num_k_blocks = (fp.K + BLOCK_K - 1) / BLOCK_K;
  const bool store_y = fp.store_y && tile_n == 0;      // one CTA per M tile materialises Y / mask

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (HAS_RES) tma_prefetch_desc(&map_r);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
    tma_prefetch_desc(&map_y);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xform_bar[s], 1);
      // a slot is free again when the MMAs that read it retired AND (if this CTA stores Y) the
      // TMA store has finished reading it
      mbar_init(&empty_bar[s], store_y ? 2 : 1);
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
        uint8_t* a_dst = smem + s * L::kStageBytes;
        uint8_t* r_dst = a_dst + L::kABytes;
        uint8_t* b_dst = r_dst + L::kRBytes;
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
        tma_load_2d(a_dst, &map_a, &full_bar[s], kb * BLOCK_K, tile_m * BLOCK_M);
        if (HAS_RES) tma_load_2d(r_dst, &map_r, &full_bar[s], kb * BLOCK_K, tile_m * BLOCK_M);
        tma_load_2d(b_dst, &map_b, &full_bar[s], kb * BLOCK_K, tile_n * BLOCK_N);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&xform_bar[s], ph);               // the TRANSFORMED tile, not the raw one
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes + L::kRBytes;
        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
        const uint64_t bdesc = make_kmajor_sw128_desc(b_addr);
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
    const int et = threadIdx.x - 64;       // 0..127: row of the tile this thread owns
    const int quad = warp & 3;
    const int row = quad * 32 + lane;      // == et for warps 2..5 in TMEM-quadrant order
    const int grow = tile_m * BLOCK_M + row;
    const bool row_ok = grow < fp.M;
    // per-channel scale / shift of the folded BatchNorm for this CTA's whole K range
    for (int k = et; k < num_k_blocks * BLOCK_K; k += 128) {
      float sc = 0.f, sh = 0.f;
      if (k < fp.K) {
        const float g = fp.param_is_bf16 ? __bfloat162float(((const __nv_bfloat16*)fp.gamma)[k])
                                         : ((const float*)fp.gamma)[k];
        const float b = fp.param_is_bf16 ? __bfloat162float(((const __nv_bfloat16*)fp.beta)[k])
                                         : ((const float*)fp.beta)[k];
        sc = g * fp.stat[fp.K + k];
        sh = fmaf(-fp.stat[k], sc, b);
      }
      scale_s[k] = sc;
      shift_s[k] = sh;
    }
    for (int i = et; i < BLOCK_N; i += 128) {
      const int col = tile_n * BLOCK_N + i;
      bias_s[i] = (fp.bias && col < fp.N) ? __bfloat162float(fp.bias[col]) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      uint8_t* a_tile = smem + s * L::kStageBytes;
      const uint8_t* r_tile = a_tile + L::kABytes;
      uint64_t mask_bits = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int p = c ^ (row & 7);                         // physical 16-byte chunk (SWIZZLE_128B)
        bf16x8* ptr = reinterpret_cast<bf16x8*>(a_tile + row * 128 + p * 16);
        float f[8];
        unpack8(*ptr, f);
        const int k0 = kb * BLOCK_K + c * 8;
        if (HAS_RES) {
          float g[8];
          unpack8(*reinterpret_cast<const bf16x8*>(r_tile + row * 128 + p * 16), g);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], scale_s[k0 + j], shift_s[k0 + j]) + g[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], scale_s[k0 + j], shift_s[k0 + j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = f[j];
          if (fp.act == 1) v = fmaxf(v, 0.f);
          else if (fp.act == 2) v = fminf(fmaxf(v, 0.f), 6.f);
          // rows beyond M and channels beyond K must feed zeros to the MMA (their C was
          // zero-filled by TMA but the shift would make them non-zero)
          f[j] = (row_ok && k0 + j < fp.K) ? v : 0.f;
        }
        const bf16x8 packed = pack8(f);
        *ptr = packed;
        if (store_y && fp.mask) {
          float q[8];
          unpack8(packed, q);
          unsigned m = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool on = fp.act == 2 ? (q[j] > 0.f && q[j] < 6.f) : (fp.act == 1 ? q[j] > 0.f : true);
            m |= (on ? 1u : 0u) << j;
          }
          mask_bits |= (uint64_t)m << (8 * c);
        }
      }
      if (store_y && fp.mask && row_ok) {
        // 8 mask bytes of this row for this k-block (K % 64 == 0 is required by the host)
        *reinterpret_cast<uint64_t*>(fp.mask + (size_t)grow * (fp.K >> 3) + kb * 8) = mask_bits;
      }
      // generic-proxy writes -> visible to the async proxy (tcgen05.mma operand reads, TMA store)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        if (store_y) {
          tma_store_2d(&map_y, a_tile, kb * BLOCK_K, tile_m * BLOCK_M);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        mbar_arrive(&xform_bar[s]);                          // MMA may consume the tile
        if (store_y) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&empty_bar[s]);                        // ... and the store is done reading it
        }
      }
    }

    // ---------------------------------------------------------------- epilogue (as gemm_tcgen05.cu)
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int cb = 0; cb < BLOCK_N / 32; ++cb) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb * 32), r);
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = __uint_as_float(r[i]) + bias_s[cb * 32 + i];
        f[i] = fp.relu ? fmaxf(v, 0.f) : v;
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
    if (fp.col_sums) {
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
      float* cs = fp.col_sums + (size_t)(tile_m % DDL_BN_REPLICAS) * 2 * fp.N;
      if (col < fp.N) {
        atomicAdd(&cs[col], s0);
        atomicAdd(&cs[fp.N + col], q0);
      }
      if (col + 1 < fp.N) {
        atomicAdd(&cs[col + 1], s1);
        atomicAdd(&cs[fp.N + col + 1], q1);
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

template <int BLOCK_N, int kStages, bool HAS_RES>
static int launch_fold(const void* A, const void* R, const void* B, void* D, void* Y, int lda, int ldb,
                       int ldd, int ldy, FoldParams fp, cudaStream_t st) {
  using L = FoldSmem<BLOCK_N, kStages, HAS_RES>;
  CUtensorMap ma, mr, mb, md, my;
  int rc;
  if ((rc = make_map_2d(&ma, A, fp.M, fp.K, lda, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mr, HAS_RES ? R : A, fp.M, fp.K, lda, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&mb, B, fp.N, fp.K, ldb, BLOCK_K, BLOCK_N))) return rc;
  if ((rc = make_map_2d(&md, D, fp.M, fp.N, ldd, 64, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&my, Y, fp.M, fp.K, ldy, BLOCK_K, BLOCK_M))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bnfold_tn_kernel<BLOCK_N, kStages, HAS_RES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((fp.M + BLOCK_M - 1) / BLOCK_M) * ((fp.N + BLOCK_N - 1) / BLOCK_N);
  cudaError_t err = launch_pdl((gemm_bnfold_tn_kernel<BLOCK_N, kStages, HAS_RES>), dim3(tiles),
                               dim3(kThreads), (size_t)L::kDynamic, st, ma, mr, mb, md, my, fp);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

// C (raw producer output), R (residual or null), Y (written), all [M, K] bf16 with leading
// dimension ld; B [N, K]; D [M, N].  Returns -1 for shapes the kernel declines.
extern "C" int ddl_gemm_bf16_tn_bnfold(const void* C, const void* R, const void* B, void* D, void* Y,
                                       void* mask, int M, int N, int K, int ld, int ldb, int ldd,
                                       const float* stat, const void* gamma, const void* beta,
                                       int param_is_bf16, int act, float* col_sums, const void* bias,
                                       int relu, int sms, cudaStream_t st) {
  using namespace ddl::gemm;
  (void)sms;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % BLOCK_K || K > kFoldMaxK || (N % 8) || (ld % 8) || (ldb % 8) || (ldd % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D) |
       reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(R) |
       reinterpret_cast<uintptr_t>(mask)) & 15)
    return -1;
  FoldParams fp{col_sums, (const __nv_bfloat16*)bias, relu, M, N, K, stat, gamma, beta, param_is_bf16,
                act, (uint8_t*)mask, 1};
  const bool narrow = N <= 64;
  if (R) {
    if (narrow) return launch_fold<64, 3, true>(C, R, B, D, Y, ld, ldb, ldd, ld, fp, st);
    return launch_fold<128, 3, true>(C, R, B, D, Y, ld, ldb, ldd, ld, fp, st);
  }
  if (narrow) return launch_fold<64, 3, false>(C, R, B, D, Y, ld, ldb, ldd, ld, fp, st);
  return launch_fold<128, 3, false>(C, R, B, D, Y, ld, ldb, ldd, ld, fp, st);
}
