// Implicit-GEMM convolution (KH x KW, stride 1 or 2, any padding) on the tcgen05 tensor cores
// for NHWC bf16 activations — the 3x3 convolutions of ResNet / VGG and the strided 3x3 / 1x1
// (downsample) layers:
//
//     Y[n, h, w, :] = sum_{kh, kw}  X[n, s*h + kh - ph, s*w + kw - pw, :] * Wt[:, kh, kw, :]^T
//
// For stride 2 the input tensor map carries element strides {1, 2, 2, 1}: the TMA unit samples
// every second pixel of the shifted box, so the strided case costs nothing extra.
//
// (reference: cuDNN library kernels chosen by PyTorch; BatchNorm statistics are a separate pass.)
//
// No im2col buffer exists anywhere.  An output tile is 128 pixels shaped as a TW x TH x TN box
// of the (w, h, n) index space; for every filter tap (kh, kw) and every 64-channel slice of Cin
// the producer warp issues ONE 4-D TMA load of the input box SHIFTED by (kw - pw, kh - ph):
// coordinates that fall outside the image (the padding halo, negative ones included) are
// zero-filled by the TMA unit, and the 128 x 64 tile lands in shared memory in exactly the
// K-major SWIZZLE_128B layout the MMA descriptor expects.  The matching 64-wide slice of the
// KRSC filter matrix [Cout, KH*KW*Cin] is a plain 2-D box.  So the K loop of the GEMM runs over
// (tap, channel slice) and the accumulator in TMEM sees KH*KW*Cin/16 tcgen05.mma instructions —
// the same pipeline as gemm_tcgen05.cu (TMA warp / MMA warp / 4 epilogue warps, mbarrier ring).
//
// Epilogue: TMEM -> registers -> bias/ReLU -> bf16 tile staged in swizzled shared memory ->
// fused BatchNorm sum / sum-of-squares (rows of pixels outside the image are zeroed) -> 4-D TMA
// store, which clips the out-of-image part of the box.
//
// STATUS: validated on hardware at the start of round 2 (tests/conv_check.py, 14 shapes incl.
// stride 2, statistics, bias + ReLU); superseded on the default path by the persistent family in
// igemm_tcgen05.cu (two CTAs per SM, double-buffered TMEM, dgrad without weight transposes) and
// kept behind DDLB_IGEMM=0 + DDLB_CONV3X3_TCGEN05=1 as the one-tile-per-CTA A/B baseline.
#include <cuda.h>
#include <cstdio>

#include "api.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

struct ConvParams {
  float* col_sums;            // [R][2][Cout] replicated partial sums, or null
  const __nv_bfloat16* bias;  // [Cout] or null
  int relu;
  int NB, H, W, Cin, Cout;    // H, W: INPUT height / width
  int OH, OW, stride;         // output height / width, convolution stride (1 or 2)
  int KH, KW, pad_h, pad_w;
  int TW, TH, TN;             // pixel box of one tile (TW*TH*TN == 128)
  int tiles_w, tiles_h, tiles_nb;
};

template <int BLOCK_N, int kStages>
struct ConvSmem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = BLOCK_M * BLOCK_N * 2;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static_assert(kOutBytes <= kRingBytes, "output staging must fit in the ring");
  static constexpr int kBarOffset = kRingBytes;
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * kStages + 1) * 8;
  static constexpr int kBiasOffset = kTmemPtrOffset + 16;
  static constexpr int kTotal = kBiasOffset + BLOCK_N * 4;
  static constexpr int kDynamic = kTotal + 1024;
};

template <int BLOCK_N, int kStages>
__global__ void __launch_bounds__(kThreads, 2)
conv_bf16_nhwc_kernel(const __grid_constant__ CUtensorMap map_x,
                      const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_y, ConvParams cp) {
  using L = ConvSmem<BLOCK_N, kStages>;
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
  const int tiles_n = (cp.Cout + BLOCK_N - 1) / BLOCK_N;
  const int tile_pix = blockIdx.x / tiles_n;
  const int tile_n = blockIdx.x % tiles_n;
  const int w0 = (tile_pix % cp.tiles_w) * cp.TW;
  const int h0 = ((tile_pix / cp.tiles_w) % cp.tiles_h) * cp.TH;
  const int n0 = (tile_pix / (cp.tiles_w * cp.tiles_h)) * cp.TN;
  const int cblocks = cp.Cin / BLOCK_K;
  const int num_k_blocks = cp.KH * cp.KW * cblocks;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_y);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        const int tap = kb / cblocks, cb = kb % cblocks;
        const int kh = tap / cp.KW, kw = tap % cp.KW;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
        // input box shifted by the tap offset; the halo is zero-filled by the TMA unit
        tma_load_4d(a_dst, &map_x, &full_bar[s], cb * BLOCK_K, w0 * cp.stride + kw - cp.pad_w,
                    h0 * cp.stride + kh - cp.pad_h, n0);
        tma_load_2d(b_dst, &map_w, &full_bar[s], tap * cp.Cin + cb * BLOCK_K, tile_n * BLOCK_N);
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
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                   (kb | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (kb == num_k_blocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 thr)
    const int et = threadIdx.x - 64;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;       // pixel of the tile held by this thread
    // which pixel is it, and is it inside the image?
    const int lw = row % cp.TW, lh = (row / cp.TW) % cp.TH, ln = row / (cp.TW * cp.TH);
    const bool valid = (w0 + lw < cp.OW) && (h0 + lh < cp.OH) && (n0 + ln < cp.NB);
    for (int i = et; i < BLOCK_N; i += 128) {
      const int col = tile_n * BLOCK_N + i;
      bias_s[i] = (cp.bias && col < cp.Cout) ? __bfloat162float(cp.bias[col]) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
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
        v = cp.relu ? fmaxf(v, 0.f) : v;
        f[i] = valid ? v : 0.f;            // out-of-image rows contribute nothing to the sums
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
        tma_store_4d(&map_y, smem + h * (BLOCK_M * 128), tile_n * BLOCK_N + h * 64, w0, h0, n0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (cp.col_sums) {
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
      float* cs = cp.col_sums + (size_t)(tile_pix % DDL_BN_REPLICAS) * 2 * cp.Cout;
      if (col < cp.Cout) {
        atomicAdd(&cs[col], s0);
        atomicAdd(&cs[cp.Cout + col], q0);
      }
      if (col + 1 < cp.Cout) {
        atomicAdd(&cs[col + 1], s1);
        atomicAdd(&cs[cp.Cout + col + 1], q1);
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

template <int BLOCK_N, int kStages>
static int launch_conv(const void* x, const void* w, void* y, ConvParams cp, cudaStream_t st) {
  using L = ConvSmem<BLOCK_N, kStages>;
  CUtensorMap mx, mw, my;
  int rc;
  if ((rc = make_map_nhwc(&mx, x, cp.NB, cp.H, cp.W, cp.Cin, BLOCK_K, cp.TW, cp.TH, cp.TN,
                          cp.stride)))
    return rc;
  const int64_t kdim = (int64_t)cp.KH * cp.KW * cp.Cin;
  if ((rc = make_map_2d(&mw, w, cp.Cout, kdim, kdim, BLOCK_K, BLOCK_N))) return rc;
  if ((rc = make_map_nhwc(&my, y, cp.NB, cp.OH, cp.OW, cp.Cout, 64, cp.TW, cp.TH, cp.TN))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_bf16_nhwc_kernel<BLOCK_N, kStages>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = cp.tiles_w * cp.tiles_h * cp.tiles_nb * ((cp.Cout + BLOCK_N - 1) / BLOCK_N);
  cudaError_t err = launch_pdl((conv_bf16_nhwc_kernel<BLOCK_N, kStages>), dim3(tiles),
                               dim3(kThreads), (size_t)L::kDynamic, st, mx, mw, my, cp);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ddl

// x: [NB, H, W, Cin] bf16 NHWC; w: [Cout, KH, KW, Cin] bf16 (KRSC); y: [NB, OH, OW, Cout] with
// OH = (H + 2*pad_h - KH) / stride + 1.  Stride 1 or 2 (strided input sampling by the TMA unit),
// dilation 1.  Returns -1 for unsupported shapes (the caller falls back to the library).
extern "C" int ddl_conv_bf16_nhwc(const void* x, const void* w, void* y, int NB, int H, int W,
                                  int Cin, int Cout, int KH, int KW, int pad_h, int pad_w,
                                  int stride, float* col_sums, const void* bias, int relu, int sms,
                                  cudaStream_t st) {
  using namespace ddl::gemm;
  if (NB <= 0 || H <= 0 || W <= 0) return 0;
  if (Cin % BLOCK_K || Cout % 64 || KH < 1 || KW < 1 || KH > 7 || KW > 7) return -1;
  if ((stride != 1 && stride != 2) || pad_h < 0 || pad_w < 0 || pad_h >= KH + stride ||
      pad_w >= KW + stride)
    return -1;
  const int OH = (H + 2 * pad_h - KH) / stride + 1, OW = (W + 2 * pad_w - KW) / stride + 1;
  if (OH <= 0 || OW <= 0) return -1;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) |
       reinterpret_cast<uintptr_t>(y)) & 15)
    return -1;
  ConvParams cp{};
  cp.col_sums = col_sums;
  cp.bias = (const __nv_bfloat16*)bias;
  cp.relu = relu;
  cp.NB = NB; cp.H = H; cp.W = W; cp.Cin = Cin; cp.Cout = Cout;
  cp.OH = OH; cp.OW = OW; cp.stride = stride;
  cp.KH = KH; cp.KW = KW; cp.pad_h = pad_h; cp.pad_w = pad_w;
  choose_pixel_box(128, NB, OH, OW, &cp.TW, &cp.TH, &cp.TN);
  if (cp.TW * stride > 256 || cp.TH * stride > 256) return -1;
  cp.tiles_w = (OW + cp.TW - 1) / cp.TW;
  cp.tiles_h = (OH + cp.TH - 1) / cp.TH;
  cp.tiles_nb = (NB + cp.TN - 1) / cp.TN;
  const int pix_tiles = cp.tiles_w * cp.tiles_h * cp.tiles_nb;
  // narrow N tiles when the grid would not fill the machine (7x7 / 14x14 layers)
  const bool narrow = Cout <= 64 || pix_tiles * ((Cout + 127) / 128) < sms;
  if (narrow) return launch_conv<64, 4>(x, w, y, cp, st);
  return launch_conv<128, 3>(x, w, y, cp, st);
}

// Tile box chosen for an [NB, H, W] output (host-only; exposed for tests): {TW, TH, TN}.
extern "C" void ddl_conv_box(int NB, int H, int W, int* out3) {
  ddl::gemm::choose_pixel_box(128, NB, H, W, &out3[0], &out3[1], &out3[2]);
}
