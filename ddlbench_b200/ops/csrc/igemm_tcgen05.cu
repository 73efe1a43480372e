// Persistent, warp-specialised implicit-GEMM on the tcgen05 tensor cores — ONE kernel family for
// every dense convolution forward AND data gradient of the model zoo (and the FC layers):
//
//   fprop :  Y[n,oh,ow,:]  = sum_{tap}  X[n, s*oh+kh-ph, s*ow+kw-pw, :] * W[:, kh, kw, :]^T
//   dgrad :  dX[n,h,w,:]   = sum_{tap} dY[n, (h+ph-kh)/s, (w+pw-kw)/s, :] * W[:, kh, kw, :]
//   GEMM  :  D[M,N] = A[M,K] B^T   (1x1 convolutions / linear layers: the 1-tap, flat-pixel case)
//
// (reference: these run as cuDNN / cuBLAS library kernels picked by PyTorch — e.g.
//  benchmark/imagenet/gpipemodels/resnet/block.py:66-74 — with BatchNorm statistics as a
//  separate pass)
//
// What is Blackwell-specific here
//   * A tiles (128 output pixels x 64 reduction channels) are 4-D TMA boxes of the NHWC
//     activation SHIFTED by the tap offset: the padding halo is zero-filled by the TMA unit, a
//     strided convolution is the tensor map's element stride, and no im2col buffer exists.
//   * B (the KRSC filter matrix [Cout][KH*KW*Cin], never transposed or flipped in memory) is
//     consumed K-major for fprop and MN-MAJOR for dgrad (instruction-descriptor b_major bit +
//     the canonical SWIZZLE_128B MN-major shared-memory descriptor), so the data gradient reads
//     the very same weight buffer as the forward pass.
//   * A stride-2 data gradient is decomposed into its 4 output-parity classes; each class is a
//     stride-1 problem over the taps of matching parity whose result is TMA-stored through a
//     strided tensor-map VIEW of dX (no scatter, no zero-insertion).
//   * Every CTA owns a CHUNK of consecutive pixel tiles of one output-channel tile and loops over
//     them with a DOUBLE-BUFFERED TMEM accumulator: the 4 epilogue warps drain tile i
//     (tcgen05.ld -> bias/ReLU -> bf16 -> swizzled staging -> TMA store) while the MMA warp already
//     accumulates tile i+1 and the TMA warp runs ahead through the shared-memory ring across tile
//     boundaries.  CTAs are ~100 KB of shared memory so TWO are co-resident per SM (8 epilogue
//     warps per SM; any two tensor-core kernels of the step can share an SM, which keeps the
//     weight-gradient kernels on the side stream overlapping with the main stream), and the chunks
//     are handed out by the hardware block scheduler (more chunks than slots: no stragglers when a
//     co-running kernel holds some SMs).  (v1 of this kernel was one 214 KB persistent CTA per SM:
//     per layer it was faster, inside the training graph it serialised against the side stream —
//     profiles/r2/.)
//   * BatchNorm statistics are accumulated per CTA in shared memory over all of its tiles and
//     flushed once; the last CTA to finish turns them into (mean, invstd), updates the running
//     statistics and re-zeroes the workspace — the BatchNorm after a convolution needs no
//     statistics or finalize launch at all.
//
// Every spin loop is bounded (mbar_wait traps after ~2 s): a protocol bug cannot hang the GPU.
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "api.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace ddl {
namespace gemm {

constexpr int kIgMaxTaps = 49;
// return value of a launch whose LAST CTA also finalized the BatchNorm statistics (0: sums only)
#define DDL_IGEMM_FINALIZED (-100)

struct IgemmParams {
  int OW, OH, NB;                 // output pixel space (tiles iterate here)
  int Nout;                       // GEMM N = output channels
  int kc_blocks;                  // 64-channel blocks of the reduction per tap
  int a_step;                     // A coordinate = tile origin * a_step + tap offset
  int ntaps;
  int TW, TH, TN, tiles_w, tiles_h, tiles_n, pix_tiles, tiles_per_cta;
  float* col_sums;                // [R][2][Nout] replicated workspace, or null
  const __nv_bfloat16* bias;      // [Nout] or null
  int relu;
  int kb_per_split;               // k-blocks per split-K slice (gridDim.z slices)
  int ncls;                       // > 1: gridDim.z indexes output CLASSES (stride-2 dgrad parities),
                                  // each with its own tap range, output extent and store map
  struct Cls { int OW, OH, tiles_w, tiles_h, pix_tiles, tap0, ntaps, pad; } cls[4];
  int split_l2;                   // cluster split-K flavour: 1 = exchange through L2 (global
                                  // workspace), 0 = through distributed shared memory
  unsigned long long* trace;      // debug: per-CTA globaltimer stamps [cta][16] (null: off)
  int split_w;                    // > 0: CLUSTER split-K (cluster = the gridDim.z slices of a tile);
                                  // each CTA finishes split_w = BLOCK_N / slices columns of the tile
  float* splitk_ws;               // fp32 scratch [tiles][slices][128 x BLOCK_N] (split-K only)
  unsigned int* splitk_tickets;   // one per tile, zero on entry, self-resetting
  unsigned int* ticket;           // fused BatchNorm finalize: one counter PER n-tile (null: sums only)
  float* stat_out;
  float* running_mean;
  float* running_var;
  float eps, momentum;
  long long stat_count;           // rows behind the statistics (NB*OH*OW)
  short tap_dw[kIgMaxTaps], tap_dh[kIgMaxTaps];
  int tap_bcol[kIgMaxTaps];       // column of the tap's block inside the filter matrix
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MN-major SWIZZLE_128B descriptor (same encoding as gemm_wgrad_tcgen05.cu, validated there):
// 64 contiguous N elements per 128-byte row, 8-row K groups 1024 B apart (SBO), 64-element N
// atoms `lbo` bytes apart.
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kSmall: ONE tile per CTA — the staged output tile aliases the (by then idle) ring and there is a
// single TMEM accumulator, so a CTA is 34-66 KB / <= 85 registers and 3-4 of them share an SM: the
// short-K layers (1x1 convolutions of the first stages) are bound by the CUDA-core epilogue, and
// 12-16 epilogue warps per SM hide its latency far better than the 8 of the chunked variant
// (profiles/r2/timeline_igemm_v4.json vs timeline_cudnn_path.json).
template <int BLOCK_N, int kStages, bool kStats, bool kSmall>
struct IgSmem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;            // 16 KB
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;            // 8 / 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kOutBytes = BLOCK_M * BLOCK_N * 2;          // the staged bf16 output tile
  static constexpr int kOutOffset = kSmall ? 0 : kRingBytes;
  static constexpr int kDataBytes = kSmall ? (kRingBytes > kOutBytes ? kRingBytes : kOutBytes)
                                           : kRingBytes + kOutBytes;
  static constexpr int kStatOffset = kDataBytes;
  static constexpr int kStatBytes = kStats ? 2 * BLOCK_N * 4 : 0;  // this CTA's n-tile only
  static constexpr int kBiasOffset = kStatOffset + kStatBytes;
  static constexpr int kBarOffset = kBiasOffset + BLOCK_N * 4;     // full[], empty[], tfull[2], tempty[2]
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * kStages + 4) * 8;
  static constexpr int kTotal = kTmemPtrOffset + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static constexpr int kMinCtas = kSmall ? (kStages == 1 ? 4 : 3) : 2;
  static_assert(kDynamic * kMinCtas <= 227 * 1024, "CTAs per SM vs shared memory");
};

// 32 lanes x 32 fp32 columns -> registers, WITHOUT the wait (callers batch two loads per wait)
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// stage stamps of one thread per CTA (tools/igemm_trace.py): 0 entry, 1 after prologue + pdl_wait,
// 2 accumulator ready, 3 cluster barrier 1, 4 scatter done, 5 cluster barrier 2, 6 tile staged +
// store issued, 7 statistics flushed, 8 ticket / finalize done, 9 exit
#define IG_STAMP(k)                                                                            \
  do {                                                                                         \
    if (p.trace && threadIdx.x == 64)                                                          \
      p.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (k)] = gtimer(); \
  } while (0)

// ---- thread-block-cluster helpers (split-K through distributed shared memory)
__device__ __forceinline__ void cluster_arrive_rel() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acq() {
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <int BLOCK_N, int kStages, bool kBMn, bool kStats, bool kSmall>
__global__ void __launch_bounds__(kThreads, (IgSmem<BLOCK_N, kStages, kStats, kSmall>::kMinCtas))
igemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_d,
                  const __grid_constant__ CUtensorMap map_ds,   // cluster split-K: split_w-column store box
                  const __grid_constant__ CUtensorMap map_d1,   // store maps of output classes 1..3
                  const __grid_constant__ CUtensorMap map_d2,
                  const __grid_constant__ CUtensorMap map_d3,
                  const __grid_constant__ IgemmParams p) {
  using L = IgSmem<BLOCK_N, kStages, kStats, kSmall>;
  constexpr uint32_t kTmemCols = kSmall ? BLOCK_N : 2 * BLOCK_N;   // one / two accumulators
  constexpr int kStep = kSmall ? 32 : 64;              // columns drained per epilogue step
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for the staging accesses)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;           // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);
  float* bias_s = reinterpret_cast<float*>(smem + L::kBiasOffset);
  float* stat_s = reinterpret_cast<float*>(smem + L::kStatOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // this CTA: n-tile blockIdx.x (fastest: CTAs sharing an A tile run together, L2 reuse), pixel
  // tiles [tp0, tp1), k-blocks [kb0, kb1) of every tile (split-K over blockIdx.z)
  const int tile_n = blockIdx.x;
  // output class of this CTA (multi-class launches: gridDim.z = classes, never split-K)
  const bool multi = !kSmall && p.ncls > 1;
  const int cls = multi ? (int)blockIdx.z : 0;
  const int c_OW = multi ? p.cls[cls].OW : p.OW, c_OH = multi ? p.cls[cls].OH : p.OH;
  const int c_tiles_w = multi ? p.cls[cls].tiles_w : p.tiles_w;
  const int c_tiles_h = multi ? p.cls[cls].tiles_h : p.tiles_h;
  const int c_pix_tiles = multi ? p.cls[cls].pix_tiles : p.pix_tiles;
  const int c_tap0 = multi ? p.cls[cls].tap0 : 0;
  const int c_ntaps = multi ? p.cls[cls].ntaps : p.ntaps;
  const CUtensorMap* md_cls = cls == 0 ? &map_d : cls == 1 ? &map_d1 : cls == 2 ? &map_d2 : &map_d3;
  const int tp0 = blockIdx.y * p.tiles_per_cta;
  const int tp1 = min(tp0 + p.tiles_per_cta, c_pix_tiles);
  const int total_kb = c_ntaps * p.kc_blocks;
  const int kb0 = multi ? 0 : blockIdx.z * p.kb_per_split;
  const int kb1 = multi ? total_kb : min(kb0 + p.kb_per_split, total_kb);
  const bool split = !kSmall && !multi && gridDim.z > 1;      // the one-tile variant never splits K
  const bool csplit = split && p.split_w > 0;       // slices of a tile = one thread-block cluster

  pdl_trigger();
  IG_STAMP(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(md_cls);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 4);
    mbar_init(&tempty_bar[1], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  if (kStats && warp >= 2) {
    for (int i = threadIdx.x - 64; i < 2 * BLOCK_N; i += 128) stat_s[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();      // everything above overlapped the previous grid; from here we read its output
  IG_STAMP(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int tp = tp0; tp < tp1; ++tp) {
        const int w0 = (tp % c_tiles_w) * p.TW;
        const int h0 = ((tp / c_tiles_w) % c_tiles_h) * p.TH;
        const int n0 = (tp / (c_tiles_w * c_tiles_h)) * p.TN;
        int t = c_tap0 + kb0 / p.kc_blocks, cb = kb0 % p.kc_blocks;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * L::kStageBytes;
          uint8_t* b_dst = a_dst + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
          // input box shifted by the tap offset; the halo is zero-filled by the TMA unit
          tma_load_4d(a_dst, &map_a, &full_bar[s], cb * BLOCK_K, w0 * p.a_step + p.tap_dw[t],
                      h0 * p.a_step + p.tap_dh[t], n0);
          if (kBMn) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(b_dst + j * 8192, &map_b, &full_bar[s],
                          p.tap_bcol[t] + tile_n * BLOCK_N + j * 64, cb * BLOCK_K);
          } else {
            tma_load_2d(b_dst, &map_b, &full_bar[s], p.tap_bcol[t] + cb * BLOCK_K, tile_n * BLOCK_N);
          }
          if (++cb == p.kc_blocks) { cb = 0; ++t; }
        }
      }
    }
    if (csplit) {          // every thread of the cluster takes part in the cluster barriers
      __syncwarp();
      if (!p.split_l2) { cluster_arrive_rel(); cluster_wait_acq(); }
      cluster_arrive_rel(); cluster_wait_acq();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, kBMn ? 1 : 0);
    uint32_t it = 0, iter = 0;
    for (int tp = tp0; tp < tp1; ++tp, ++iter) {
      const uint32_t acc = iter & 1, acc_ph = (iter >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1);          // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABytes;
          const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // A: +32 B along K inside the swizzle atom; B K-major: same; B MN-major: 16 K rows
            // = 2048 B further down every 8 KB box
            const uint64_t bdesc = kBMn ? make_mn_sw128_desc(b_addr + k * (UMMA_K * 128), 8192)
                                        : make_kmajor_sw128_desc(b_addr) + (uint64_t)(k * 2);
            umma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc, idesc, (kb > kb0 || k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
    }
    if (csplit) {
      __syncwarp();
      if (!p.split_l2) { cluster_arrive_rel(); cluster_wait_acq(); }
      cluster_arrive_rel(); cluster_wait_acq();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int et = threadIdx.x - 64;
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;           // output pixel of the tile held by this thread
    const int lw = row % p.TW, lh = (row / p.TW) % p.TH, ln = row / (p.TW * p.TH);
    __shared__ bool flag_last;
    if (p.bias) {
      for (int i = et; i < BLOCK_N; i += 128) {
        const int col = tile_n * BLOCK_N + i;
        bias_s[i] = col < p.Nout ? __bfloat162float(p.bias[col]) : 0.f;
      }
    }
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;     // this thread's share of the column sums
    int done_tiles = 0;
    uint32_t iter = 0;
    uint8_t* stage = smem + L::kOutOffset;
    for (int tp = tp0; tp < tp1; ++tp, ++iter) {
      const uint32_t acc = iter & 1, acc_ph = (iter >> 1) & 1;
      const int w0 = (tp % c_tiles_w) * p.TW;
      const int h0 = ((tp / c_tiles_w) % c_tiles_h) * p.TH;
      const int n0 = (tp / (c_tiles_w * c_tiles_h)) * p.TN;
      const bool valid = (w0 + lw < c_OW) && (h0 + lh < c_OH) && (n0 + ln < p.NB);
      const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N;
      // split-K: slice z of tile t parks its partial sums in slot [t][z] of the workspace, laid
      // out [(col/4)][row][4] fp32 so that a warp's 16-byte accesses are 512 contiguous bytes.
      // Plain stores, no atomics (fp32 vector atomics at this volume are several times slower
      // than the GEMM itself); the LAST slice to arrive adds the other slots to its own
      // accumulator and runs the epilogue.  Nothing needs zeroing.
      const size_t slot_elems = (size_t)BLOCK_M * BLOCK_N;
      float* wst = split ? p.splitk_ws + (size_t)(tp * gridDim.x + tile_n) * gridDim.z * slot_elems : nullptr;
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      IG_STAMP(2);
      if (csplit) {
        // ---- split-K across a thread-block cluster, reduced through DISTRIBUTED SHARED MEMORY.
        // The S slices of this tile are the S CTAs of one cluster (co-scheduled by the hardware).
        // Each CTA owns W = BLOCK_N / S output columns: it scatters the other column slices of
        // its partial accumulator straight into their owners' (by now idle) rings with
        // st.shared::cluster, and after a cluster barrier adds the S-1 partials it received to
        // its own TMEM slice and runs the epilogue for those W columns.  No global workspace, no
        // atomics, no serial last-arriver: the reduction is spread over all S CTAs.
        const int S = (int)gridDim.z, rank = (int)blockIdx.z, W = p.split_w;
        float* rbuf = reinterpret_cast<float*>(smem);          // [S-1][128][W] fp32 over the ring
        const uint32_t rbuf_u32 = smem_u32(rbuf);
        const bool via_l2 = p.split_l2 != 0;
        // L2 flavour: every CTA parks the column slices it does NOT own in its own 64 KB slot of
        // the global workspace ([column quad][row][4 floats]: coalesced both ways, stays in L2),
        // ONE cluster barrier, then each CTA pulls the S-1 partials of its slice with ld.global.cg.
        // (Measured: remote st.shared::cluster stores run at ~12 B/clk per SM, several times
        // slower than the same bytes through L2 — profiles/r2/igemm_trace_*.log.)
        const size_t slot_f4 = (size_t)BLOCK_M * BLOCK_N / 4;
        float4* ws_tile = reinterpret_cast<float4*>(p.splitk_ws) +
                          (size_t)(tp * gridDim.x + tile_n) * S * slot_f4;
        // (1) DSMEM flavour: all tensor cores of the cluster are done reading their rings
        if (!via_l2) { cluster_arrive_rel(); cluster_wait_acq(); }
        IG_STAMP(3);
        // (2) scatter
#pragma unroll 1
        for (int g = 0; g < BLOCK_N / 32; ++g) {
          const int dest = (g * 32) / W;
          if (dest == rank) continue;
          uint32_t r[32];
          tmem_ld_32x32(tacc + g * 32, r);
          // receive buffer layout [slot][column quad][row][4 floats]: a warp's 16-byte stores are
          // 512 contiguous bytes (coalesced on the SM-to-SM network, conflict-free when read back)
          const int slot = rank - (rank > dest ? 1 : 0);
          const int q0 = ((g * 32) % W) / 4;
          const uint32_t local = rbuf_u32 + (uint32_t)((((slot * (W / 4) + q0) * BLOCK_M) + row) * 16);
          if (via_l2) {
            float4* dst = ws_tile + (size_t)rank * slot_f4 + (size_t)(g * 8) * BLOCK_M + row;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              __stcg(dst + j * BLOCK_M, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                    __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
            continue;
          }
          const uint32_t remote = mapa_u32(local, (uint32_t)dest);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_cluster_f4(remote + j * (BLOCK_M * 16), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                          __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        // (3) every partial sum has landed
        IG_STAMP(4);
        cluster_arrive_rel(); cluster_wait_acq();
        IG_STAMP(5);
        // (4) own slice
        const bool wide = W == 64;                   // 64 columns: the swizzled staging + map_d
#pragma unroll 1
        for (int g = 0; g < W / 32; ++g) {
          float f[32];
          {
            uint32_t r[32];
            tmem_ld_32x32(tacc + rank * W + g * 32, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
          }
          for (int sl = 0; sl < S - 1; ++sl) {
            const int srck = sl + (sl >= rank ? 1 : 0);        // the sl-th OTHER slice
            const float4* src = via_l2 ? ws_tile + (size_t)srck * slot_f4 +
                                             (size_t)((rank * W + g * 32) / 4) * BLOCK_M + row
                                       : reinterpret_cast<const float4*>(rbuf) +
                                             ((size_t)(sl * (W / 4) + g * 8) * BLOCK_M + row);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = via_l2 ? __ldcg(src + j * BLOCK_M) : src[j * BLOCK_M];
              f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = f[i];
            if (p.bias) v += bias_s[rank * W + g * 32 + i];
            if (p.relu) v = fmaxf(v, 0.f);
            f[i] = valid ? v : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bf16x8 pk = pack8(f + j * 8);
            const int chunk = g * 4 + j;             // 16-byte chunk inside the row
            if (wide)
              *reinterpret_cast<bf16x8*>(stage + row * 128 + ((chunk ^ (row & 7)) * 16)) = pk;
            else   // 64-byte rows, SWIZZLE_64B: chunk ^= address bits [7:8]
              *reinterpret_cast<bf16x8*>(stage + row * 64 + ((chunk ^ ((row >> 1) & 3)) * 16)) = pk;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int col0 = tile_n * BLOCK_N + rank * W;
        if (et == 0 && col0 < p.Nout) {
          tma_store_4d(wide ? &map_d : &map_ds, stage, col0, w0, h0, n0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        IG_STAMP(6);
        ++done_tiles;
        if (kStats && p.col_sums) {
          const int pairs = W / 2, groups = 128 / pairs, rows_per = BLOCK_M / groups;
          const int pair = et % pairs, grp = et / pairs;
          const int c = pair * 2;
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
          for (int rr = 0; rr < rows_per; ++rr) {
            const int r_ = grp * rows_per + rr;
            const uint8_t* src = wide ? stage + r_ * 128 + (((c >> 3) ^ (r_ & 7)) * 16) + (c & 7) * 2
                                      : stage + r_ * 64 + (((c >> 3) ^ ((r_ >> 1) & 3)) * 16) + (c & 7) * 2;
            const float2 fv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src));
            a0 += fv.x; a1 += fv.y;
            b0 = fmaf(fv.x, fv.x, b0); b1 = fmaf(fv.y, fv.y, b1);
          }
          atomicAdd(&stat_s[rank * W + c], a0);
          atomicAdd(&stat_s[rank * W + c + 1], a1);
          atomicAdd(&stat_s[BLOCK_N + rank * W + c], b0);
          atomicAdd(&stat_s[BLOCK_N + rank * W + c + 1], b1);
        }
        continue;
      }
      if (split) {
        float* mine = wst + (size_t)blockIdx.z * slot_elems;
#pragma unroll
        for (int cp = 0; cp < BLOCK_N / kStep; ++cp) {
          uint32_t r[kStep];
#pragma unroll
          for (int h = 0; h < kStep / 32; ++h) tmem_ld_32x32_nowait(tacc + cp * kStep + h * 32, r + h * 32);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kStep / 4; ++j)
            __stcg(reinterpret_cast<float4*>(mine + ((size_t)(cp * (kStep / 4) + j) * BLOCK_M + row) * 4),
                   make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          unsigned int* tk = p.splitk_tickets + (tp * gridDim.x + tile_n);
          const unsigned int t = atomicAdd(tk, 1u);
          flag_last = (t == gridDim.z - 1);
          if (flag_last) *tk = 0u;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (!flag_last) continue;          // (one tile per CTA in split mode: the CTA is done)
        __threadfence();
      }
      // the staging buffer was last read by the TMA store (and the statistics pass) of the
      // previous tile
      if (et == 0 && iter >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int cp = 0; cp < BLOCK_N / kStep; ++cp) {
        float f[kStep];
        {
          uint32_t r[kStep];
#pragma unroll
          for (int h = 0; h < kStep / 32; ++h) tmem_ld_32x32_nowait(tacc + cp * kStep + h * 32, r + h * 32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < kStep; ++i) f[i] = __uint_as_float(r[i]);
        }
        if (split) {
          for (int z = 0; z < (int)gridDim.z; ++z) {
            if (z == (int)blockIdx.z) continue;
            const float* other = wst + (size_t)z * slot_elems;
#pragma unroll
            for (int j = 0; j < kStep / 4; ++j) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(
                  other + ((size_t)(cp * (kStep / 4) + j) * BLOCK_M + row) * 4));
              f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < kStep; ++i) {
          float v = f[i];
          if (p.bias) v += bias_s[cp * kStep + i];
          if (p.relu) v = fmaxf(v, 0.f);
          f[i] = valid ? v : 0.f;              // rows outside the image contribute nothing
        }
        const int c0 = cp * kStep;             // first column of this step inside the tile
        uint8_t* region = stage + (c0 / 64) * (BLOCK_M * 128);
        const int chunk0 = (c0 % 64) / 8;
#pragma unroll
        for (int j = 0; j < kStep / 8; ++j) {
          bf16x8 pk = pack8(f + j * 8);
          const int chunk = (chunk0 + j) ^ (row & 7);
          *reinterpret_cast<bf16x8*>(region + row * 128 + chunk * 16) = pk;
        }
      }
      // accumulator drained: hand it back to the MMA warp (one arrival per epilogue warp)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
#pragma unroll
        for (int h = 0; h < BLOCK_N / 64; ++h)
          if (tile_n * BLOCK_N + h * 64 < p.Nout)
            tma_store_4d(md_cls, stage + h * (BLOCK_M * 128), tile_n * BLOCK_N + h * 64, w0, h0, n0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      IG_STAMP(6);
      ++done_tiles;
      if (kStats && p.col_sums) {
        // per-channel sum / sum-of-squares of the (bf16-rounded) tile: every thread owns a column
        // pair and a group of rows, and keeps its partial sums in registers across the CTA's tiles
        constexpr int kPairs = BLOCK_N / 2;
        constexpr int kGroups = 128 / kPairs;
        constexpr int kRowsPer = BLOCK_M / kGroups;
        const int pair = et % kPairs, grp = et / kPairs;
        const int c = pair * 2;
        const uint8_t* region = stage + (c / 64) * (BLOCK_M * 128);
        const int cc = c % 64;
#pragma unroll 16
        for (int rr = 0; rr < kRowsPer; ++rr) {
          const int r_ = grp * kRowsPer + rr;
          const int chunk = (cc >> 3) ^ (r_ & 7);
          const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(
              region + r_ * 128 + chunk * 16 + (cc & 7) * 2);
          const float2 fv = __bfloat1622float2(v);
          s0 += fv.x; s1 += fv.y;
          q0 = fmaf(fv.x, fv.x, q0); q1 = fmaf(fv.y, fv.y, q1);
        }
      }
    }
    if (kStats && p.col_sums && done_tiles > 0) {
      // combine the row groups in shared memory, flush this CTA's sums with one vector of atomics;
      // the CTA that completes the LAST pixel tile of this n-tile finalizes its BLOCK_N channels
      // (one channel per thread: every n-tile is finalized in parallel by a different CTA)
      if (!csplit) {
        const int c = (et % (BLOCK_N / 2)) * 2;
        atomicAdd(&stat_s[c], s0);
        atomicAdd(&stat_s[c + 1], s1);
        atomicAdd(&stat_s[BLOCK_N + c], q0);
        atomicAdd(&stat_s[BLOCK_N + c + 1], q1);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* cs = p.col_sums + (size_t)(blockIdx.y % DDL_BN_REPLICAS) * 2 * p.Nout;
      // (cluster split-K: this CTA produced only its own split_w columns of the tile)
      const int i_lo = csplit ? (int)blockIdx.z * p.split_w : 0;
      const int i_hi = csplit ? i_lo + p.split_w : BLOCK_N;
      for (int i = i_lo + et; i < i_hi; i += 128) {
        const int col = tile_n * BLOCK_N + i;
        if (col < p.Nout) {
          atomicAdd(&cs[col], stat_s[i]);
          atomicAdd(&cs[p.Nout + col], stat_s[BLOCK_N + i]);
        }
      }
      IG_STAMP(7);
      if (!kSmall && p.ticket) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const unsigned int t = atomicAdd(&p.ticket[tile_n], (unsigned int)done_tiles);
          flag_last = (t + (unsigned int)done_tiles ==
                       (unsigned int)c_pix_tiles * (csplit ? gridDim.z : 1u));
          if (flag_last) p.ticket[tile_n] = 0u;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (flag_last) {
          __threadfence();
          const float inv_m = 1.f / (float)p.stat_count;
          for (int i = et; i < BLOCK_N; i += 128) {
            const int c = tile_n * BLOCK_N + i;
            if (c >= p.Nout) continue;
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int r = 0; r < DDL_BN_REPLICAS; ++r) {
              float* pp = p.col_sums + (size_t)r * 2 * p.Nout;
              s += __ldcg(pp + c);
              q += __ldcg(pp + p.Nout + c);
              pp[c] = 0.f;
              pp[p.Nout + c] = 0.f;
            }
            const float mean = s * inv_m;
            const float var = fmaxf(fmaf(-mean, mean, q * inv_m), 0.f);
            p.stat_out[c] = mean;
            p.stat_out[p.Nout + c] = rsqrtf(var + p.eps);
            if (p.running_mean) {
              const float unbiased =
                  p.stat_count > 1 ? var * ((float)p.stat_count / (float)(p.stat_count - 1)) : var;
              p.running_mean[c] = fmaf(p.momentum, mean - p.running_mean[c], p.running_mean[c]);
              p.running_var[c] = fmaf(p.momentum, unbiased - p.running_var[c], p.running_var[c]);
            }
          }
        }
      }
    }
    IG_STAMP(8);
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    IG_STAMP(9);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------ host side
// Generic 4-D bf16 tensor map: dims {C, W, H, N} (elements), strides of W / H / N in ELEMENTS,
// box {bc, bw, bh, bn} elements loaded, `step` = element stride on W and H.
static inline int make_map_4d(CUtensorMap* map, const void* ptr, int64_t C, int64_t W, int64_t H,
                              int64_t N, int64_t sW, int64_t sH, int64_t sN, int bc, int bw, int bh,
                              int bn, int step, int swizzle_bytes = 128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)(bw * step), (cuuint32_t)(bh * step),
                       (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)step, (cuuint32_t)step, 1};
  if (box[1] > 256 || box[2] > 256 || box[3] > 256) return -1;
  if ((strides[0] % 16) || (strides[1] % 16) || (strides[2] % 16)) return -1;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

static unsigned long long* g_ig_trace = nullptr;

struct IgTensor {       // NHWC view: extents and element strides (C is contiguous)
  const void* ptr;
  int64_t C, W, H, N, sW, sH, sN;
};

struct IgProblem {
  IgTensor a, d;        // activation operand, output (tiles iterate over d's W x H x N)
  int a_step;
  const void* b;        // filter matrix [b_rows][b_cols], leading dimension ldb
  int64_t b_rows, b_cols, ldb;
  int b_mn;             // 0: B[n][k] (K-major, fprop); 1: B[k][n] (MN-major, dgrad)
  int kc;               // reduction channels per tap
  int ntaps;
  short tap_dw[kIgMaxTaps], tap_dh[kIgMaxTaps];
  int tap_bcol[kIgMaxTaps];
  float* col_sums;
  const void* bias;
  int relu;
  const DdlBnFinalize* fin;
  float* splitk_ws;               // zero fp32 scratch for split-K (may be null: never split)
  unsigned int* splitk_tickets;
  int splitk_tiles;               // capacity of both, in 128x128 tiles
  // multi-class launch (stride-2 data gradient: one class per output parity): class c writes d_cls[c]
  // from taps [cls_tap0[c], cls_tap0[c+1]); class 0's tensor is also `d`
  int ncls;
  IgTensor d_cls[4];
  int cls_tap0[5];
};

template <int BLOCK_N, int kStages, bool kBMn, bool kStats, bool kSmall>
static int launch_ig(const IgProblem& q, IgemmParams& p, int sms, cudaStream_t st) {
  using L = IgSmem<BLOCK_N, kStages, kStats, kSmall>;
  CUtensorMap ma, mb, md;
  int rc;
  if ((rc = make_map_4d(&ma, q.a.ptr, q.a.C, q.a.W, q.a.H, q.a.N, q.a.sW, q.a.sH, q.a.sN, BLOCK_K,
                        p.TW, p.TH, p.TN, q.a_step)))
    return rc;
  if (kBMn) {
    if ((rc = make_map_2d(&mb, q.b, q.b_rows, q.b_cols, q.ldb, 64, 64))) return rc;
  } else {
    if ((rc = make_map_2d(&mb, q.b, q.b_rows, q.b_cols, q.ldb, BLOCK_K, BLOCK_N))) return rc;
  }
  if ((rc = make_map_4d(&md, q.d.ptr, q.d.C, q.d.W, q.d.H, q.d.N, q.d.sW, q.d.sH, q.d.sN, 64, p.TW,
                        p.TH, p.TN, 1)))
    return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_bf16_kernel<BLOCK_N, kStages, kBMn, kStats, kSmall>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.tiles_n = (int)((q.d.C + BLOCK_N - 1) / BLOCK_N);
  p.pix_tiles = p.tiles_w * p.tiles_h * (int)((q.d.N + p.TN - 1) / p.TN);
  long long total = (long long)p.pix_tiles * p.tiles_n;
  const int total_kb = p.ntaps * p.kc_blocks;
  CUtensorMap mdc[4] = {md, md, md, md};
  int max_pix_tiles = p.pix_tiles;
  p.ncls = (!kSmall && q.ncls > 1) ? q.ncls : 1;
  if (p.ncls > 1) {
    total = 0;
    max_pix_tiles = 0;
    for (int c = 0; c < p.ncls; ++c) {
      const IgTensor& d = q.d_cls[c];
      if ((rc = make_map_4d(&mdc[c], d.ptr, d.C, d.W, d.H, d.N, d.sW, d.sH, d.sN, 64, p.TW, p.TH, p.TN, 1)))
        return rc;
      IgemmParams::Cls& k = p.cls[c];
      k.OW = (int)d.W; k.OH = (int)d.H;
      k.tiles_w = (int)((d.W + p.TW - 1) / p.TW);
      k.tiles_h = (int)((d.H + p.TH - 1) / p.TH);
      k.pix_tiles = k.tiles_w * k.tiles_h * (int)((d.N + p.TN - 1) / p.TN);
      k.tap0 = q.cls_tap0[c];
      k.ntaps = q.cls_tap0[c + 1] - q.cls_tap0[c];
      k.pad = 0;
      total += (long long)k.pix_tiles * p.tiles_n;
      if (k.pix_tiles > max_pix_tiles) max_pix_tiles = k.pix_tiles;
    }
    md = mdc[0];
  }
  // Work decomposition.  Enough tiles: ONE wave of at most 2 CTAs per SM, every CTA a chunk of
  // consecutive pixel tiles (set-up, statistics flush and ticket are paid once per chunk).  Too
  // few tiles to fill the machine and a long reduction (the 14x14 / 7x7 layers): split-K over
  // gridDim.z, partial tiles meet in the fp32 workspace, the last slice runs the epilogue.
  int per = (int)((total + 2LL * sms - 1) / (2LL * sms));
  if (per > 16) per = 16;
  if (kSmall) per = 1;
  // split-K flavour: default = the slices of a tile form a thread-block cluster (co-scheduled), every
  // slice finishes its own columns and the partial sums are exchanged through L2 ("dsmem": through
  // distributed shared memory instead); "slots" = round-2's first scheme (partial tiles
  // parked in a global fp32 workspace, the last slice to arrive reduces) kept for A/B runs
  static const int split_mode = [] {
    const char* e = std::getenv("DDLB_IG_SPLIT");
    return (e && !strcmp(e, "slots")) ? 1 : (e && !strcmp(e, "off")) ? 2 : (e && !strcmp(e, "dsmem")) ? 3 : 0;
  }();
  int splits = 1;
  p.split_w = 0;
  p.split_l2 = 0;
  if (!kSmall && p.ncls == 1 && total < sms && total_kb >= 8 && BLOCK_N == 128 && split_mode != 2) {
    if (split_mode == 0 || split_mode == 3) {
      // 2 or 4 slices: enough CTAs for every SM, at most the 2-per-SM residency
      splits = (total * 4 <= 2LL * sms && total_kb >= 16) ? 4 : 2;
      for (; splits > 1; splits /= 2) {
        const int kbps = (total_kb + splits - 1) / splits;
        if ((total_kb + kbps - 1) / kbps == splits) break;          // no empty slice
      }
      if (splits > 1) {
        p.split_w = BLOCK_N / splits;
        // exchange through L2 when the workspace has a slot per (tile, slice); DDLB_IG_SPLIT=dsmem
        // forces the distributed-shared-memory exchange
        p.split_l2 = (split_mode == 0 && q.splitk_ws && total * splits <= q.splitk_tiles) ? 1 : 0;
      }
    } else if (q.splitk_ws) {
      splits = (int)((3LL * sms / 2 + total - 1) / total);
      if (splits > total_kb / 4) splits = total_kb / 4;
      if (splits > 8) splits = 8;
      while (splits > 1 && total * splits > q.splitk_tiles) --splits;   // slots available
      if (splits < 1) splits = 1;
    }
  }
  p.trace = g_ig_trace;
  if (splits > 1) per = 1;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;      // no empty slices
  if (p.split_w && BLOCK_N / splits != p.split_w) return -4;      // (cannot happen: checked above)
  p.splitk_ws = q.splitk_ws;
  p.splitk_tickets = q.splitk_tickets;
  p.tiles_per_cta = per;
  const int chunks = (max_pix_tiles + per - 1) / per;
  if (chunks > 65535 || p.tiles_n > 65535) return -1;
  if (p.ncls > 1) splits = p.ncls;          // gridDim.z = classes (kb_per_split covers every tap)
  CUtensorMap mds = md;
  if (p.split_w == 32) {
    if ((rc = make_map_4d(&mds, q.d.ptr, q.d.C, q.d.W, q.d.H, q.d.N, q.d.sW, q.d.sH, q.d.sN, 32, p.TW,
                          p.TH, p.TN, 1, /*swizzle_bytes=*/64)))
      return rc;
  }
  cudaError_t err;
  if (p.split_w) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.tiles_n, chunks, splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)L::kDynamic;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)splits;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    err = cudaLaunchKernelEx(&cfg, igemm_bf16_kernel<BLOCK_N, kStages, kBMn, kStats, kSmall>, ma, mb, md,
                             mds, mdc[1], mdc[2], mdc[3], p);
  } else {
    err = launch_pdl((igemm_bf16_kernel<BLOCK_N, kStages, kBMn, kStats, kSmall>),
                     dim3(p.tiles_n, chunks, splits), dim3(kThreads), (size_t)L::kDynamic, st, ma, mb,
                     md, mds, mdc[1], mdc[2], mdc[3], p);
  }
  if (err != cudaSuccess) return (int)err;
  err = cudaGetLastError();
  if (err != cudaSuccess) return (int)err;
  return (kStats && p.col_sums && p.ticket) ? DDL_IGEMM_FINALIZED : 0;
}

static int run_ig(const IgProblem& q, int sms, cudaStream_t st) {
  if (q.ntaps < 1 || q.ntaps > kIgMaxTaps) return -1;
  if ((q.a.C % 8) || (q.d.C % 8) || (q.ldb % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(q.a.ptr) | reinterpret_cast<uintptr_t>(q.b) |
       reinterpret_cast<uintptr_t>(q.d.ptr)) & 15)
    return -1;
  IgemmParams p{};
  p.OW = (int)q.d.W; p.OH = (int)q.d.H; p.NB = (int)q.d.N;
  p.Nout = (int)q.d.C;
  p.kc_blocks = (q.kc + BLOCK_K - 1) / BLOCK_K;
  p.a_step = q.a_step;
  p.ntaps = q.ntaps;
  memcpy(p.tap_dw, q.tap_dw, sizeof(p.tap_dw));
  memcpy(p.tap_dh, q.tap_dh, sizeof(p.tap_dh));
  memcpy(p.tap_bcol, q.tap_bcol, sizeof(p.tap_bcol));
  if (q.d.H == 1 && q.d.N == 1) {           // flat GEMM rows: no edge waste
    p.TW = 128; p.TH = 1; p.TN = 1;
  } else {
    choose_pixel_box(128, (int)q.d.N, (int)q.d.H, (int)q.d.W, &p.TW, &p.TH, &p.TN);
  }
  if (p.TW * q.a_step > 256 || p.TH * q.a_step > 256) return -1;
  p.tiles_w = (int)((q.d.W + p.TW - 1) / p.TW);
  p.tiles_h = (int)((q.d.H + p.TH - 1) / p.TH);
  const bool stats = q.col_sums != nullptr;
  p.col_sums = q.col_sums;
  p.bias = (const __nv_bfloat16*)q.bias;
  p.relu = q.relu;
  p.stat_count = (long long)q.d.W * q.d.H * q.d.N;
  if (q.fin && stats) {
    p.ticket = q.fin->ticket;
    p.stat_out = q.fin->stat_out;
    p.running_mean = q.fin->running_mean;
    p.running_var = q.fin->running_var;
    p.eps = q.fin->eps;
    p.momentum = q.fin->momentum;
  }
  const bool narrow = q.d.C <= 64;
  // variant: short reductions (<= small_max_kb k-blocks of 64) -> one tile per CTA, 3-4 CTAs per SM;
  // otherwise chunked + double-buffered accumulator (and split-K when the grid would be tiny)
  static const int small_max_kb = [] {
    const char* e = std::getenv("DDLB_IG_SMALL_MAXKB");
    return e ? atoi(e) : 4;
  }();
  const int total_kb = q.ntaps * p.kc_blocks;
  // (the one-tile variant pays off only where the statistics make the epilogue the bottleneck)
  const int v = (!stats || total_kb > small_max_kb) ? 0 : (total_kb == 1 ? 1 : 2);   // 0 chunked, 1/2 small
  if (v != 0) p.ticket = nullptr;      // sums only: the caller runs the 2 us finalize kernel
#define DDL_IG_DISPATCH(BMN, STATS)                                                          \
  do {                                                                                       \
    if (narrow) {                                                                            \
      if (v == 0) return launch_ig<64, 3, BMN, STATS, false>(q, p, sms, st);                 \
      if (v == 1) return launch_ig<64, 1, BMN, STATS, true>(q, p, sms, st);                  \
      return launch_ig<64, 2, BMN, STATS, true>(q, p, sms, st);                              \
    }                                                                                        \
    if (v == 0) return launch_ig<128, 2, BMN, STATS, false>(q, p, sms, st);                  \
    if (v == 1) return launch_ig<128, 1, BMN, STATS, true>(q, p, sms, st);                   \
    return launch_ig<128, 2, BMN, STATS, true>(q, p, sms, st);                               \
  } while (0)
  if (q.b_mn) {
    if (stats) return -1;
    DDL_IG_DISPATCH(true, false);
  }
  if (stats) DDL_IG_DISPATCH(false, true);
  DDL_IG_DISPATCH(false, false);
#undef DDL_IG_DISPATCH
}

static inline IgTensor dense_nhwc(const void* ptr, int64_t N, int64_t H, int64_t W, int64_t C) {
  return IgTensor{ptr, C, W, H, N, C, W * C, H * W * C};
}

}  // namespace gemm
}  // namespace ddl

using ddl::gemm::IgProblem;
using ddl::gemm::IgTensor;

// D[M, N] = A[M, K] op(B) with optional bias/ReLU/BatchNorm statistics (+finalize).
// b_mn_major = 0: B is [N, K] (K contiguous);  1: B is [K, N] (N contiguous).
extern "C" int ddl_igemm_gemm(const void* A, const void* B, void* D, int M, int N, int K, int lda,
                              int ldb, int ldd, int b_mn_major, float* col_sums, const void* bias,
                              int relu, const DdlBnFinalize* fin, float* splitk_ws,
                              unsigned int* splitk_tickets, int splitk_tiles, int sms,
                              cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldd % 8) || (K % 8) || (N % 8)) return -1;
  IgProblem q{};
  q.a = IgTensor{A, K, M, 1, 1, lda, (int64_t)lda * M, (int64_t)lda * M};
  q.d = IgTensor{D, N, M, 1, 1, ldd, (int64_t)ldd * M, (int64_t)ldd * M};
  q.a_step = 1;
  q.b = B;
  q.b_mn = b_mn_major;
  q.b_rows = b_mn_major ? K : N;
  q.b_cols = b_mn_major ? N : K;
  q.ldb = ldb;
  q.kc = K;
  q.ntaps = 1;
  q.col_sums = col_sums; q.bias = bias; q.relu = relu; q.fin = fin;
  q.splitk_ws = splitk_ws; q.splitk_tickets = splitk_tickets; q.splitk_tiles = splitk_tiles;
  return ddl::gemm::run_ig(q, sms, st);
}

// y = conv2d(x, w) (+bias, ReLU, BatchNorm statistics).  x: [NB,H,W,Cin], w: [Cout,KH,KW,Cin] (KRSC),
// y: [NB,OH,OW,Cout], all bf16.  stride 1 or 2, dilation 1, groups 1.  -1: unsupported shape.
extern "C" int ddl_igemm_conv_fprop(const void* x, const void* w, void* y, int NB, int H, int W,
                                    int Cin, int Cout, int KH, int KW, int pad_h, int pad_w,
                                    int stride, float* col_sums, const void* bias, int relu,
                                    const DdlBnFinalize* fin, float* splitk_ws,
                                    unsigned int* splitk_tickets, int splitk_tiles, int sms,
                                    cudaStream_t st) {
  if (NB <= 0 || H <= 0 || W <= 0) return 0;
  if ((Cin % 8) || (Cout % 8) || KH < 1 || KW < 1 || KH * KW > ddl::gemm::kIgMaxTaps) return -1;
  if (stride != 1 && stride != 2) return -1;
  const int OH = (H + 2 * pad_h - KH) / stride + 1, OW = (W + 2 * pad_w - KW) / stride + 1;
  if (OH <= 0 || OW <= 0) return -1;
  IgProblem q{};
  const int64_t kdim = (int64_t)KH * KW * Cin;
  if (KH == 1 && KW == 1 && stride == 1 && pad_h == 0 && pad_w == 0) {
    const int64_t M = (int64_t)NB * H * W;
    q.a = IgTensor{x, Cin, M, 1, 1, Cin, M * Cin, M * Cin};
    q.d = IgTensor{y, Cout, M, 1, 1, Cout, M * Cout, M * Cout};
  } else {
    q.a = ddl::gemm::dense_nhwc(x, NB, H, W, Cin);
    q.d = ddl::gemm::dense_nhwc(y, NB, OH, OW, Cout);
  }
  q.a_step = stride;
  q.b = w; q.b_mn = 0; q.b_rows = Cout; q.b_cols = kdim; q.ldb = kdim;
  q.kc = Cin;
  q.ntaps = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int t = kh * KW + kw;
      q.tap_dw[t] = (short)(kw - pad_w);
      q.tap_dh[t] = (short)(kh - pad_h);
      q.tap_bcol[t] = t * Cin;
    }
  q.col_sums = col_sums; q.bias = bias; q.relu = relu; q.fin = fin;
  q.splitk_ws = splitk_ws; q.splitk_tickets = splitk_tickets; q.splitk_tiles = splitk_tiles;
  return ddl::gemm::run_ig(q, sms, st);
}

// dx = conv2d_backward_data(dy, w).  dy: [NB,OH,OW,Cout], w: [Cout,KH,KW,Cin] (the forward filter,
// untouched), dx: [NB,H,W,Cin] — fully written.  stride 1 or 2.
extern "C" int ddl_igemm_conv_dgrad(const void* dy, const void* w, void* dx, int NB, int H, int W,
                                    int Cin, int Cout, int KH, int KW, int pad_h, int pad_w,
                                    int stride, float* splitk_ws, unsigned int* splitk_tickets,
                                    int splitk_tiles, int sms, cudaStream_t st) {
  if (NB <= 0 || H <= 0 || W <= 0) return 0;
  if ((Cin % 8) || (Cout % 8) || KH < 1 || KW < 1 || KH * KW > ddl::gemm::kIgMaxTaps) return -1;
  if (stride != 1 && stride != 2) return -1;
  const int OH = (H + 2 * pad_h - KH) / stride + 1, OW = (W + 2 * pad_w - KW) / stride + 1;
  if (OH <= 0 || OW <= 0) return -1;
  const int64_t kdim = (int64_t)KH * KW * Cin;
  IgProblem q{};
  q.splitk_ws = splitk_ws; q.splitk_tickets = splitk_tickets; q.splitk_tiles = splitk_tiles;
  q.b = w; q.b_mn = 1; q.b_rows = Cout; q.b_cols = kdim; q.ldb = kdim;
  q.kc = Cout;
  q.a_step = 1;
  if (stride == 1) {
    if (KH == 1 && KW == 1 && pad_h == 0 && pad_w == 0) {
      const int64_t M = (int64_t)NB * H * W;
      q.a = IgTensor{dy, Cout, M, 1, 1, Cout, M * Cout, M * Cout};
      q.d = IgTensor{dx, Cin, M, 1, 1, Cin, M * Cin, M * Cin};
    } else {
      q.a = ddl::gemm::dense_nhwc(dy, NB, OH, OW, Cout);
      q.d = ddl::gemm::dense_nhwc(dx, NB, H, W, Cin);
    }
    q.ntaps = KH * KW;
    for (int kh = 0; kh < KH; ++kh)
      for (int kw = 0; kw < KW; ++kw) {
        const int t = kh * KW + kw;
        q.tap_dw[t] = (short)(pad_w - kw);
        q.tap_dh[t] = (short)(pad_h - kh);
        q.tap_bcol[t] = t * Cin;
      }
    return ddl::gemm::run_ig(q, sms, st);
  }
  // stride 2: one stride-1 problem per output parity class (a, b), stored through a strided view
  bool need_zero = false;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      int n = 0;
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KW; ++kw)
          if (((a + pad_h - kh) % 2 == 0) && ((b + pad_w - kw) % 2 == 0)) ++n;
      if (n == 0 && (H - a + 1) / 2 > 0 && (W - b + 1) / 2 > 0) need_zero = true;
    }
  if (need_zero) {
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)NB * H * W * Cin * 2, st);
    if (e != cudaSuccess) return (int)e;
  }
  q.a = ddl::gemm::dense_nhwc(dy, NB, OH, OW, Cout);
  // the parity classes that have taps, largest first (class 0 fixes the pixel box): ONE launch with
  // gridDim.z = classes when there are several (DDLB_IG_S2_MULTI=0: one launch per class)
  static const bool multi_ok = [] {
    const char* e = std::getenv("DDLB_IG_S2_MULTI");
    return !(e && e[0] == '0');
  }();
  int ncls = 0, ntaps_all = 0;
  q.cls_tap0[0] = 0;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const int Ha = (H - a + 1) / 2, Wb = (W - b + 1) / 2;
      if (Ha <= 0 || Wb <= 0) continue;
      int n = 0;
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KW; ++kw) {
          const int nh = a + pad_h - kh, nw = b + pad_w - kw;
          if ((nh % 2) || (nw % 2)) continue;
          q.tap_dh[ntaps_all + n] = (short)(nh / 2);
          q.tap_dw[ntaps_all + n] = (short)(nw / 2);
          q.tap_bcol[ntaps_all + n] = (kh * KW + kw) * Cin;
          ++n;
        }
      if (n == 0) continue;
      const __nv_bfloat16* base = (const __nv_bfloat16*)dx + ((int64_t)a * W + b) * Cin;
      q.d_cls[ncls] = IgTensor{base, Cin, Wb, Ha, NB, 2 * (int64_t)Cin, 2 * (int64_t)W * Cin,
                               (int64_t)H * W * Cin};
      ntaps_all += n;
      q.cls_tap0[++ncls] = ntaps_all;
    }
  if (ncls == 0) return 0;
  if (ncls > 1 && multi_ok) {
    q.ncls = ncls;
    q.d = q.d_cls[0];
    q.ntaps = q.cls_tap0[1];                   // class 0's taps (variant choice / k-loop bound of class 0)
    return ddl::gemm::run_ig(q, sms, st);
  }
  IgProblem one = q;
  one.ncls = 1;
  for (int c = 0; c < ncls; ++c) {
    one.d = q.d_cls[c];
    one.ntaps = q.cls_tap0[c + 1] - q.cls_tap0[c];
    for (int t = 0; t < one.ntaps; ++t) {
      one.tap_dh[t] = q.tap_dh[q.cls_tap0[c] + t];
      one.tap_dw[t] = q.tap_dw[q.cls_tap0[c] + t];
      one.tap_bcol[t] = q.tap_bcol[q.cls_tap0[c] + t];
    }
    const int rc = ddl::gemm::run_ig(one, sms, st);
    if (rc) return rc;
  }
  return 0;
}

// debug: per-CTA stage timestamps of the next igemm launches go to `buf` ([ctas][16] uint64; null: off)
extern "C" void ddl_igemm_set_trace(unsigned long long* buf) { ddl::gemm::g_ig_trace = buf; }
