// Single-pass BatchNorm forward / backward on thread-block CLUSTERS.
//
// The two-kernel forms in bn.cu (stats -> apply, bwd_reduce -> bwd_apply) read every operand
// twice and pay two launches, a global-atomic reduction and a last-CTA finalize tail per layer.
// At the reference's batch size (32/GPU) 40 of ResNet-50's 53 BatchNorm layers have activations
// of 1.6-13 MB: small enough that a whole CHANNEL SLAB (32-128 channels x all N*H*W rows) fits in
// the shared memory of one cluster of <= 16 CTAs (16 x 227 KB).  So here
//
//   * a cluster owns a channel slab, its CTAs split the rows;
//   * phase 1 streams the operands ONCE from global memory, parks them in shared memory and
//     accumulates the per-channel sums in registers;
//   * the sums are combined inside the CTA (shuffles + one smem pass) and across the cluster
//     through DISTRIBUTED SHARED MEMORY (every CTA reads its peers' partials) — no global
//     atomics, no workspace, no ticket, no finalize launch;
//   * phase 2 applies the per-channel coefficients to the parked tiles and writes the result.
//
// Global traffic: backward = read dy (+dy2), 1-bit mask, x ; write dx  (vs. 2x the reads);
//                 forward  = read x (+res) ; write y (+mask)          (vs. x read twice).
// Layers whose slab does not fit (N*H*W = 100352 at 56x56) keep the two-kernel path.
#include <cooperative_groups.h>

#include "api.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace ddl {

constexpr int kClThreads = 512;
constexpr int kClWarps = kClThreads / 32;
constexpr int kActBitmaskFlag = 0x100;

__device__ __forceinline__ float cl_act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
__device__ __forceinline__ bool cl_act_on(float y, int act) {
  if (act == 1) return y > 0.f;
  if (act == 2) return y > 0.f && y < 6.f;
  return true;
}

struct ClusterPlan {
  int slab_vecs;      // 16-byte channel vectors per slab (blockDim.x): 4, 8 or 16
  int cluster;        // CTAs per cluster (row splits)
  int slabs;          // channel slabs (clusters)
  int rows_per_cta;
  size_t smem;        // dynamic shared memory per CTA
  bool ok;
};

// shared memory: [tiles][rows_per_cta][slab_vecs] 16-byte vectors, then the reduction scratch
__host__ __device__ inline size_t cl_scratch_bytes(int slab_vecs) {
  // part[16*slab_vecs] + tot[16*slab_vecs] + red[kClWarps][slab_vecs][16] + gather[16][16*slab_vecs]
  return (size_t)(2 * 16 * slab_vecs + kClWarps * slab_vecs * 16 + 16 * 16 * slab_vecs) * sizeof(float);
}

static int max_active_clusters(int cluster, size_t smem);

static ClusterPlan plan_cluster(int64_t M, int C, int tiles, int sms, int max_cluster) {
  ClusterPlan best{};
  best.ok = false;
  const size_t cap = 227 * 1024 - 1024;
  const int cv = C / 8;
  double best_score = -1.0;
  for (int sv : {16, 8, 4}) {
    if (sv > cv && sv != 4) continue;
    const int slabs = (cv + sv - 1) / sv;
    for (int cl : {1, 2, 4, 8, 16}) {
      if (cl > max_cluster) continue;
      const int64_t rows = (M + cl - 1) / cl;
      const size_t smem = (size_t)rows * sv * 16 * tiles + (((size_t)rows * sv + 15) & ~(size_t)15) +
                          cl_scratch_bytes(sv);
      if (smem > cap) continue;
      if (rows < 1) continue;
      // one wave: a second round of clusters would double the kernel time (measured)
      if (slabs > max_active_clusters(cl, smem)) continue;
      const int ctas = slabs * cl;
      // fill the machine (but not several waves of 1-CTA/SM clusters), then prefer wide slabs
      double score = (double)(ctas < sms ? ctas : sms) / sms;
      if (ctas > 2 * sms) score -= 0.25 * ((double)ctas / sms - 2.0);
      score += 0.02 * (sv / 4) - 0.01 * cl / 16.0;
      if (rows * sv < 2 * kClThreads) score -= 0.3;       // under two vectors per thread: overhead-bound
      if (score > best_score) {
        best_score = score;
        best.slab_vecs = sv;
        best.cluster = cl;
        best.slabs = slabs;
        best.rows_per_cta = (int)rows;
        best.smem = smem;
        best.ok = true;
      }
    }
  }
  return best;
}

// Reduce NV per-thread values over all row lanes of the CTA, then over the cluster.
// On return tot[k * slab_ch .. ] (k = 0 .. NV/8-1 groups of channel-indexed sums) holds the
// cluster-wide totals in every CTA.  Layout of part/tot: [NV/8][slab_vecs*8].
template <int NV>
__device__ __forceinline__ void cluster_reduce(float* v, float* part, float* tot, float* red,
                                               cg::cluster_group& cluster) {
  const int sv = blockDim.x;
  const int tid = threadIdx.y * sv + threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // 1. lanes of a warp that share threadIdx.x (stride sv) -> lane < sv holds the warp sum
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float a = v[k];
    for (int off = 16; off >= sv; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    v[k] = a;
  }
  if (lane < sv) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[(warp * sv + lane) * 16 + k] = v[k];
  }
  __syncthreads();
  // 2. across the warps of the CTA
  const int slab_ch = sv * 8;
  const int nout = (NV / 8) * slab_ch;
  for (int o = tid; o < nout; o += kClThreads) {
    const int grp = o / slab_ch, ch = o % slab_ch;
    const int tx = ch >> 3, kk = ch & 7;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kClWarps; ++w) a += red[(w * sv + tx) * 16 + grp * 8 + kk];
    part[o] = a;
  }
  __syncthreads();
  // 3. across the CTAs of the cluster: every CTA PUSHES its partial into slot [rank] of every
  // peer's gather buffer (remote shared-memory stores are fire-and-forget), one cluster barrier,
  // then a local sum.  (Pulling the peers' partials instead costs a dependent chain of remote
  // loads per output and a second barrier.)
  const unsigned n = cluster.num_blocks();
  const unsigned me = cluster.block_rank();
  float* gather = red + kClWarps * sv * 16;
  // completes the split barrier every CTA arrived at on entry: all peers are running, so their
  // shared memory may be written
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  for (int o = tid; o < nout; o += kClThreads) {
    const float a = part[o];
    for (unsigned r = 0; r < n; ++r) cluster.map_shared_rank(gather, r)[me * nout + o] = a;
  }
  cluster.sync();
  for (int o = tid; o < nout; o += kClThreads) {
    float a = 0.f;
    for (unsigned r = 0; r < n; ++r) a += gather[r * nout + o];
    tot[o] = a;
  }
  __syncthreads();
}

// 16-byte (L2-only) and 4/8/16-byte asynchronous global->shared copies: the whole CTA tile is
// requested up front, so the memory system sees every load of the kernel at once instead of a
// dependent load->use chain per row (first version, ncu: 7 GB/s per SM with 16 resident warps).
__device__ __forceinline__ void cp_async16(void* smem, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem)),
               "l"(gptr)
               : "memory");
}
__device__ __forceinline__ void cp_async_small(void* smem, const void* gptr, int bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
  if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gptr) : "memory");
  else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gptr) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------- backward
// smem: tile_d [rows][sv] int4 (dy, overwritten with dz) | tile_x | [tile_e (dy2)] | mask rows
// [rows][sv] bytes | reduction scratch.  `tiles` (2 or 3) is decided by the host plan.
template <typename PT, typename GT>
__global__ void __launch_bounds__(kClThreads, 2)
bn_bwd_cluster_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                      const uint8_t* __restrict__ mask, const __nv_bfloat16* __restrict__ x,
                      __nv_bfloat16* __restrict__ dz_out, __nv_bfloat16* __restrict__ dx,
                      const float* __restrict__ stat, const PT* __restrict__ gamma,
                      GT* __restrict__ dgamma, GT* __restrict__ dbeta, int accumulate, int64_t M,
                      int C, int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t cl_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // waited on in cluster_reduce
  pdl_enter();
  const int sv = blockDim.x, rl = blockDim.y;
  const int cv_total = C >> 3;
  const int slab = blockIdx.x / cluster.num_blocks();
  const int rank = cluster.block_rank();
  const int cv = slab * sv + threadIdx.x;
  const bool live = cv < cv_total;
  const size_t tile_elems = (size_t)rows_per_cta * sv;
  int4* tile_d = reinterpret_cast<int4*>(cl_smem);
  int4* tile_x = tile_d + tile_elems;
  int4* tile_e = tile_x + tile_elems;                       // only if dy2
  uint8_t* tile_m = reinterpret_cast<uint8_t*>(tile_x + (dy2 ? 2 : 1) * tile_elems);
  float* part = reinterpret_cast<float*>(tile_m + ((tile_elems + 15) & ~(size_t)15));
  float* tot = part + 16 * sv;
  float* red = tot + 16 * sv;

  const int64_t r0 = (int64_t)rank * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  // the mask row of a slab is `sv` contiguous bytes: one small async copy per row when aligned
  const bool mask_async = mask && (cv_total % sv == 0);
  // ---- phase 0: request the whole tile
  if (live) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      const int64_t o0 = r * C + (int64_t)cv * 8;
      const size_t sidx = (size_t)(r - r0) * sv + threadIdx.x;
      cp_async16(&tile_d[sidx], dy + o0);
      cp_async16(&tile_x[sidx], x + o0);
      if (dy2) cp_async16(&tile_e[sidx], dy2 + o0);
      if (mask_async && threadIdx.x == 0)
        cp_async_small(tile_m + (size_t)(r - r0) * sv, mask + r * cv_total + (int64_t)slab * sv, sv);
    }
  }
  float v[16], mean[8];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  if (live) {
#pragma unroll
    for (int k = 0; k < 8; ++k) mean[k] = stat[cv * 8 + k];
  }
  cp_async_wait_all();
  if (mask_async) __syncthreads();     // mask bytes were copied by the threadIdx.x == 0 lanes
  // ---- phase 1: dz = mask * (dy + dy2), sums
  if (live) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      const size_t sidx = (size_t)(r - r0) * sv + threadIdx.x;
      int4 rd = tile_d[sidx], rx = tile_x[sidx];
      float d0[8], x0[8];
      unpack8(*reinterpret_cast<bf16x8*>(&rd), d0);
      if (dy2) {
        int4 r2 = tile_e[sidx];
        float e0[8];
        unpack8(*reinterpret_cast<bf16x8*>(&r2), e0);
#pragma unroll
        for (int k = 0; k < 8; ++k) d0[k] += e0[k];
      }
      if (mask) {
        const unsigned mb = mask_async ? tile_m[sidx] : mask[r * cv_total + cv];
#pragma unroll
        for (int k = 0; k < 8; ++k) d0[k] = (mb >> k) & 1u ? d0[k] : 0.f;
      }
      bf16x8 pd = pack8(d0);
      if (dz_out) *reinterpret_cast<bf16x8*>(dz_out + r * C + (int64_t)cv * 8) = pd;
      if (dy2 || mask) tile_d[sidx] = *reinterpret_cast<int4*>(&pd);
      unpack8(*reinterpret_cast<bf16x8*>(&rx), x0);
      unpack8(pd, d0);    // the sums see the same rounded dz the apply phase will use
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] += d0[k];
        v[8 + k] = fmaf(d0[k], x0[k] - mean[k], v[8 + k]);
      }
    }
  }
  cluster_reduce<16>(v, part, tot, red, cluster);
  if (!live) return;
  // ---- phase 2: dx = A*dz + Bx*x + K
  const int slab_ch = sv * 8;
  float A[8], Bx[8], K[8];
  const float inv_m = 1.f / (float)M;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cv * 8 + k;
    const float istd = stat[C + c];
    const float sdz = tot[threadIdx.x * 8 + k];
    const float sdzx = tot[slab_ch + threadIdx.x * 8 + k] * istd;
    const float c0 = to_f(gamma[c]) * istd;
    const float c1 = sdz * inv_m, c2 = sdzx * inv_m;
    A[k] = c0;
    Bx[k] = -c0 * istd * c2;
    K[k] = c0 * (mean[k] * istd * c2 - c1);
    if (rank == 0 && threadIdx.y == 0) {
      if (accumulate) {
        dgamma[c] = from_f<GT>(to_f(dgamma[c]) + sdzx);
        dbeta[c] = from_f<GT>(to_f(dbeta[c]) + sdz);
      } else {
        dgamma[c] = from_f<GT>(sdzx);
        dbeta[c] = from_f<GT>(sdz);
      }
    }
  }
  for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
    const size_t sidx = (size_t)(r - r0) * sv + threadIdx.x;
    int4 rd = tile_d[sidx], rx = tile_x[sidx];
    float d0[8], x0[8];
    unpack8(*reinterpret_cast<bf16x8*>(&rd), d0);
    unpack8(*reinterpret_cast<bf16x8*>(&rx), x0);
#pragma unroll
    for (int k = 0; k < 8; ++k) d0[k] = fmaf(A[k], d0[k], fmaf(Bx[k], x0[k], K[k]));
    *reinterpret_cast<bf16x8*>(dx + r * C + (int64_t)cv * 8) = pack8(d0);
  }
}

// ----------------------------------------------------------------------------------- forward
// smem: tile_x | [tile_r (res)] | reduction scratch
template <typename PT>
__global__ void __launch_bounds__(kClThreads, 2)
bn_fwd_cluster_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                      __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ mask_out,
                      float* __restrict__ stat_out, float* __restrict__ running_mean,
                      float* __restrict__ running_var, const PT* __restrict__ gamma,
                      const PT* __restrict__ beta, int64_t M, int C, int rows_per_cta, float eps,
                      float momentum, int act) {
  extern __shared__ __align__(16) uint8_t cl_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // waited on in cluster_reduce
  pdl_enter();
  const int sv = blockDim.x, rl = blockDim.y;
  const int cv_total = C >> 3;
  const int slab = blockIdx.x / cluster.num_blocks();
  const int rank = cluster.block_rank();
  const int cv = slab * sv + threadIdx.x;
  const bool live = cv < cv_total;
  const size_t tile_elems = (size_t)rows_per_cta * sv;
  int4* tile_x = reinterpret_cast<int4*>(cl_smem);
  int4* tile_r = tile_x + tile_elems;
  float* part = reinterpret_cast<float*>(tile_x + (res ? 2 : 1) * tile_elems);
  float* tot = part + 16 * sv;
  float* red = tot + 16 * sv;
  const int64_t r0 = (int64_t)rank * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  if (live) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      const int64_t o0 = r * C + (int64_t)cv * 8;
      const size_t sidx = (size_t)(r - r0) * sv + threadIdx.x;
      cp_async16(&tile_x[sidx], x + o0);
      if (res) cp_async16(&tile_r[sidx], res + o0);
    }
  }
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  cp_async_wait_all();
  if (live) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      int4 a = tile_x[(size_t)(r - r0) * sv + threadIdx.x];
      float f0[8];
      unpack8(*reinterpret_cast<bf16x8*>(&a), f0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] += f0[k];
        v[8 + k] = fmaf(f0[k], f0[k], v[8 + k]);
      }
    }
  }
  cluster_reduce<16>(v, part, tot, red, cluster);
  if (!live) return;
  const int slab_ch = sv * 8;
  const float inv_m = 1.f / (float)M;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cv * 8 + k;
    const float mean = tot[threadIdx.x * 8 + k] * inv_m;
    float var = fmaf(-mean, mean, tot[slab_ch + threadIdx.x * 8 + k] * inv_m);
    var = fmaxf(var, 0.f);
    const float istd = rsqrtf(var + eps);
    sc[k] = to_f(gamma[c]) * istd;
    sh[k] = fmaf(-mean, sc[k], to_f(beta[c]));
    if (rank == 0 && threadIdx.y == 0) {
      stat_out[c] = mean;
      stat_out[C + c] = istd;
      if (running_mean) {
        const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
        running_mean[c] = fmaf(momentum, mean - running_mean[c], running_mean[c]);
        running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
      }
    }
  }
  for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
    const int64_t o0 = r * C + (int64_t)cv * 8;
    const size_t sidx = (size_t)(r - r0) * sv + threadIdx.x;
    int4 a = tile_x[sidx];
    float f[8];
    unpack8(*reinterpret_cast<bf16x8*>(&a), f);
    if (res) {
      int4 rr = tile_r[sidx];
      float g[8];
      unpack8(*reinterpret_cast<bf16x8*>(&rr), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = cl_act_fwd(fmaf(f[k], sc[k], sh[k]) + g[k], act);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = cl_act_fwd(fmaf(f[k], sc[k], sh[k]), act);
    }
    const bf16x8 p = pack8(f);
    *reinterpret_cast<bf16x8*>(y + o0) = p;
    if (mask_out) {
      float q[8];
      unpack8(p, q);
      unsigned m = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) m |= (cl_act_on(q[k], act) ? 1u : 0u) << k;
      mask_out[r * cv_total + cv] = (uint8_t)m;
    }
  }
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kernel)(KArgs...), const ClusterPlan& p, cudaStream_t st,
                                  Args... args) {
  static thread_local void* configured[64];
  static thread_local int n_configured = 0;
  bool seen = false;
  for (int i = 0; i < n_configured; ++i) seen |= configured[i] == (void*)kernel;
  if (!seen) {
    cudaError_t e = cudaFuncSetAttribute((const void*)kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    if (n_configured < 64) configured[n_configured++] = (void*)kernel;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.slabs * p.cluster);
  cfg.blockDim = dim3(p.slab_vecs, kClThreads / p.slab_vecs);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// How many clusters of `cluster` CTAs with `smem` bytes each the device can hold at once
// (cudaOccupancyMaxActiveClusters on the backward kernel; cached per (cluster, smem / 8 KB)).
static int max_active_clusters(int cluster, size_t smem) {
  static int cache[5][32];
  static bool init = false;
  if (!init) {
    for (auto& row : cache) for (int& v : row) v = -1;
    init = true;
  }
  int ci = cluster == 1 ? 0 : cluster == 2 ? 1 : cluster == 4 ? 2 : cluster == 8 ? 3 : 4;
  int si = (int)((smem + 8191) / 8192);
  if (si > 31) si = 31;
  if (cache[ci][si] >= 0) return cache[ci][si];
  auto* k = bn_bwd_cluster_kernel<__nv_bfloat16, __nv_bfloat16>;
  cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster * 64);
  cfg.blockDim = dim3(4, kClThreads / 4);
  cfg.dynamicSmemBytes = (size_t)si * 8192 > 227 * 1024 - 1024 ? 227 * 1024 - 1024 : (size_t)si * 8192;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (const void*)k, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cache[ci][si] = n;
  return n;
}

// Largest usable cluster: 16 (non-portable) if the device can co-schedule one such cluster of
// 227 KB CTAs inside a GPC, otherwise the portable 8.
static int max_cluster_size() {
  static const int value = [] {
    auto* k = bn_bwd_cluster_kernel<__nv_bfloat16, __nv_bfloat16>;
    if (cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
            cudaSuccess) {
      cudaGetLastError();
      return 8;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(4, kClThreads / 4);
    cfg.dynamicSmemBytes = 226 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)k, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      return 8;
    }
    return 16;
  }();
  return value;
}

}  // namespace ddl

using namespace ddl;
typedef __nv_bfloat16 bf16;

// Returns kNotEligible when the layer does not fit (caller uses the two-kernel path), 0 on
// success, a cudaError_t otherwise.
static constexpr int kNotEligible = -12345;
extern "C" int ddl_bn_bwd_fused(const void* dy, const void* dy2, const void* ymask, const void* x,
                                void* dz_out, void* dx, const float* stat, const void* gamma,
                                int param_is_bf16, void* dgamma, void* dbeta, int grad_is_bf16,
                                int accumulate, int64_t M, int C, int act_flags, int sms,
                                cudaStream_t st) {
  const int act = act_flags & 0xff;
  // the activation derivative must come as the 1-bit mask (or there is no activation)
  if (C % 8 || (act != 0 && !(act_flags & kActBitmaskFlag))) return kNotEligible;
  ClusterPlan p = plan_cluster(M, C, dy2 ? 3 : 2, sms, max_cluster_size());
  if (!p.ok) return kNotEligible;
  const uint8_t* mask = act != 0 ? (const uint8_t*)ymask : nullptr;
  cudaError_t e;
#define L(PT, GT)                                                                                 \
  e = launch_cluster(bn_bwd_cluster_kernel<PT, GT>, p, st, (const bf16*)dy, (const bf16*)dy2,    \
                     mask, (const bf16*)x, (bf16*)dz_out, (bf16*)dx, stat, (const PT*)gamma,     \
                     (GT*)dgamma, (GT*)dbeta, accumulate, M, C, p.rows_per_cta)
  if (param_is_bf16 && grad_is_bf16) L(bf16, bf16);
  else if (param_is_bf16) L(bf16, float);
  else if (grad_is_bf16) L(float, bf16);
  else L(float, float);
#undef L
  return (int)e;
}

extern "C" int ddl_bn_fwd_fused(const void* x, const void* res, void* y, void* mask_out,
                                float* stat_out, float* running_mean, float* running_var,
                                const void* gamma, const void* beta, int param_is_bf16, int64_t M,
                                int C, float eps, float momentum, int act, int sms,
                                cudaStream_t st) {
  if (C % 8) return kNotEligible;
  ClusterPlan p = plan_cluster(M, C, res ? 2 : 1, sms, max_cluster_size());
  if (!p.ok) return kNotEligible;
  cudaError_t e;
  if (param_is_bf16)
    e = launch_cluster(bn_fwd_cluster_kernel<bf16>, p, st, (const bf16*)x, (const bf16*)res,
                       (bf16*)y, (uint8_t*)mask_out, stat_out, running_mean, running_var,
                       (const bf16*)gamma, (const bf16*)beta, M, C, p.rows_per_cta, eps, momentum,
                       act);
  else
    e = launch_cluster(bn_fwd_cluster_kernel<float>, p, st, (const bf16*)x, (const bf16*)res,
                       (bf16*)y, (uint8_t*)mask_out, stat_out, running_mean, running_var,
                       (const float*)gamma, (const float*)beta, M, C, p.rows_per_cta, eps, momentum,
                       act);
  return (int)e;
}

// Plan introspection for tests / docs: returns slab_vecs | cluster << 8 | ok << 16.
extern "C" int ddl_bn_cluster_plan(int64_t M, int C, int tiles, int sms, int* rows_per_cta,
                                   int64_t* smem) {
  ClusterPlan p = plan_cluster(M, C, tiles, sms, max_cluster_size());
  if (rows_per_cta) *rows_per_cta = p.rows_per_cta;
  if (smem) *smem = (int64_t)p.smem;
  return p.ok ? (p.slab_vecs | (p.cluster << 8) | (1 << 16)) : 0;
}
