// Gradient all-reduce over NVLink/NVSwitch peer memory, FUSED with the SGD-momentum
// update — the data-parallel hot path.
//
// Reference behaviour being replaced (K1/K4 in SURVEY.md §2.9): Horovod enqueues one
// NCCL ring all-reduce per fused <=64 MB buffer on a side stream, synchronises all
// handles in optimizer.step(), divides by world size and only then runs torch.optim.SGD
// as a separate per-tensor pass (benchmark/imagenet/imagenet_horovod.py:135-139,182);
// PipeDream's replicated stages do the same through DistributedDataParallel
// (pipedream-fork/runtime/runtime.py:259-263), on Gloo for hybrid configs.
//
// Here every rank owns 1/W of each gradient bucket.  One kernel per bucket:
//   barrier-in  : "my gradients for this bucket are final" (per-CTA flags in peer memory)
//   reduce      : CTA i of rank r pulls portion i of slice r from ALL ranks' gradient
//                 buffers with 16-byte NVLink loads and sums them in fp32 registers
//                 (reduce-scatter by loads; bf16 on the wire, fp32 accumulate);
//                 NVLS variant: one multimem.ld_reduce does the sum inside the switch
//   update      : SGD(momentum, weight decay) on the rank's fp32 master + momentum SHARD
//                 (optimizer state is sharded 1/W per rank; the gradient never makes a
//                 second trip through HBM)
//   broadcast   : the new bf16 weights of the slice are stored into EVERY rank's weight
//                 buffer (all-gather by stores; NVLS variant: one multimem.st)
//   barrier-out : "my slice has landed everywhere"
//
// Buckets are launched on a communication stream as backward finishes them, so the
// transfer overlaps the remaining wgrad/dgrad GEMMs.  CTA i only ever synchronises with
// CTA i of the other ranks, so no grid-wide sync is needed, and the epoch counters live in
// device memory so the launch sequence replays from a captured CUDA graph.
#include "common.cuh"
#include "api.h"

namespace ddl {

constexpr int kMaxRanks = 8;   // one NVSwitch domain of 8 GPUs
constexpr int kArThreads = 512;

struct PeerPtrs {
  void* p[kMaxRanks];
};

// flags layout on every rank: [2 phases][max_blocks][kMaxRanks] uint32
__device__ __forceinline__ uint32_t* flag_slot(uint32_t* base, int phase, int max_blocks,
                                               int block, int rank) {
  return base + ((size_t)phase * max_blocks + block) * kMaxRanks + rank;
}

// All threads call.  Thread r (< world) publishes `epoch` to rank r and waits for rank r's.
// status[0] is set non-zero on a timeout so a broken peer cannot hang the GPU forever.
__device__ __forceinline__ void peer_barrier(const PeerPtrs& flags, int rank, int world, int phase,
                                             int max_blocks, uint32_t epoch, int* status) {
  __syncthreads();  // every thread's prior global/peer stores are issued
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    __threadfence_system();
    st_release_sys(flag_slot((uint32_t*)flags.p[r], phase, max_blocks, blockIdx.x, rank), epoch);
    const uint32_t* mine = flag_slot((uint32_t*)flags.p[rank], phase, max_blocks, blockIdx.x, r);
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
        atomicExch(status, 1 + phase);
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int4 ld_peer16(const void* p) {
  int4 r;
  asm volatile("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_peer16(void* p, const int4& v) {
  asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// NVLS: reduce 8 bf16 across all ranks inside the switch (fp32 accumulate)
__device__ __forceinline__ int4 mc_ld_reduce_bf16x8(const void* mc) {
  int4 r;
  asm volatile(
      "multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(mc)
      : "memory");
  return r;
}
__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const void* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st16(void* mc, const int4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc),
               "f"(__int_as_float(v.x)), "f"(__int_as_float(v.y)), "f"(__int_as_float(v.z)),
               "f"(__int_as_float(v.w))
               : "memory");
}

struct ArSgdArgs {
  PeerPtrs grad;     // per-rank base of the flat bf16 gradient buffer
  PeerPtrs weight;   // per-rank base of the flat bf16 weight buffer
  PeerPtrs flags;    // per-rank flag pads
  void* grad_mc;     // multicast alias of the gradient buffer (NVLS) or null
  void* weight_mc;   // multicast alias of the weight buffer (NVLS) or null
  float* master;     // this rank's fp32 master SHARD for the bucket   [slice_len]
  float* mom;        // this rank's fp32 momentum SHARD for the bucket [slice_len]
  const float* hyper;  // {lr, momentum, weight_decay, grad_scale}
  uint32_t* epoch;   // per-CTA epoch counters [max_blocks]
  int* status;
  int64_t bucket_off;  // element offset of the bucket in the flat buffers
  int64_t slice_len;   // elements per rank slice (multiple of 8)
  int rank, world, max_blocks;
};

// MODE 0: peer-to-peer loads/stores.  MODE 1: NVLS multimem.
// U independent 16-byte vectors per thread per trip, KW = compile-time bound on the world size:
// NVLink round trips are ~2-3 us, so the bytes in flight (CTAs x threads x U x 16 B x world for
// peer loads) set the bandwidth.  With ONE peer (the 2-replica stages of a hybrid pipeline) U = 2
// left 1 MB in flight and the kernel at 0.4x NCCL (profiles/allreduce_sweep_n2.json, round 1);
// the generic kernel holds U x KW at 16 vectors per thread — (8, 2), (4, 4), (2, 8) —, the fused
// one (which also carries the fp32 update of U vectors in registers) at (4, 2), (2, 4), (2, 8) and
// makes up for it with more CTAs when there is a single peer.
template <int MODE, int U, int KW>
__global__ void __launch_bounds__(kArThreads, 1) ar_sgd_twoshot_kernel(ArSgdArgs a) {
  const uint32_t epoch = a.epoch[blockIdx.x] + 1;
  peer_barrier(a.flags, a.rank, a.world, 0, a.max_blocks, epoch, a.status);

  const float lr = a.hyper[0], mu = a.hyper[1], wd = a.hyper[2], gs = a.hyper[3];
  const int64_t nvec = a.slice_len >> 3;
  const int64_t base = a.bucket_off + (int64_t)a.rank * a.slice_len;  // my slice, in elements
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * U) {
    float g[U][8];
    if (MODE == 1) {
      int4 red[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v < nvec) red[u] = mc_ld_reduce_bf16x8((const __nv_bfloat16*)a.grad_mc + base + (v << 3));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) unpack8(*reinterpret_cast<bf16x8*>(&red[u]), g[u]);
    } else {
      int4 raw[U][KW];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v < nvec) {
#pragma unroll
          for (int r = 0; r < KW; ++r)
            if (r < a.world)
              raw[u][r] = ld_peer16((const __nv_bfloat16*)a.grad.p[(a.rank + r) % a.world] + base + (v << 3));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[u][k] = 0.f;
#pragma unroll
        for (int r = 0; r < KW; ++r)
          if (r < a.world) {
            float f[8];
            unpack8(*reinterpret_cast<bf16x8*>(&raw[u][r]), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[u][k] += f[k];
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v >= nvec) break;
      const int64_t e = base + (v << 3);
      // SGD on the local shard
      float* mp = a.master + (v << 3);
      float* vp = a.mom + (v << 3);
      float4 p0 = *reinterpret_cast<float4*>(mp), p1 = *reinterpret_cast<float4*>(mp + 4);
      float4 m0 = *reinterpret_cast<float4*>(vp), m1 = *reinterpret_cast<float4*>(vp + 4);
      float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gk = fmaf(wd, p[k], g[u][k] * gs);
        m[k] = fmaf(mu, m[k], gk);
        p[k] = fmaf(-lr, m[k], p[k]);
      }
      *reinterpret_cast<float4*>(mp) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(mp + 4) = make_float4(p[4], p[5], p[6], p[7]);
      *reinterpret_cast<float4*>(vp) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(vp + 4) = make_float4(m[4], m[5], m[6], m[7]);
      bf16x8 w = pack8(p);
      const int4 wv = *reinterpret_cast<int4*>(&w);
      if (MODE == 1) {
        mc_st16((__nv_bfloat16*)a.weight_mc + e, wv);
      } else {
#pragma unroll
        for (int r = 0; r < KW; ++r)
          if (r < a.world) st_peer16((__nv_bfloat16*)a.weight.p[(a.rank + r) % a.world] + e, wv);
      }
    }
  }
  peer_barrier(a.flags, a.rank, a.world, 1, a.max_blocks, epoch, a.status);
  if (threadIdx.x == 0) a.epoch[blockIdx.x] = epoch;
}

// Generic in-place sum all-reduce of a symmetric buffer (no optimizer): reduce-scatter by
// loads + all-gather by stores.  T = bf16 or float.  scale is applied to the sum.
struct ArArgs {
  PeerPtrs buf;
  PeerPtrs flags;
  void* buf_mc;
  uint32_t* epoch;
  int* status;
  int64_t off;        // element offset
  int64_t slice_len;  // elements per rank (multiple of 16 bytes worth)
  float scale;
  int rank, world, max_blocks;
};

template <typename T, int MODE, int U, int KW>
__global__ void __launch_bounds__(kArThreads, 1) ar_twoshot_kernel(ArArgs a) {
  constexpr int EPV = 16 / sizeof(T);  // elements per 16-byte vector
  const uint32_t epoch = a.epoch[blockIdx.x] + 1;
  peer_barrier(a.flags, a.rank, a.world, 0, a.max_blocks, epoch, a.status);
  const int64_t nvec = a.slice_len / EPV;
  const int64_t base = a.off + (int64_t)a.rank * a.slice_len;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * U) {
    int4 outv[U];
    if (MODE == 1) {
      int4 red[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v < nvec) {
          if (sizeof(T) == 2) red[u] = mc_ld_reduce_bf16x8((const T*)a.buf_mc + base + v * EPV);
          else {
            float4 f = mc_ld_reduce_f32x4((const T*)a.buf_mc + base + v * EPV);
            red[u] = *reinterpret_cast<int4*>(&f);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (sizeof(T) == 2) {
          float f[8];
          unpack8(*reinterpret_cast<bf16x8*>(&red[u]), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] *= a.scale;
          bf16x8 o = pack8(f);
          outv[u] = *reinterpret_cast<int4*>(&o);
        } else {
          float4 f = *reinterpret_cast<float4*>(&red[u]);
          f.x *= a.scale; f.y *= a.scale; f.z *= a.scale; f.w *= a.scale;
          outv[u] = *reinterpret_cast<int4*>(&f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v < nvec) mc_st16((T*)a.buf_mc + base + v * EPV, outv[u]);
      }
    } else {
      int4 raw[U][KW];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v < nvec) {
#pragma unroll
          for (int r = 0; r < KW; ++r)
            if (r < a.world) raw[u][r] = ld_peer16((const T*)a.buf.p[(a.rank + r) % a.world] + base + v * EPV);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (sizeof(T) == 2) {
          float acc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
          for (int r = 0; r < KW; ++r)
            if (r < a.world) {
              float f[8];
              unpack8(*reinterpret_cast<bf16x8*>(&raw[u][r]), f);
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] *= a.scale;
          bf16x8 o = pack8(acc);
          outv[u] = *reinterpret_cast<int4*>(&o);
        } else {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < KW; ++r)
            if (r < a.world) {
              float4 f = *reinterpret_cast<float4*>(&raw[u][r]);
              acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
            }
          acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
          outv[u] = *reinterpret_cast<int4*>(&acc);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v < nvec) {
#pragma unroll
          for (int r = 0; r < KW; ++r)
            if (r < a.world) st_peer16((T*)a.buf.p[(a.rank + r) % a.world] + base + v * EPV, outv[u]);
        }
      }
    }
  }
  peer_barrier(a.flags, a.rank, a.world, 1, a.max_blocks, epoch, a.status);
  if (threadIdx.x == 0) a.epoch[blockIdx.x] = epoch;
}

// One-shot (latency-optimal, small messages): every rank reads all peers and writes only
// its own output buffer `out` (may be ordinary local memory).
template <typename T>
__global__ void __launch_bounds__(kArThreads, 1)
ar_oneshot_kernel(ArArgs a, T* __restrict__ out, int64_t n) {
  constexpr int EPV = 16 / sizeof(T);
  const uint32_t epoch = a.epoch[blockIdx.x] + 1;
  peer_barrier(a.flags, a.rank, a.world, 0, a.max_blocks, epoch, a.status);
  const int64_t nvec = n / EPV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int64_t e = a.off + v * EPV;
    int4 raw[kMaxRanks];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < a.world) raw[r] = ld_peer16((const T*)a.buf.p[r] + e);  // fixed order: bitwise equal on all ranks
    int4 outv;
    if (sizeof(T) == 2) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < a.world) {
          float f[8];
          unpack8(*reinterpret_cast<bf16x8*>(&raw[r]), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= a.scale;
      bf16x8 o = pack8(acc);
      outv = *reinterpret_cast<int4*>(&o);
    } else {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < a.world) {
          float4 f = *reinterpret_cast<float4*>(&raw[r]);
          acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
        }
      acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
      outv = *reinterpret_cast<int4*>(&acc);
    }
    *reinterpret_cast<int4*>(out + v * EPV) = outv;
  }
  // peers must not overwrite their buffers until everyone has read them
  peer_barrier(a.flags, a.rank, a.world, 1, a.max_blocks, epoch, a.status);
  if (threadIdx.x == 0) a.epoch[blockIdx.x] = epoch;
}

// Device-side barrier only (used to fence symmetric-buffer reuse, and as the "ack" token
// of pipelined evaluation — reference: pipedream-fork/runtime/runtime.py:630-654).
__global__ void peer_barrier_kernel(ArArgs a) {
  const uint32_t epoch = a.epoch[blockIdx.x] + 1;
  peer_barrier(a.flags, a.rank, a.world, 0, a.max_blocks, epoch, a.status);
  if (threadIdx.x == 0) a.epoch[blockIdx.x] = epoch;
}

// Copy `n` bytes (multiple of 16) from local `src` into a peer's buffer and then raise a
// flag there: the producer side of a stage-boundary transfer done by stores over NVLink.
__global__ void __launch_bounds__(kArThreads)
p2p_put_kernel(const int4* __restrict__ src, int4* __restrict__ dst_peer, int64_t nvec,
               uint32_t* __restrict__ peer_flag, uint32_t* __restrict__ local_seq,
               unsigned int* __restrict__ ticket) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride)
    st_peer16(dst_peer + v, ld_stream(src + v));
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t seq = *local_seq + 1;
    *local_seq = seq;
    *ticket = 0u;
    st_release_sys(peer_flag, seq);
  }
}

// Consumer side: spin until the local flag reaches the next sequence number.
__global__ void p2p_wait_kernel(const uint32_t* __restrict__ flag, uint32_t* __restrict__ local_seq,
                                int* __restrict__ status) {
  const uint32_t want = *local_seq + 1;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - want) < 0) {
    if (clock64() - t0 > 20000000000LL) {
      atomicExch(status, 3);
      break;
    }
  }
  *local_seq = want;
}

}  // namespace ddl

using namespace ddl;

static void fill_peers(PeerPtrs& dst, void* const* src, int world) {
  for (int i = 0; i < kMaxRanks; ++i) dst.p[i] = i < world ? src[i] : nullptr;
}

extern "C" int ddl_ar_sgd_twoshot(void* const* grad_ptrs, void* const* weight_ptrs,
                                  void* const* flag_ptrs, void* grad_mc, void* weight_mc,
                                  float* master, float* mom, const float* hyper, uint32_t* epoch,
                                  int* status, int64_t bucket_off, int64_t slice_len, int rank,
                                  int world, int blocks, int max_blocks, int use_multimem,
                                  cudaStream_t st) {
  if (world > kMaxRanks || slice_len % 8 || blocks > max_blocks) return -1;
  ArSgdArgs a;
  fill_peers(a.grad, grad_ptrs, world);
  fill_peers(a.weight, weight_ptrs, world);
  fill_peers(a.flags, flag_ptrs, world);
  a.grad_mc = grad_mc; a.weight_mc = weight_mc;
  a.master = master; a.mom = mom; a.hyper = hyper; a.epoch = epoch; a.status = status;
  a.bucket_off = bucket_off; a.slice_len = slice_len;
  a.rank = rank; a.world = world; a.max_blocks = max_blocks;
  if (use_multimem && grad_mc && weight_mc)
    ar_sgd_twoshot_kernel<1, 4, 8><<<blocks, kArThreads, 0, st>>>(a);
  else if (world <= 2)
    ar_sgd_twoshot_kernel<0, 4, 2><<<blocks, kArThreads, 0, st>>>(a);
  else if (world <= 4)
    ar_sgd_twoshot_kernel<0, 2, 4><<<blocks, kArThreads, 0, st>>>(a);
  else
    ar_sgd_twoshot_kernel<0, 2, 8><<<blocks, kArThreads, 0, st>>>(a);
  return (int)cudaGetLastError();
}

static int fill_ar(ArArgs& a, void* const* buf_ptrs, void* const* flag_ptrs, void* buf_mc,
                   uint32_t* epoch, int* status, int64_t off, int64_t slice_len, float scale,
                   int rank, int world, int max_blocks) {
  if (world > kMaxRanks) return -1;
  fill_peers(a.buf, buf_ptrs, world);
  fill_peers(a.flags, flag_ptrs, world);
  a.buf_mc = buf_mc; a.epoch = epoch; a.status = status; a.off = off; a.slice_len = slice_len;
  a.scale = scale; a.rank = rank; a.world = world; a.max_blocks = max_blocks;
  return 0;
}

extern "C" int ddl_ar_twoshot(void* const* buf_ptrs, void* const* flag_ptrs, void* buf_mc,
                              int is_bf16, uint32_t* epoch, int* status, int64_t off,
                              int64_t slice_len, float scale, int rank, int world, int blocks,
                              int max_blocks, int use_multimem, cudaStream_t st) {
  ArArgs a;
  if (fill_ar(a, buf_ptrs, flag_ptrs, buf_mc, epoch, status, off, slice_len, scale, rank, world,
              max_blocks))
    return -1;
  if (blocks > max_blocks || (slice_len * (is_bf16 ? 2 : 4)) % 16) return -1;
  const bool mm = use_multimem && buf_mc;
#define DDL_AR2(T)                                                                             \
  do {                                                                                         \
    if (mm) ar_twoshot_kernel<T, 1, 4, 8><<<blocks, kArThreads, 0, st>>>(a);                   \
    else if (world <= 2) ar_twoshot_kernel<T, 0, 8, 2><<<blocks, kArThreads, 0, st>>>(a);      \
    else if (world <= 4) ar_twoshot_kernel<T, 0, 4, 4><<<blocks, kArThreads, 0, st>>>(a);      \
    else ar_twoshot_kernel<T, 0, 2, 8><<<blocks, kArThreads, 0, st>>>(a);                      \
  } while (0)
  if (is_bf16) DDL_AR2(__nv_bfloat16);
  else DDL_AR2(float);
#undef DDL_AR2
  return (int)cudaGetLastError();
}

extern "C" int ddl_ar_oneshot(void* const* buf_ptrs, void* const* flag_ptrs, int is_bf16,
                              void* out, uint32_t* epoch, int* status, int64_t off, int64_t n,
                              float scale, int rank, int world, int blocks, int max_blocks,
                              cudaStream_t st) {
  ArArgs a;
  if (fill_ar(a, buf_ptrs, flag_ptrs, nullptr, epoch, status, off, 0, scale, rank, world,
              max_blocks))
    return -1;
  if (blocks > max_blocks || (n * (is_bf16 ? 2 : 4)) % 16) return -1;
  if (is_bf16)
    ar_oneshot_kernel<__nv_bfloat16><<<blocks, kArThreads, 0, st>>>(a, (__nv_bfloat16*)out, n);
  else
    ar_oneshot_kernel<float><<<blocks, kArThreads, 0, st>>>(a, (float*)out, n);
  return (int)cudaGetLastError();
}

extern "C" int ddl_peer_barrier(void* const* flag_ptrs, uint32_t* epoch, int* status, int rank,
                                int world, int max_blocks, cudaStream_t st) {
  ArArgs a;
  void* dummy[kMaxRanks] = {nullptr};
  if (fill_ar(a, dummy, flag_ptrs, nullptr, epoch, status, 0, 0, 1.f, rank, world, max_blocks))
    return -1;
  peer_barrier_kernel<<<1, 32, 0, st>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int ddl_p2p_put(const void* src, void* dst_peer, int64_t nbytes, uint32_t* peer_flag,
                           uint32_t* local_seq, unsigned int* ticket, int blocks,
                           cudaStream_t st) {
  if (nbytes % 16) return -1;
  p2p_put_kernel<<<blocks, kArThreads, 0, st>>>((const int4*)src, (int4*)dst_peer, nbytes / 16,
                                               peer_flag, local_seq, ticket);
  return (int)cudaGetLastError();
}

extern "C" int ddl_p2p_wait(const uint32_t* flag, uint32_t* local_seq, int* status,
                            cudaStream_t st) {
  p2p_wait_kernel<<<1, 1, 0, st>>>(flag, local_seq, status);
  return (int)cudaGetLastError();
}
