// Small memory-bound kernels: input normalisation, variable-length sequence reversal /
// masks (the GNMT helper ops), dropout, global average pooling.
#include "common.cuh"
#include "api.h"

namespace ddl {

// ---------------------------------------------------------------------------------------
// uint8 HWC image batch -> normalised bf16 NHWC (channels-last) in one pass.
// The reference normalises on the CPU inside DataLoader workers (transforms.ToTensor +
// Normalize, benchmark/imagenet/imagenet_pytorch.py:68-80) and ships fp32 NCHW over PCIe
// (4 bytes/value); here the host->device copy carries the raw uint8 pixels (1 byte/value)
// and the conversion is fused into the first device touch.
__global__ void __launch_bounds__(256)
u8_to_bf16_norm_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                       int64_t n, int C, float4 scale, float4 shift) {
  pdl_enter();
  // out = in * scale[c] + shift[c]   with scale = 1/(255*std), shift = -mean/std
  const float sc[4] = {scale.x, scale.y, scale.z, scale.w};
  const float sh[4] = {shift.x, shift.y, shift.z, shift.w};
  const int64_t nvec = n >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int64_t e0 = v << 4;
    int4 raw = ld_stream(in + e0);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&raw);
    // channel of the first element: (16 v) mod C in 32-bit arithmetic (64-bit % is ~100 instrs)
    int c = (int)((((unsigned)v % (unsigned)C) * (16u % (unsigned)C)) % (unsigned)C);
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      f[k] = fmaf((float)b[k], sc[c], sh[c]);
      c = (c + 1 == C) ? 0 : c + 1;
    }
    *reinterpret_cast<bf16x8*>(out + e0) = pack8(f);
    *reinterpret_cast<bf16x8*>(out + e0 + 8) = pack8(f + 8);
  }
  for (int64_t e = (nvec << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += stride) {
    const int c = (int)(e % C);
    out[e] = __float2bfloat16_rn(fmaf((float)in[e], sc[c], sh[c]));
  }
}

// ---------------------------------------------------------------------------------------
// Per-sequence time reversal of a padded (T, B, F) tensor honouring lengths:
//   out[t, b, :] = in[len_b - 1 - t, b, :] for t < len_b, else 0
// Used forward AND backward by the emulated bidirectional LSTM of GNMT.  Reference:
// pipedream-fork/runtime/translation/seq2seq/csrc/pack_utils_kernel.cu:7-13 launches B*T
// blocks of 512 threads doing scalar strided copies, after building an int64 offsets
// vector on the CPU and copying it H2D on every call (:37-54), compiled for sm_70 only.
// Here: lengths stay on the device, one thread moves 16 bytes, no host round trip.
template <int VEC_BYTES>
__global__ void __launch_bounds__(256)
revert_varlen_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                     const int64_t* __restrict__ lengths, int T, int B, int64_t row_bytes) {
  pdl_enter();
  const int64_t vec_per_row = row_bytes / VEC_BYTES;
  const int64_t total = (int64_t)T * B * vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t v = i % vec_per_row;
    const int64_t tb = i / vec_per_row;
    const int b = (int)(tb % B);
    const int t = (int)(tb / B);
    const int len = (int)lengths[b];
    uint8_t* dst = out + (tb * vec_per_row + v) * VEC_BYTES;
    if (t < len) {
      const int64_t src_tb = (int64_t)(len - 1 - t) * B + b;
      const uint8_t* src = in + (src_tb * vec_per_row + v) * VEC_BYTES;
      if (VEC_BYTES == 16) *reinterpret_cast<int4*>(dst) = ld_stream(src);
      else if (VEC_BYTES == 4) *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
      else if (VEC_BYTES == 2) *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(src);
      else *dst = *src;
    } else {
      if (VEC_BYTES == 16) *reinterpret_cast<int4*>(dst) = make_int4(0, 0, 0, 0);
      else if (VEC_BYTES == 4) *reinterpret_cast<uint32_t*>(dst) = 0u;
      else if (VEC_BYTES == 2) *reinterpret_cast<uint16_t*>(dst) = 0;
      else *dst = 0;
    }
  }
}

// (T, B) uint8 validity mask from lengths — reference builds it with a CPU loop
// (pack_utils.cpp:13-31, `set_mask_cpp`).
__global__ void set_mask_kernel(const int64_t* __restrict__ lengths, uint8_t* __restrict__ mask,
                                int T, int B) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * B) return;
  const int b = i % B, t = i / B;
  mask[i] = t < (int)lengths[b] ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// Global average pool over H*W for NHWC bf16: [N, HW, C] -> [N, C] (bf16), and its backward.
// grid (N, ceil(C/8/32)), block (32 channel vectors, 8 row lanes): the 8 lanes split the H*W
// rows, a shared-memory tree adds them (the first version used one thread per channel vector
// looping over all rows: 32 CTAs, 17 us for 6 MB).
__global__ void __launch_bounds__(256)
gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int HW, int C) {
  pdl_enter();
  __shared__ float red[8][8][33];
  const int n = blockIdx.x;
  const int cv_total = C >> 3;
  const int cv = blockIdx.y * 32 + threadIdx.x;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (cv < cv_total) {
    const __nv_bfloat16* p = x + (int64_t)n * HW * C + cv * 8;
    for (int i = threadIdx.y; i < HW; i += 8) {
      int4 raw = ld_stream(p + (int64_t)i * C);
      float f[8];
      unpack8(*reinterpret_cast<bf16x8*>(&raw), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.y][k][threadIdx.x] = acc[k];
  __syncthreads();
  if (threadIdx.y == 0 && cv < cv_total) {
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) s += red[l][k][threadIdx.x];
      acc[k] = s * inv;
    }
    *reinterpret_cast<bf16x8*>(y + (int64_t)n * C + cv * 8) = pack8(acc);
  }
}

__global__ void __launch_bounds__(256)
gap_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int64_t nvec,
               int HW, int C) {
  pdl_enter();
  const int cv_total = C >> 3;
  const float inv = 1.f / (float)HW;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % cv_total);
    const int64_t n = v / ((int64_t)cv_total * HW);
    bf16x8 g = *reinterpret_cast<const bf16x8*>(dy + n * C + cv * 8);
    float f[8];
    unpack8(g, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= inv;
    *reinterpret_cast<bf16x8*>(dx + v * 8) = pack8(f);
  }
}

}  // namespace ddl

using namespace ddl;

static inline int cap_blocks(int64_t work, int per_block, int sms, int mult) {
  int64_t b = (work + per_block - 1) / per_block;
  if (b > (int64_t)sms * mult) b = (int64_t)sms * mult;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int ddl_u8_to_bf16_norm(const void* in, void* out, int64_t n, int C, const float* mean,
                                   const float* stdv, int sms, cudaStream_t st) {
  if (C < 1 || C > 4) return -1;
  float sc[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0};
  for (int c = 0; c < C; ++c) {
    sc[c] = 1.f / (255.f * stdv[c]);
    sh[c] = -mean[c] / stdv[c];
  }
  launch_pdl(u8_to_bf16_norm_kernel, dim3(cap_blocks(n / 16 + 1, 256, sms, 16)), dim3(256), 0, st, 
      (const uint8_t*)in, (__nv_bfloat16*)out, n, C, make_float4(sc[0], sc[1], sc[2], sc[3]),
      make_float4(sh[0], sh[1], sh[2], sh[3]));
  return (int)cudaGetLastError();
}

extern "C" int ddl_revert_varlen(const void* in, void* out, const int64_t* lengths, int T, int B,
                                 int64_t row_bytes, int sms, cudaStream_t st) {
  if (T <= 0 || B <= 0 || row_bytes <= 0) return 0;
  const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
#define L(V)                                                                                  \
  launch_pdl(revert_varlen_kernel<V>, dim3(cap_blocks((int64_t)T * B * (row_bytes / V), 256, sms, 16)), dim3(256), 0, \
                            st, (const uint8_t*)in, (uint8_t*)out, lengths, T, B, row_bytes)
  if (row_bytes % 16 == 0 && aligned) L(16);
  else if (row_bytes % 4 == 0) L(4);
  else if (row_bytes % 2 == 0) L(2);
  else L(1);
#undef L
  return (int)cudaGetLastError();
}

extern "C" int ddl_set_mask(const int64_t* lengths, void* mask, int T, int B, cudaStream_t st) {
  if (T * B <= 0) return 0;
  launch_pdl(set_mask_kernel, dim3((T * B + 255) / 256), dim3(256), 0, st, lengths, (uint8_t*)mask, T, B);
  return (int)cudaGetLastError();
}

extern "C" int ddl_gap_fwd(const void* x, void* y, int N, int HW, int C, cudaStream_t st) {
  if (C % 8) return -1;
  launch_pdl(gap_fwd_kernel, dim3(dim3(N, (C / 8 + 31) / 32)), dim3(dim3(32, 8)), 0, st, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, HW, C);
  return (int)cudaGetLastError();
}

extern "C" int ddl_gap_bwd(const void* dy, void* dx, int N, int HW, int C, int sms,
                           cudaStream_t st) {
  if (C % 8) return -1;
  const int64_t nvec = (int64_t)N * HW * (C / 8);
  launch_pdl(gap_bwd_kernel, dim3(cap_blocks(nvec, 256, sms, 16)), dim3(256), 0, st, 
      (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, nvec, HW, C);
  return (int)cudaGetLastError();
}
