// Patch matrix for the first convolution of the ImageNet models (7x7, stride 2, 3 input channels):
//
//     A[(n, oh, ow)][(kh*KW + kw)*C + c] = X[n, oh*s - p + kh, ow*s - p + kw, c]      (0 outside)
//
// padded along K to a multiple of 8 so the tcgen05 GEMM (gemm_tcgen05.cu) can consume it through
// TMA.  With 3 channels there is nothing for a tensor-core implicit GEMM to chew on per tap
// (6 bytes per pixel), and cuDNN answers this layer with an sm_80 kernel (ncu: 200 us forward at
// batch 32, 4 % of the ResNet-50 step); as patch matrix + GEMM the layer is ~120 MB of streaming
// writes + reads, and its BatchNorm statistics come out of the GEMM epilogue for free.
//
// One CTA per output row (n, oh): the KH input rows it needs are staged once in shared memory
// (zero-padded left/right/top/bottom), then every thread assembles 16-byte vectors of A from
// shared memory:  k -> (kh = k / (KW*C), rem = k % (KW*C)),  smem index = kh*rowlen + ow*s*C + rem.
//
// STATUS: opt-in (DDLB_STEM_IM2COL=1), numerics test opt-in — written after round 1's GPU budget
// was spent, not yet run on hardware.
#include "api.h"
#include "common.cuh"

namespace ddl {

__global__ void __launch_bounds__(256)
im2col_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ a, int H, int W,
                   int C, int OH, int OW, int KH, int KW, int stride, int pad, int Kpad) {
  extern __shared__ __align__(16) __nv_bfloat16 rows[];     // [KH][(W + 2*pad) * C]
  pdl_enter();
  const int n = blockIdx.x / OH, oh = blockIdx.x % OH;
  const int rowlen = (W + 2 * pad) * C;
  const int ih0 = oh * stride - pad;
  for (int i = threadIdx.x; i < KH * rowlen; i += blockDim.x) {
    const int kh = i / rowlen, j = i % rowlen;
    const int ih = ih0 + kh;
    const int iwc = j - pad * C;                              // (iw * C + c) of the unpadded row
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (ih >= 0 && ih < H && iwc >= 0 && iwc < W * C)
      v = x[((size_t)n * H + ih) * W * C + iwc];
    rows[i] = v;
  }
  __syncthreads();
  const int K = KH * KW * C, R = KW * C;
  const int vec_per_row = Kpad >> 3;
  __nv_bfloat16* out = a + ((size_t)n * OH + oh) * OW * Kpad;
  for (int v = threadIdx.x; v < OW * vec_per_row; v += blockDim.x) {
    const int ow = v / vec_per_row, k0 = (v % vec_per_row) * 8;
    const int base = ow * stride * C;
    __align__(16) __nv_bfloat16 e[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = k0 + t;
      e[t] = k < K ? rows[(k / R) * rowlen + base + (k % R)] : __float2bfloat16_rn(0.f);
    }
    *reinterpret_cast<int4*>(out + (size_t)ow * Kpad + k0) = *reinterpret_cast<const int4*>(e);
  }
}

}  // namespace ddl

// x: [N, H, W, C] bf16 NHWC; a: [N*OH*OW, Kpad] bf16 with Kpad = roundup(KH*KW*C, 8).
extern "C" int ddl_im2col_rows(const void* x, void* a, int N, int H, int W, int C, int KH, int KW,
                               int stride, int pad, int Kpad, cudaStream_t st) {
  using namespace ddl;
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  if (N <= 0 || OH <= 0 || OW <= 0) return 0;
  if (Kpad % 8 || Kpad < KH * KW * C) return -1;
  const size_t smem = (size_t)KH * (W + 2 * pad) * C * sizeof(__nv_bfloat16);
  if (smem > 48 * 1024) return -1;
  cudaError_t e = launch_pdl(im2col_rows_kernel, dim3(N * OH), dim3(256), smem, st,
                             (const __nv_bfloat16*)x, (__nv_bfloat16*)a, H, W, C, OH, OW, KH, KW,
                             stride, pad, Kpad);
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}
