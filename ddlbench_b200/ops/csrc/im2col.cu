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
// One CTA per group of output rows of one image: the input rows it needs are staged once in shared
// memory (zero-padded left/right/top/bottom), then every thread assembles 16-byte vectors of A from
// shared memory:  k -> (kh = k / (KW*C), rem = k % (KW*C)),  smem index = kh*rowlen + ow*s*C + rem.
#include "api.h"
#include "common.cuh"

namespace ddl {

// kRows output rows per CTA: the (kRows-1)*stride + KH input rows are staged once (3.5x -> ~2.2x
// read amplification at kRows = 4) and the k -> shared-memory offset table is built once per CTA,
// so the inner loop is 8 two-byte shared loads + one 16-byte store with no integer division.
constexpr int kIm2colRows = 4;
constexpr int kIm2colMaxK = 512;

__global__ void __launch_bounds__(256)
im2col_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ a, int H, int W,
                   int C, int OH, int OW, int KH, int KW, int stride, int pad, int Kpad) {
  extern __shared__ __align__(16) __nv_bfloat16 rows[];     // [IR][(W + 2*pad) * C], then the zero slot
  __shared__ int koff[kIm2colMaxK];
  pdl_enter();
  const int groups = (OH + kIm2colRows - 1) / kIm2colRows;
  const int n = blockIdx.x / groups, oh0 = (blockIdx.x % groups) * kIm2colRows;
  const int nrows = min(kIm2colRows, OH - oh0);
  const int rowlen = (W + 2 * pad) * C;
  const int in_rows = (nrows - 1) * stride + KH;
  const int ih0 = oh0 * stride - pad;
  const int K = KH * KW * C, R = KW * C;
  const int zero_slot = in_rows * rowlen;                     // padding columns of A read this
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) koff[k] = k < K ? (k / R) * rowlen + (k % R) : -1;
  if (threadIdx.x == 0) rows[zero_slot] = __float2bfloat16_rn(0.f);
  const __nv_bfloat16* xin = x + (size_t)n * H * W * C;
  for (int r = 0; r < in_rows; ++r) {
    const int ih = ih0 + r;
    const bool row_ok = ih >= 0 && ih < H;
    const __nv_bfloat16* src = xin + (size_t)ih * W * C - pad * C;
    for (int j = threadIdx.x; j < rowlen; j += blockDim.x) {
      const int iwc = j - pad * C;
      rows[r * rowlen + j] = (row_ok && iwc >= 0 && iwc < W * C) ? src[j] : __float2bfloat16_rn(0.f);
    }
  }
  __syncthreads();
  const int vec_per_row = Kpad >> 3;
  // thread -> (vector slot inside a patch row, first output pixel): both fixed for the whole loop
  const int lanes = blockDim.x / vec_per_row;                 // output pixels in flight per pass
  const int slot = threadIdx.x % vec_per_row, px0 = threadIdx.x / vec_per_row;
  if (px0 >= lanes) return;
  int off[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) off[t] = koff[slot * 8 + t];
  const int npx = nrows * OW;
  __nv_bfloat16* out = a + ((size_t)n * OH + oh0) * OW * Kpad + slot * 8;
  for (int px = px0; px < npx; px += lanes) {
    const int r = px / OW, ow = px - r * OW;
    const int base = r * stride * rowlen + ow * stride * C;
    __align__(16) __nv_bfloat16 e[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) e[t] = rows[off[t] >= 0 ? base + off[t] : zero_slot];
    *reinterpret_cast<int4*>(out + (size_t)px * Kpad) = *reinterpret_cast<const int4*>(e);
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
  const size_t smem = ((size_t)((kIm2colRows - 1) * stride + KH) * (W + 2 * pad) * C + 8) * sizeof(__nv_bfloat16);
  if (smem > 48 * 1024 || Kpad > kIm2colMaxK || Kpad / 8 > 256) return -1;
  const int groups = (OH + kIm2colRows - 1) / kIm2colRows;
  cudaError_t e = launch_pdl(im2col_rows_kernel, dim3(N * groups), dim3(256), smem, st,
                             (const __nv_bfloat16*)x, (__nv_bfloat16*)a, H, W, C, OH, OW, KH, KW,
                             stride, pad, Kpad);
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}
