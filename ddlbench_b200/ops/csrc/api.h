// C launch API of the ddlbench_b200 sm_100a kernels.  Pure CUDA (no torch headers) so
// each .cu compiles in seconds; bindings.cpp adapts torch tensors to these calls.
// Every function returns 0 on success, a cudaError_t value, or -1 for bad arguments.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#ifdef __cplusplus
extern "C" {
#endif

// ---- sgd.cu
int ddl_sgd_flat(void* master, void* mom, const void* grad, int grad_is_bf16, void* w_out,
                 void* w_out2, int w_is_bf16, const float* hyper, int64_t n, int sms,
                 cudaStream_t stream);
int ddl_sumsq(const void* x, int is_bf16, float* out, int64_t n, int sms, cudaStream_t stream);

// ---- xent.cu
int ddl_xent(const void* logits, int is_bf16, const int64_t* target, void* dlogits,
             float* row_stats, float* out, unsigned int* ticket, int rows, int C, float smoothing,
             float grad_scale, cudaStream_t stream);

// ---- bn.cu
int ddl_bn_replicas();
int ddl_bn_stats(const void* x, float* sums, unsigned int* counter, float* stat,
                 float* running_mean, float* running_var, int64_t M, int C, float eps,
                 float momentum, int sms, cudaStream_t st);
int ddl_bn_finalize(float* sums, float* stat, float* running_mean, float* running_var, int64_t M,
                    int C, float eps, float momentum, cudaStream_t st);
int ddl_bn_apply(const void* x, const void* res, void* y, const float* stat, const void* gamma,
                 const void* beta, int param_is_bf16, void* mask_out, int64_t M, int C, int act,
                 int sms, cudaStream_t st);
int ddl_bn_apply_sums(const void* x, const void* res, void* y, float* sums, float* stat_out,
                      float* running_mean, float* running_var, unsigned int* ticket,
                      const void* gamma, const void* beta, int param_is_bf16, void* mask_out,
                      int64_t M, int C, float eps, float momentum, int act, int sms, cudaStream_t st);
int ddl_bn_bwd_reduce(const void* dy, const void* dy2, const void* y, const void* x, void* dz_out,
                      const float* stat, float* sums, unsigned int* counter, const void* gamma,
                      int param_is_bf16, float* coef, void* dgamma, void* dbeta, int grad_is_bf16,
                      int accumulate, int64_t M, int C, int act, int sms, cudaStream_t st);
int ddl_bn_bwd_apply(const void* dz_in, const void* dy, const void* dy2, const void* y, const void* x, void* dx,
                     const float* stat, const float* coef, int64_t M, int C, int act, int sms,
                     cudaStream_t st);
int ddl_conv_bf16_nhwc(const void* x, const void* w, void* y, int NB, int H, int W, int Cin, int Cout,
                       int KH, int KW, int pad_h, int pad_w, int stride, float* col_sums,
                       const void* bias, int relu, int sms, cudaStream_t st);
int ddl_dwconv3x3(int mode, const void* a, const void* b, void* out, int N, int H, int W, int C,
                  int stride, int pad, int sms, cudaStream_t st);
void ddl_conv_box(int NB, int H, int W, int* out3);
int ddl_conv_wgrad_bf16_nhwc(const void* dy, const void* x, float* D, int NB, int H, int W, int Cin,
                             int Cout, int KH, int KW, int pad_h, int pad_w, int stride, int sms,
                             cudaStream_t st);
int ddl_im2col_rows(const void* x, void* a, int N, int H, int W, int C, int KH, int KW, int stride,
                    int pad, int Kpad, cudaStream_t st);
int ddl_bn_bwd_fused(const void* dy, const void* dy2, const void* ymask, const void* x, void* dz_out,
                     void* dx, const float* stat, const void* gamma, int param_is_bf16, void* dgamma,
                     void* dbeta, int grad_is_bf16, int accumulate, int64_t M, int C, int act_flags,
                     int sms, cudaStream_t st);
int ddl_bn_fwd_fused(const void* x, const void* res, void* y, void* mask_out, float* stat_out,
                     float* running_mean, float* running_var, const void* gamma, const void* beta,
                     int param_is_bf16, int64_t M, int C, float eps, float momentum, int act, int sms,
                     cudaStream_t st);
int ddl_bn_cluster_plan(int64_t M, int C, int tiles, int sms, int* rows_per_cta, int64_t* smem);
int ddl_bias_act(const void* x, const void* bias, int bias_is_bf16, void* y, int64_t M, int C,
                 int act, int sms, cudaStream_t st);
int ddl_act_bwd(const void* dy, const void* dy2, const void* y, void* dz, int64_t n, int act,
                int sms, cudaStream_t st);
int ddl_colsum(const void* x, float* sums, unsigned int* counter, void* out, int out_is_bf16,
               int accumulate, int64_t M, int C, int sms, cudaStream_t st);

// ---- allreduce.cu
int ddl_ar_sgd_twoshot(void* const* grad_ptrs, void* const* weight_ptrs, void* const* flag_ptrs,
                       void* grad_mc, void* weight_mc, float* master, float* mom,
                       const float* hyper, uint32_t* epoch, int* status, int64_t bucket_off,
                       int64_t slice_len, int rank, int world, int blocks, int max_blocks,
                       int use_multimem, cudaStream_t st);
int ddl_ar_twoshot(void* const* buf_ptrs, void* const* flag_ptrs, void* buf_mc, int is_bf16,
                   uint32_t* epoch, int* status, int64_t off, int64_t slice_len, float scale,
                   int rank, int world, int blocks, int max_blocks, int use_multimem,
                   cudaStream_t st);
int ddl_ar_oneshot(void* const* buf_ptrs, void* const* flag_ptrs, int is_bf16, void* out,
                   uint32_t* epoch, int* status, int64_t off, int64_t n, float scale, int rank,
                   int world, int blocks, int max_blocks, cudaStream_t st);
int ddl_peer_barrier(void* const* flag_ptrs, uint32_t* epoch, int* status, int rank, int world,
                     int max_blocks, cudaStream_t st);
int ddl_p2p_put(const void* src, void* dst_peer, int64_t nbytes, uint32_t* peer_flag,
                uint32_t* local_seq, unsigned int* ticket, int blocks, cudaStream_t st);
int ddl_p2p_wait(const uint32_t* flag, uint32_t* local_seq, int* status, cudaStream_t st);

// ---- misc.cu
int ddl_u8_to_bf16_norm(const void* in, void* out, int64_t n, int C, const float* mean,
                        const float* stdv, int sms, cudaStream_t st);
int ddl_revert_varlen(const void* in, void* out, const int64_t* lengths, int T, int B,
                      int64_t row_bytes, int sms, cudaStream_t st);
int ddl_set_mask(const int64_t* lengths, void* mask, int T, int B, cudaStream_t st);
int ddl_gap_fwd(const void* x, void* y, int N, int HW, int C, cudaStream_t st);
int ddl_gap_bwd(const void* dy, void* dx, int N, int HW, int C, int sms, cudaStream_t st);

// ---- elementwise.cu (average pooling, dropout)
int ddl_avgpool(int backward, const void* in, void* out, int N, int H, int W, int C, int Ho, int Wo,
                int k, int s, int p, int adaptive, int count_include_pad, int sms, cudaStream_t st);
int ddl_dropout_fwd(const void* x, void* y, void* mask, int64_t n, float p, uint64_t seed, uint64_t salt,
                    const long long* step, int sms, cudaStream_t st);
int ddl_dropout_bwd(const void* dy, const void* mask, void* dx, int64_t n, float scale, int sms,
                    cudaStream_t st);

// ---- pool.cu
int ddl_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, int Ho, int Wo,
                    int k, int s, int p, int sms, cudaStream_t st);
int ddl_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, int Ho,
                    int Wo, int k, int s, int p, int sms, cudaStream_t st);

// ---- gemm_tcgen05.cu  (TMA + tcgen05/TMEM GEMM; see the file header)
// D[M,N] = A[M,K] * B[N,K]^T  (both K-major bf16), fp32 accumulate in TMEM, bf16 output.
// Optional epilogue: per-column (output-channel) sum / sum-of-squares into col_sums[2][N]
// (BatchNorm statistics), optional bias[N] add and ReLU.
// `fin` (nullable): also finalize the BatchNorm statistics in the last CTA of the grid.
typedef struct DdlBnFinalize {
  unsigned int* ticket;   // grid arrival counter, zero on entry, self-resetting
  float* stat_out;        // [2][N] (mean, invstd)
  float* running_mean;    // [N] or NULL
  float* running_var;     // [N] or NULL
  float eps, momentum;
} DdlBnFinalize;
int ddl_gemm_bf16_tn(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                     int ldd, float* col_sums, const void* bias, int relu,
                     const DdlBnFinalize* fin, int sms, cudaStream_t st);
// D[M,N] (+)= A^T B with A stored [K,M] (M contiguous) and B stored [K,N] (N contiguous):
// both operands MN-major (weight gradient: A = dY[pixels, Cout], B = X[pixels, Cin]).
// Split-K over `splits` CTAs per tile with fp32 atomics into D32 when splits > 1.
int ddl_gemm_bf16_nt_mn(const void* A, const void* B, void* D, int M, int N, int K, int lda,
                        int ldb, int ldd, int out_is_bf16, int splits, float* workspace, int sms,
                        cudaStream_t st);
int ddl_gemm_bf16_tn_bnfold(const void* C, const void* R, const void* B, void* D, void* Y, void* mask,
                            int M, int N, int K, int ld, int ldb, int ldd, const float* stat,
                            const void* gamma, const void* beta, int param_is_bf16, int act,
                            float* col_sums, const void* bias, int relu, int sms, cudaStream_t st);
int ddl_gemm_bf16_tn_bnbwd(const void* dy, const void* dy2, const void* c, const void* mask,
                           const void* bt, void* dx, void* dc, void* dz, int M, int N, int K, int ld,
                           int ldb, int ldd, const float* stat, const float* coef, int act, int sms,
                           cudaStream_t st);

// ---- igemm_tcgen05.cu  (persistent implicit-GEMM: every dense conv fprop / dgrad and the FC layers)
// Return: 0 ok (with col_sums: sums only, run ddl_bn_finalize), DDL_IGEMM_FINALIZED (-100) ok and the
// statistics were finalized by the kernel, -1 unsupported shape, > 0 CUDA error.
// `fin->ticket` here is an ARRAY with one counter per 128-wide output-channel tile; splitk_*: zero
// fp32 scratch of `splitk_tiles` x 16384 floats + as many zero tickets (null / 0: never split K).
int ddl_igemm_gemm(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                   int ldd, int b_mn_major, float* col_sums, const void* bias, int relu,
                   const DdlBnFinalize* fin, float* splitk_ws, unsigned int* splitk_tickets,
                   int splitk_tiles, int sms, cudaStream_t st);
int ddl_igemm_conv_fprop(const void* x, const void* w, void* y, int NB, int H, int W, int Cin,
                         int Cout, int KH, int KW, int pad_h, int pad_w, int stride, float* col_sums,
                         const void* bias, int relu, const DdlBnFinalize* fin, float* splitk_ws,
                         unsigned int* splitk_tickets, int splitk_tiles, int sms, cudaStream_t st);
int ddl_igemm_conv_dgrad(const void* dy, const void* w, void* dx, int NB, int H, int W, int Cin,
                         int Cout, int KH, int KW, int pad_h, int pad_w, int stride,
                         float* splitk_ws, unsigned int* splitk_tickets, int splitk_tiles, int sms,
                         cudaStream_t st);
void ddl_igemm_set_trace(unsigned long long* buf);
void ddl_wgrad_set_trace(unsigned long long* buf);
// ---- gemm2cta_tcgen05.cu  (CTA-pair `cta_group::2` persistent GEMM, 256 x 256 tiles; -1: declined)
int ddl_gemm2cta(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                 int ldd, const void* bias, int relu, int sms, cudaStream_t st);
// ---- gemm_wgrad_tcgen05.cu, one-launch weight gradients (self-cleaning fp32 workspace)
int ddl_wgrad_gemm(const void* A, const void* B, float* ws, long long ws_floats, void* out,
                   int out_bf16, int accumulate, int M, int N, int K, int lda, int ldb, int ldd,
                   int sms, cudaStream_t st);
int ddl_wgrad_conv(const void* dy, const void* x, float* ws, long long ws_floats, void* out,
                   int out_bf16, int accumulate, int NB, int H, int W, int Cin, int Cout, int KH,
                   int KW, int pad_h, int pad_w, int stride, int sms, cudaStream_t st);
int ddl_gemm_selftest_info(char* buf, int len);
int ddl_set_pdl(int on);

#ifdef __cplusplus
}
#endif
