// Declarations of the tcgen05 / TMEM / TMA kernels (tc_gemm.cu, tc_attn.cu).
#pragma once
#include "common.cuh"
#include "dropout.cuh"

namespace mst {

// Epilogues of the bf16 tensor-core GEMM  D[M,N] = A[M,K] * W[N,K]^T  (+ ...)
enum TcEpi {
  TC_EPI_BIAS_BF16 = 0,   // out bf16 [M, ldo]  = acc + bias                     (QKV)
  TC_EPI_BIAS_GELU_BF16,  // out bf16 [M, ldo]  = gelu(acc + bias)               (FFN1)
  TC_EPI_BIAS_RES_LN,     // out bf16 [M, N]    = LN(acc + bias + residual)*g+b  (N == 512; out-proj, FFN2)
  TC_EPI_INPROJ,          // token rows (b, t+1) of every pass = acc + bias + pe[t+1]
  TC_EPI_OUTPROJ_F32,     // out fp32 [seq][n][s-1] = acc + bias (token 0 dropped, n < n_valid)
  TC_EPI_BIAS_F32,        // out fp32 [M, ldo]  = acc + bias                     (test hook)
  TC_EPI_TRAIN_F32,       // out fp32 [M, ldo]  (+)= acc (+ bias) (+ add)        (training GEMMs: forward, dX, dW)
};

struct TcGemmParams {
  const __nv_bfloat16* a = nullptr;  // [M, K] row-major, K % 64 == 0, 16B-aligned rows
  const __nv_bfloat16* w = nullptr;  // [N, K] row-major
  const float* bias = nullptr;       // [N]
  void* out = nullptr;
  int M = 0, N = 0, K = 0;
  int ldo = 0;
  int epi = TC_EPI_BIAS_BF16;
  // LN epilogue
  const __nv_bfloat16* residual = nullptr;  // [M, N]
  const float* ln_g = nullptr;
  const float* ln_b = nullptr;
  // in/out projection geometry
  const float* pe = nullptr;
  int B = 0, T = 0, n_pass = 1, n_valid = 0;
  int td_mode = 0;           // experiment knob (MST_TEARDOWN)
  long long* dbg = nullptr;  // test hook: clock64 timeline of cluster 0 (see mst_test_set_gemm_debug)
  float* out2 = nullptr;  // OUTPROJ: rows of sequences >= B go here (uncond pass)
  const float* add = nullptr;  // TRAIN_F32: fp32 tensor with out's layout added to the result (residual paths)
  int accumulate = 0;          // TRAIN_F32: out += result (gradient accumulation)
  // TRAIN_F32 as the in-projection of the taped forward: GEMM row (b, t), t < tok_T, lands in output row
  // b * (tok_T + tok_off) + t + tok_off (+ pe[t + tok_off] when pe is given); tok_T == 0: rows map one to one
  int tok_T = 0, tok_off = 0;
  // OUTPROJ_F32: leading token rows of every sequence that carry no output (0 = the default, one)
  int drop_tokens = 0;
  // TRAIN_F32, linear1 of the taped forward: besides out = u, also h = dropout(gelu(u)) as fp32 (tape) and bf16 (operand
  // of linear2) - the separate GELU launch of every layer goes away.  Rows are grouped rows_per_seq per sequence.
  float* gelu_out = nullptr;
  __nv_bfloat16* gelu_bf = nullptr;
  Drop drop;
  uint32_t drop_site = 0;
  int rows_per_seq = 1;
  // 16-bit format switches of the sampler's residual stream (see tc_gemm_ln5_kernel): IEEE fp16 instead of bf16
  int ab_f16 = 0;  // A and W hold fp16 (QKV, FFN1 and the final projection read the fp16 stream)
  int io_f16 = 0;  // RES_LN: residual and output are fp16;  INPROJ: output is fp16
};

int tc_gemm(const TcGemmParams& p, cudaStream_t s);
void set_gemm_debug(long long* dev_buf);  // device buffer of >= 3*2*1024 int64 (or nullptr to switch off)

// x [B,F,T] fp32 -> A operand of the in-projection: bf16 [B*T, f_pad], zero padded
struct Token0Params;
// ... and, when t0 is given, token 0 of the n_seqs sequences in the same launch (see simt.cuh)
int motion_to_tokens_bf16(const float* x, __nv_bfloat16* a, int B, int F, int T, int f_pad, const Token0Params* t0,
                          int n_seqs, cudaStream_t s);

// fp32 [rows, cols] -> bf16 [rows_pad, cols_pad] zero padded (weight packing)
int pack_bf16(const float* src, __nv_bfloat16* dst, int rows, int cols, int rows_pad, int cols_pad, cudaStream_t s);

// the same into IEEE fp16
int pack_f16(const float* src, __half* dst, int rows, int cols, int rows_pad, int cols_pad, cudaStream_t s);

// fp32 [rows, cols] (leading dimension ld) -> bf16 copy dst [rows, cols] and / or transposed copy dst_t [cols, rows_pad]
// (columns [rows, rows_pad) zero): operands of the training GEMMs (dX needs W^T, dW needs dY^T and X^T)
int cvt_bf16(const float* src, int rows, int cols, int ld, __nv_bfloat16* dst, __nv_bfloat16* dst_t, int rows_pad,
             cudaStream_t s);

// Several conversions in one launch: each job turns fp32 src [rows, cols] (leading dimension ld) into any of
//   dst   bf16 [rows_pad, cols_pad] zero padded,      dst_h  fp16 [rows_pad, cols_pad] zero padded,
//   dst_t bf16 [cols, t_ld] transposed, columns [rows, t_ld) zero,      colsum[c] += sum_r src[r, c].
struct CvtJob {
  const float* src;
  __nv_bfloat16* dst;
  __nv_bfloat16* dst_t;
  __half* dst_h;
  float* colsum;
  int rows, cols, ld, rows_pad, cols_pad, t_ld, tile0, tiles_x;
};
constexpr int CVT_MAX_JOBS = 40;
struct CvtJobs {
  CvtJob j[CVT_MAX_JOBS];
  int n = 0;
  int tiles_of_last = 0;
};
void cvt_jobs_add(CvtJobs& js, const float* src, int rows, int cols, int ld, __nv_bfloat16* dst, __half* dst_h, int rows_pad,
                  int cols_pad, __nv_bfloat16* dst_t, int t_ld, float* colsum);
int cvt_multi(const CvtJobs& js, cudaStream_t s, const char* name);

struct TcAttnParams {
  const __nv_bfloat16* qkv = nullptr;  // [n_seqs*S, 3d]  (Q | K | V column blocks)
  __nv_bfloat16* out = nullptr;        // [n_seqs*S, d]
  int n_seqs = 0, S = 0, d_model = 0, n_heads = 0;
};
int tc_attention(const TcAttnParams& p, cudaStream_t s);

}  // namespace mst
