// Declarations of the fp32 SIMT kernels (simt.cu).
#pragma once
#include "common.cuh"

namespace mst {

enum AMode { A_ROWMAJOR = 0, A_MOTION = 1, A_GATHER_ROWS = 2 };
enum Epi { EPI_PLAIN = 0, EPI_GELU, EPI_SILU, EPI_RESIDUAL, EPI_INPROJ, EPI_OUTPROJ };

struct GemmF32Params {
  const float* a = nullptr;   // A operand (see a_mode)
  const float* w = nullptr;   // [N, ldw] row-major (nn.Linear weight layout)
  const float* bias = nullptr;
  float* c = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0, ldc = 0;
  int a_mode = A_ROWMAJOR;
  int epi = EPI_PLAIN;
  const int64_t* gather = nullptr;  // A_GATHER_ROWS: row index per m
  const float* residual = nullptr;  // EPI_RESIDUAL, same layout as c
  const float* pe = nullptr;        // EPI_INPROJ: positional table [*, N]
  int B = 0, T = 0;                 // A_MOTION / EPI_INPROJ / EPI_OUTPROJ geometry
  int n_pass = 1;                   // EPI_INPROJ: write rows of `n_pass` passes (cond, uncond)
  int row_invariant = 0;            // 1: always the one-warp-per-output kernel, so a row's result does not depend on how
                                    // many rows share the launch (time / text embeddings: bit-identical for any batch split)
};

int gemm_f32(const GemmF32Params& p, cudaStream_t s);

struct Token0Params {
  const float* temb = nullptr;
  const int32_t* temb_row_dev = nullptr;
  int temb_row_offset = 0;
  const float* text_emb = nullptr;
  const float* txt_b = nullptr;  // embed_text.bias (== embed_text(0)); NULL when the model has no text branch
  const float* pe = nullptr;
  float* x_f32 = nullptr;
  __nv_bfloat16* x_bf16 = nullptr;
  __half* x_f16 = nullptr;  // fp16 residual stream of the sampler
  int B = 0, T = 0, d = 0, cfg = 0, uncond = 0;
};

int token0(const Token0Params& p, int n_seqs, cudaStream_t s);
int layernorm_f32(const float* x, const float* g, const float* b, float* y, int M, int d, cudaStream_t s);
int attention_f32(const float* qkv, float* out, int n_seqs, int S, int d, int n_heads, cudaStream_t s);

}  // namespace mst
