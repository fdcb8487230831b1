// fp32 SIMT kernels of the denoiser (MST_PREC_FP32, the parity mode) and the
// small fp32 helpers both precisions share (time/text embedding, token 0).
//
// What they restate (reference file:line):
//   InputProcess            model/mdm_forstyledataset.py:435-441
//   PositionalEncoding add  model/mdm_forstyledataset.py:401-404
//   TimestepEmbedder        model/mdm_forstyledataset.py:415-422
//   embed_text / mask_cond  model/mdm_forstyledataset.py:288-296, :327
//   TransformerEncoderLayer torch semantics as used at :231-238 (post-norm,
//                           exact GELU, softmax(QK^T/sqrt(dh))V, LN eps 1e-5)
//   OutputProcess           model/mdm_forstyledataset.py:464-478
#include "common.cuh"
#include "simt.cuh"

namespace mst {

// ---------------------------------------------------------------------------
// Generic fp32 GEMM  C[m,n] = sum_k A(m,k) * W[n*ldw + k]   (+ fused epilogue)
// 128x128x16 tile, 256 threads, 8x8 outputs per thread.
// ---------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float load_a(const GemmF32Params& p, int m, int k) {
  switch (p.a_mode) {
    case A_ROWMAJOR: return p.a[(int64_t)m * p.lda + k];
    case A_MOTION: {  // A(m=(b,t), k=f) = x[b][f][t]
      int b = m / p.T, t = m - b * p.T;
      return p.a[((int64_t)b * p.K + k) * p.T + t];
    }
    case A_GATHER_ROWS: return p.a[(int64_t)p.gather[m] * p.lda + k];
    default: return 0.0f;
  }
}

__device__ __forceinline__ void store_c(const GemmF32Params& p, int m, int n, float v) {
  if (p.bias) v += p.bias[n];
  switch (p.epi) {
    case EPI_PLAIN: p.c[(int64_t)m * p.ldc + n] = v; break;
    case EPI_GELU: p.c[(int64_t)m * p.ldc + n] = gelu_erf(v); break;
    case EPI_SILU: p.c[(int64_t)m * p.ldc + n] = silu(v); break;
    case EPI_RESIDUAL: p.c[(int64_t)m * p.ldc + n] = v + p.residual[(int64_t)m * p.ldc + n]; break;
    case EPI_INPROJ: {  // m=(b,t) -> token row (b, t+1) of every pass, + pe[t+1]
      int b = m / p.T, t = m - b * p.T;
      v += p.pe[(int64_t)(t + 1) * p.N + n];
      int S = p.T + 1;
      for (int pass = 0; pass < p.n_pass; ++pass)
        p.c[((int64_t)(pass * p.B + b) * S + t + 1) * p.ldc + n] = v;
      break;
    }
    case EPI_OUTPROJ: {  // m=(seq,s) -> out[seq][n][s-1], token 0 dropped
      int S = p.T + 1;
      int seq = m / S, s = m - seq * S;
      if (s > 0) p.c[((int64_t)seq * p.N + n) * p.T + (s - 1)] = v;
      break;
    }
  }
}

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmF32Params p) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // each thread loads 8 A and 8 W elements of the [128 x 16] tiles
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      int idx = tid + l * 256;       // 0..2047
      int kk, mm;
      if (p.a_mode == A_MOTION) { mm = idx & 127; kk = idx >> 7; }   // m fastest: coalesced along t
      else { kk = idx & 15; mm = idx >> 4; }                           // k fastest: coalesced along k
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < p.K) ? load_a(p, m, k) : 0.0f;
      int kw = idx & 15, nn = idx >> 4;
      int n = n0 + nn, k2 = k0 + kw;
      Ws[kw][nn] = (n < p.N && k2 < p.K) ? p.w[(int64_t)n * p.ldw + k2] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = Ws[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + tx + 16 * j;
      if (n < p.N) store_c(p, m, n, acc[i][j]);
    }
  }
}

// Few output rows (time / text embedding of a handful of samples, the B=1 style example of the finetune step): the
// 128 x 128 tiling would run on a handful of CTAs with a long serial k-loop; one warp per output element instead.
__global__ void __launch_bounds__(256) gemm_f32_skinny_kernel(GemmF32Params p) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const int lane = threadIdx.x & 31;
  const long long total = (long long)p.M * p.N;
  for (long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < total;
       o += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int m = (int)(o / p.N), n = (int)(o - (long long)m * p.N);
    const float* w = p.w + (int64_t)n * p.ldw;
    float acc = 0.0f;
    for (int k = lane; k < p.K; k += 32) acc = fmaf(load_a(p, m, k), w[k], acc);
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) store_c(p, m, n, acc);
  }
}

int gemm_f32(const GemmF32Params& p, cudaStream_t s) {
  if (p.row_invariant || ((long long)p.M * p.N <= 65536 && p.M <= 128 && (long long)p.M * p.N * p.K <= (24ll << 20))) {
    const long long warps = (long long)p.M * p.N;
    const int blocks = (int)((warps + 7) / 8 < 4096 ? (warps + 7) / 8 : 4096);
    MST_CUDA_OK(launch_pdl(gemm_f32_skinny_kernel, dim3(blocks), dim3(256), 0, s, p));
    MST_LAUNCHED("gemm_f32_skinny", s);
    return MST_OK;
  }
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
  MST_CUDA_OK(launch_pdl(gemm_f32_kernel, grid, dim3(256), 0, s, p));
  MST_LAUNCHED("gemm_f32", s);
  return MST_OK;
}

// ---------------------------------------------------------------------------
// token 0 of every sequence:  temb[row] + (uncond ? txt_b : text_emb[b]) + pe[0]
// (mdm_forstyledataset.py:322-327, :344-345).  Writes fp32 and/or bf16.
// ---------------------------------------------------------------------------
__global__ void token0_kernel(Token0Params p) {
  pdl_launch_dependents();
  pdl_wait();  // *temb_row_dev is decremented by the previous step's update kernel
  const int seq = blockIdx.x;
  const int b = seq % p.B;
  const bool uncond = p.cfg ? (seq >= p.B) : (p.uncond != 0);
  int row = p.temb_row_dev ? (*p.temb_row_dev + p.temb_row_offset) : (b + p.temb_row_offset);
  const int S = p.T + 1;
  for (int n = threadIdx.x; n < p.d; n += blockDim.x) {
    float v = p.temb[(int64_t)row * p.d + n];
    if (p.txt_b) v += (uncond || !p.text_emb) ? p.txt_b[n] : p.text_emb[(int64_t)b * p.d + n];
    v += p.pe[n];
    int64_t o = (int64_t)seq * S * p.d + n;
    if (p.x_f32) p.x_f32[o] = v;
    if (p.x_bf16) p.x_bf16[o] = __float2bfloat16_rn(v);
    if (p.x_f16) p.x_f16[o] = __float2half_rn(v);
  }
}

int token0(const Token0Params& p, int n_seqs, cudaStream_t s) {
  MST_CUDA_OK(launch_pdl(token0_kernel, dim3(n_seqs), dim3(128), 0, s, p));
  MST_LAUNCHED("token0", s);
  return MST_OK;
}

// ---------------------------------------------------------------------------
// y = LayerNorm(x) * g + b over rows of width d (d % 32 == 0, d <= 1024);
// one warp per row, two-pass mean / biased variance, eps = 1e-5.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                            const float* __restrict__ bt, float* __restrict__ y, int M,
                                                            int d) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + (int64_t)warp * d;
  float v[32];
  const int per = d / 32;
  float s = 0.0f;
  for (int i = 0; i < per; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)d;
  float q = 0.0f;
  for (int i = 0; i < per; ++i) { float c = v[i] - mean; q += c * c; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)d + 1e-5f);
  float* yr = y + (int64_t)warp * d;
  for (int i = 0; i < per; ++i) {
    int n = lane + 32 * i;
    yr[n] = (v[i] - mean) * rstd * g[n] + bt[n];
  }
}

int layernorm_f32(const float* x, const float* g, const float* b, float* y, int M, int d, cudaStream_t s) {
  if (d % 32 != 0 || d > 1024) return fail(MST_ERR_UNSUPPORTED, "layernorm_f32: d must be a multiple of 32 and <= 1024");
  int warps_per_block = 8;
  MST_CUDA_OK(launch_pdl(layernorm_f32_kernel, dim3(ceil_div(M, warps_per_block)), dim3(warps_per_block * 32), 0, s, x, g, b, y, M, d));
  MST_LAUNCHED("layernorm_f32", s);
  return MST_OK;
}

// ---------------------------------------------------------------------------
// fp32 attention: softmax(Q K^T / sqrt(dh)) V per (sequence, head).
// qkv: [n_seqs*S, 3d]  (Q | K | V column blocks, head h at columns h*dh..)
// One CTA = (q-chunk of 16 rows, head, sequence); K then V staged in smem.
// ---------------------------------------------------------------------------
constexpr int ATT_QROWS = 16;

__global__ void __launch_bounds__(128) attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int S,
                                                            int d, int dh, float scale) {
  extern __shared__ float sm[];
  const int ldk = dh + 1;
  float* kv = sm;                       // [S][dh+1]
  float* qs = kv + (size_t)S * ldk;     // [ATT_QROWS][dh]
  float* ps = qs + ATT_QROWS * dh;      // [ATT_QROWS][S]
  const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_QROWS;
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)seq * S;
  const int ld = 3 * d;
  const int nq = min(ATT_QROWS, S - q0);

  for (int idx = tid; idx < S * dh; idx += blockDim.x) {
    int j = idx / dh, c = idx - j * dh;
    kv[j * ldk + c] = qkv[(row0 + j) * ld + d + h * dh + c];
  }
  for (int idx = tid; idx < ATT_QROWS * dh; idx += blockDim.x) {
    int i = idx / dh, c = idx - i * dh;
    qs[idx] = (i < nq) ? qkv[(row0 + q0 + i) * ld + h * dh + c] * scale : 0.0f;
  }
  __syncthreads();
  for (int idx = tid; idx < nq * S; idx += blockDim.x) {
    int i = idx / S, j = idx - i * S;
    const float* qr = qs + i * dh;
    const float* kr = kv + j * ldk;
    float acc = 0.0f;
    for (int c = 0; c < dh; ++c) acc = fmaf(qr[c], kr[c], acc);
    ps[i * S + j] = acc;
  }
  __syncthreads();
  // softmax per row: one warp per row
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < nq; i += 4) {
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, ps[i * S + j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int j = lane; j < S; j += 32) { float e = expf(ps[i * S + j] - mx); ps[i * S + j] = e; sum += e; }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float inv = 1.0f / sum;
    for (int j = lane; j < S; j += 32) ps[i * S + j] *= inv;
  }
  __syncthreads();
  for (int idx = tid; idx < S * dh; idx += blockDim.x) {
    int j = idx / dh, c = idx - j * dh;
    kv[j * ldk + c] = qkv[(row0 + j) * ld + 2 * d + h * dh + c];
  }
  __syncthreads();
  for (int c = tid; c < dh; c += blockDim.x) {
    for (int i = 0; i < nq; ++i) {
      float acc = 0.0f;
      const float* pr = ps + i * S;
      for (int j = 0; j < S; ++j) acc = fmaf(pr[j], kv[j * ldk + c], acc);
      out[(row0 + q0 + i) * d + h * dh + c] = acc;
    }
  }
}

int attention_f32(const float* qkv, float* out, int n_seqs, int S, int d, int n_heads, cudaStream_t s) {
  const int dh = d / n_heads;
  size_t smem = ((size_t)S * (dh + 1) + (size_t)ATT_QROWS * dh + (size_t)ATT_QROWS * S) * sizeof(float);
  if (smem > 220 * 1024) return fail(MST_ERR_UNSUPPORTED, "attention_f32: sequence too long for shared memory");
  static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
  if (attr_set.first()) {
    MST_CUDA_OK(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  }
  dim3 grid(ceil_div(S, ATT_QROWS), n_heads, n_seqs);
  attention_f32_kernel<<<grid, 128, smem, s>>>(qkv, out, S, d, dh, 1.0f / sqrtf((float)dh));
  MST_LAUNCHED("attention_f32", s);
  return MST_OK;
}

}  // namespace mst
