// Training path of the denoiser: forward with an activation tape, backward
// through OutputProcess -> 8 x TransformerEncoderLayer -> InputProcess, the masked-L2 style loss, the gradient of
// the per-step update, and the fused multi-tensor AdamW / norm kernels.  fp32 SIMT (the parity precision).
//
// What it restates (reference file:line):
//   few_shot_style_finetune_losses      diffusion/gaussian_diffusion.py:1317-1399
//   masked_l2                           diffusion/gaussian_diffusion.py:223-235
//   p_sample_with_grad / ddim_..._grad  diffusion/inpainting_gaussian_diffusion.py:66-123, :176-239
//   StyleDiffusion.forward              model/mdm_forstyledataset.py:602-625   (torch autograd does the backward there)
//   MotionEncoder.forward               model/mdm_forstyledataset.py:89-124
//   MixedPrecisionTrainer._compute_norms / AdamW step   diffusion/fp16_util.py:208-223, train/training_loop.py:97-99
//
// Layer algebra (nn.TransformerEncoderLayer, post-norm, exact GELU; dropout is identity: see DESIGN.md):
//   qkv = x Wqkv^T + b ; P = softmax(Q K^T / sqrt(dh)) ; ao = P V ; z1 = x + ao Wo^T + bo ; y = LN1(z1)
//   u = y W1^T + b1 ; h = gelu(u) ; z2 = y + h W2^T + b2 ; x' = LN2(z2)
#include "common.cuh"
#include "simt.cuh"
#include "tc.cuh"
#include "smem_gemm.cuh"
#include "dropout.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace mst {

// ---------------------------------------------------------------------------------------------------------
// General batched fp32 GEMM   C = alpha * op(A) op(B) (+ bias) (+ add) (+ C)
// ---------------------------------------------------------------------------------------------------------
enum AModeEx { AX_NORMAL = 0, AX_TRANS = 1, AX_MOTION_TOK = 2, AX_TOKROWS = 3 };
enum CModeEx { CX_NORMAL = 0, CX_MOTION = 1 };

struct GemmEx {
  const float* a = nullptr;
  const float* b = nullptr;
  float* c = nullptr;
  const float* bias = nullptr;  // [N]
  const float* add = nullptr;   // same layout / leading dimension as c
  int M = 0, N = 0, K = 0, lda = 0, ldb = 0, ldc = 0;
  int a_mode = AX_NORMAL;  // NORMAL a[m*lda+k] | TRANS a[k*lda+m] | MOTION_TOK m=(seq,s): s<tok_off ? 0 : a[(seq*K+k)*T+s-tok_off]
                           // | TOKROWS m=(seq,t): a[(seq*S+t+tok_off)*lda+k]
  int trans_b = 0;         // 0: B(k,n) = b[k*ldb+n]   1: B(k,n) = b[n*ldb+k]
  int c_mode = CX_NORMAL;  // MOTION: m=(seq,t) -> c[(seq*N+n)*T+t]
  int T = 0, tok_off = 1;  // token geometry of the MOTION / TOKROWS modes (S = T + tok_off)
  float alpha = 1.0f;
  int accumulate = 0;
  int batch = 1, heads = 1;  // blockIdx.z -> (z / heads, z % heads)
  long long a_bs = 0, a_hs = 0, b_bs = 0, b_hs = 0, c_bs = 0, c_hs = 0;
  int split_k = 1;  // > 1: partial products are atomically added into c (which the caller zeroed or accumulates into)
};

template <int TM>
__global__ void __launch_bounds__(256) gemm_ex_kernel(GemmEx p) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  constexpr int BT = 16 * TM;  // square tile
  constexpr int BK = 16;
  __shared__ float As[BK][BT + 4];
  __shared__ float Bs[BK][BT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z / p.split_k, split = blockIdx.z - z * p.split_k;
  const int zb = z / p.heads, zh = z - zb * p.heads;
  const float* __restrict__ a = p.a + zb * p.a_bs + zh * p.a_hs;
  const float* __restrict__ b = p.b + zb * p.b_bs + zh * p.b_hs;
  float* __restrict__ c = p.c + zb * p.c_bs + zh * p.c_hs;
  const float* __restrict__ addp = p.add ? p.add + zb * p.c_bs + zh * p.c_hs : nullptr;
  const int m0 = blockIdx.y * BT, n0 = blockIdx.x * BT;
  const int k_per = ((p.K + p.split_k - 1) / p.split_k + BK - 1) / BK * BK;
  const int k_begin = split * k_per, k_end = min(p.K, k_begin + k_per);
  const int S = p.T + p.tok_off;

  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.0f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int l = 0; l < (BT * BK) / 256; ++l) {
      const int idx = tid + l * 256;
      int mm, kk;
      if (p.a_mode == AX_NORMAL || p.a_mode == AX_TOKROWS) { kk = idx & 15; mm = idx >> 4; }
      else { mm = idx % BT; kk = idx / BT; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.0f;
      if (m < p.M && k < k_end) {
        switch (p.a_mode) {
          case AX_NORMAL: v = a[(long long)m * p.lda + k]; break;
          case AX_TRANS: v = a[(long long)k * p.lda + m]; break;
          case AX_MOTION_TOK: {
            const int seq = m / S, s = m - seq * S;
            v = s < p.tok_off ? 0.0f : a[((long long)seq * p.K + k) * p.T + (s - p.tok_off)];
            break;
          }
          default: {
            const int seq = m / p.T, t = m - seq * p.T;
            v = a[((long long)seq * S + t + p.tok_off) * p.lda + k];
          }
        }
      }
      As[kk][mm] = v;
      int nn, kb;
      if (p.trans_b) { kb = idx & 15; nn = idx >> 4; }
      else { nn = idx % BT; kb = idx / BT; }
      const int n = n0 + nn, k2 = k0 + kb;
      float w = 0.0f;
      if (n < p.N && k2 < k_end) w = p.trans_b ? b[(long long)n * p.ldb + k2] : b[(long long)k2 * p.ldb + n];
      Bs[kb][nn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TM; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (k_begin >= k_end && split != 0) return;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= p.N) continue;
      long long o;
      if (p.c_mode == CX_MOTION) {
        const int seq = m / p.T, t = m - seq * p.T;
        o = ((long long)seq * p.N + n) * p.T + t;
      } else {
        o = (long long)m * p.ldc + n;
      }
      float v = p.alpha * acc[i][j];
      if (split == 0) {
        if (p.bias) v += p.bias[n];
        if (addp) v += addp[o];
      }
      if (p.split_k > 1) atomicAdd(c + o, v);
      else if (p.accumulate) c[o] += v;
      else c[o] = v;
    }
  }
}

__device__ __forceinline__ float gemm_ex_a(const GemmEx& p, const float* a, int m, int k, int S) {
  switch (p.a_mode) {
    case AX_NORMAL: return a[(long long)m * p.lda + k];
    case AX_TRANS: return a[(long long)k * p.lda + m];
    case AX_MOTION_TOK: {
      const int seq = m / S, s = m - seq * S;
      return s < p.tok_off ? 0.0f : a[((long long)seq * p.K + k) * p.T + (s - p.tok_off)];
    }
    default: {
      const int seq = m / p.T, t = m - seq * p.T;
      return a[((long long)seq * S + t + p.tok_off) * p.lda + k];
    }
  }
}

// Small problems (the B=1 style example: 77-token sequences, per-head 77 x 77 products): one warp per output element,
// lanes stride the reduction - hundreds of warps in flight instead of a dozen CTAs with a serial k-loop.
__global__ void __launch_bounds__(256) gemm_ex_skinny_kernel(GemmEx p) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const int lane = threadIdx.x & 31;
  const int S = p.T + p.tok_off;
  const long long per = (long long)p.M * p.N, total = per * p.batch * p.heads;
  for (long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < total;
       o += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int z = (int)(o / per);
    const long long r = o - (long long)z * per;
    const int m = (int)(r / p.N), n = (int)(r - (long long)m * p.N);
    const int zb = z / p.heads, zh = z - zb * p.heads;
    const float* a = p.a + zb * p.a_bs + zh * p.a_hs;
    const float* b = p.b + zb * p.b_bs + zh * p.b_hs;
    float acc = 0.0f;
    if (p.trans_b) {
      const float* br = b + (long long)n * p.ldb;
      for (int k = lane; k < p.K; k += 32) acc = fmaf(gemm_ex_a(p, a, m, k, S), br[k], acc);
    } else {
      for (int k = lane; k < p.K; k += 32) acc = fmaf(gemm_ex_a(p, a, m, k, S), b[(long long)k * p.ldb + n], acc);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      long long oo;
      if (p.c_mode == CX_MOTION) {
        const int seq = m / p.T, t = m - seq * p.T;
        oo = ((long long)seq * p.N + n) * p.T + t;
      } else {
        oo = (long long)m * p.ldc + n;
      }
      oo += zb * p.c_bs + zh * p.c_hs;
      float v = p.alpha * acc;
      if (p.bias) v += p.bias[n];
      if (p.add) v += p.add[oo];
      if (p.accumulate) p.c[oo] += v;
      else p.c[oo] = v;
    }
  }
}

static int gemm_ex(GemmEx p, cudaStream_t s, const char* name) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return fail(MST_ERR_INVALID, "gemm_ex: empty problem");
  const long long z = (long long)p.batch * p.heads;
  if ((long long)p.M * p.N * z <= 131072 && (long long)p.M * p.N * z * p.K <= (24ll << 20)) {
    const long long warps = (long long)p.M * p.N * z;
    const int blocks = (int)((warps + 7) / 8 < 8192 ? (warps + 7) / 8 : 8192);
    MST_CUDA_OK(launch_pdl(gemm_ex_skinny_kernel, dim3(blocks), dim3(256), 0, s, p));
    MST_LAUNCHED(name, s);
    return MST_OK;
  }
  if (p.split_k == 1 && z == 1 && p.c_mode == CX_NORMAL && (p.accumulate || p.ldc == p.N) && p.K >= 512 &&
      ceil_div(p.M, 64) * ceil_div(p.N, 64) < 64) {
    // a handful of tiles with a long reduction (fp32 linears of the B=1 style example): spread the k-loop over CTAs
    p.split_k = p.K / 128 < 8 ? p.K / 128 : 8;
  }
  const long long tiles128 = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 128) * z * p.split_k;
  if (p.split_k > 1 && !p.accumulate) {
    if (p.c_mode != CX_NORMAL || z != 1 || p.ldc != p.N)
      return fail(MST_ERR_INVALID, "gemm_ex: split-k overwrite needs a dense single output");
    MST_CUDA_OK(cudaMemsetAsync(p.c, 0, (size_t)p.M * p.N * sizeof(float), s));
  }
  static const long long t128 = getenv("MST_GEMM_EX_T128") ? atoll(getenv("MST_GEMM_EX_T128")) : 400;
  if (tiles128 >= t128) {
    dim3 grid(ceil_div(p.N, 128), ceil_div(p.M, 128), (unsigned)(z * p.split_k));
    MST_CUDA_OK(launch_pdl(gemm_ex_kernel<8>, grid, dim3(256), 0, s, p));
  } else {
    dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), (unsigned)(z * p.split_k));
    MST_CUDA_OK(launch_pdl(gemm_ex_kernel<4>, grid, dim3(256), 0, s, p));
  }
  MST_LAUNCHED(name, s);
  return MST_OK;
}

// split factor for a weight-gradient GEMM (small output, long reduction over the token rows)
static int pick_split(int M, int N, int K) {
  const int tiles = ceil_div(M, 128) * ceil_div(N, 128);
  int split = 1;
  while (tiles * split < 2 * sm_count() && K / (split * 2) >= 64 && split < 32) split *= 2;
  return split;
}

// ---------------------------------------------------------------------------------------------------------
// row kernels: softmax, softmax backward, LayerNorm backward, GELU forward / backward, column sums
// ---------------------------------------------------------------------------------------------------------
// rows of scores [rows, S] -> softmax in place; key_valid [n_seqs, S] (1 = attend) or NULL; row -> seq = row / (heads*S)
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ p, const uint8_t* __restrict__ key_valid,
                                                           long long rows, int S, int rows_per_seq) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = p + row * S;
  const uint8_t* kv = key_valid ? key_valid + (row / rows_per_seq) * S : nullptr;
  float mx = -INFINITY;
  for (int j = lane; j < S; j += 32)
    if (!kv || kv[j]) mx = fmaxf(mx, r[j]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
  for (int j = lane; j < S; j += 32) {
    const float e = (!kv || kv[j]) ? expf(r[j] - mx) : 0.0f;
    r[j] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  for (int j = lane; j < S; j += 32) r[j] *= inv;
}

// dS = P * (dP - sum_j P dP), in place in dP
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* __restrict__ p, float* __restrict__ dp,
                                                               long long rows, int S) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* pr = p + row * S;
  float* dr = dp + row * S;
  float dot = 0.0f;
  for (int j = lane; j < S; j += 32) dot = fmaf(pr[j], dr[j], dot);
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int j = lane; j < S; j += 32) dr[j] = pr[j] * (dr[j] - dot);
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ u, float* __restrict__ h, long long n) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    h[i] = gelu_f(u[i]);
}
// du = dh * gelu'(u), in place in dh
// (+ a bf16 copy: the A operand of the dX GEMM that follows, when dh_bf is given)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ u, float* __restrict__ dh,
                                                       __nv_bfloat16* __restrict__ dh_bf, long long n) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = dh[i] * gelu_grad_f(u[i]);
    dh[i] = g;
    if (dh_bf) dh_bf[i] = __float2bfloat16_rn(g);
  }
}

// out[i] = in[i] * mask[i] (+ add[i]); n % 4 == 0, all pointers 16-byte aligned; in may alias out
__global__ void __launch_bounds__(256) dropout_kernel(const float* in, const float* __restrict__ add, float* out, long long n4,
                                                      long long per_seq4, Drop d, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x) {
    const int seq = (int)(v / per_seq4);
    const float4 m = drop_scale4(d, site, seq, v - (long long)seq * per_seq4);
    float4 x = reinterpret_cast<const float4*>(in)[v];
    x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
    if (add) {
      const float4 a = reinterpret_cast<const float4*>(add)[v];
      x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
    }
    reinterpret_cast<float4*>(out)[v] = x;
  }
}

// du = dh * mask * gelu'(u), in place in dh
__global__ void __launch_bounds__(256) gelu_bwd_drop_kernel(const float* __restrict__ u, float* __restrict__ dh,
                                                            __nv_bfloat16* __restrict__ dh_bf, long long n4,
                                                            long long per_seq4, Drop d, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x) {
    const int seq = (int)(v / per_seq4);
    const float4 m = drop_scale4(d, site, seq, v - (long long)seq * per_seq4);
    const float4 uu = reinterpret_cast<const float4*>(u)[v];
    float4 g = reinterpret_cast<float4*>(dh)[v];
    g.x *= m.x * gelu_grad_f(uu.x); g.y *= m.y * gelu_grad_f(uu.y);
    g.z *= m.z * gelu_grad_f(uu.z); g.w *= m.w * gelu_grad_f(uu.w);
    reinterpret_cast<float4*>(dh)[v] = g;
    if (dh_bf) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(g.x, g.y), hi = __floats2bfloat162_rn(g.z, g.w);
      reinterpret_cast<uint2*>(dh_bf)[v] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
}

__device__ __forceinline__ uint2 bf16x4(const float4& x) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// h = dropout(gelu(u)) (+ a bf16 copy: the A operand of the next tensor-core GEMM) in one pass
__global__ void __launch_bounds__(256) gelu_drop_fwd_kernel(const float* __restrict__ u, float* __restrict__ h,
                                                            __nv_bfloat16* __restrict__ h_bf, long long n4, long long per_seq4,
                                                            Drop d, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x) {
    const float4 uu = reinterpret_cast<const float4*>(u)[v];
    float4 x = make_float4(gelu_f(uu.x), gelu_f(uu.y), gelu_f(uu.z), gelu_f(uu.w));
    if (d.on()) {
      const int seq = (int)(v / per_seq4);
      const float4 m = drop_scale4(d, site, seq, v - (long long)seq * per_seq4);
      x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
    }
    reinterpret_cast<float4*>(h)[v] = x;
    if (h_bf) reinterpret_cast<uint2*>(h_bf)[v] = bf16x4(x);
  }
}

// z = res + dropout(t) (res / dropout optional; t may alias z), y = LayerNorm(z) * g + b (+ a bf16 copy of y): the
// residual + dropout + LayerNorm tail of both halves of an encoder layer in one pass, one warp per row.
// d % 128 == 0, d <= 1024; rows are grouped S per sequence for the dropout counters.
__global__ void __launch_bounds__(256) add_drop_ln_kernel(const float* t, const float* __restrict__ res, float* z,
                                                          const float* __restrict__ g, const float* __restrict__ bt,
                                                          float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf, int M, int d,
                                                          int S, Drop drop, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const int per4 = d >> 7;  // float4 per lane
  const int seq = row / S;
  const long long local4 = (long long)(row - seq * S) * (d >> 2);
  const float4* t4 = reinterpret_cast<const float4*>(t + (long long)row * d);
  const float4* r4 = res ? reinterpret_cast<const float4*>(res + (long long)row * d) : nullptr;
  float4* z4 = z ? reinterpret_cast<float4*>(z + (long long)row * d) : nullptr;
  float4 v[8];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per4) {
      const int c4 = lane + 32 * i;
      float4 x = t4[c4];
      if (drop.on()) {
        const float4 m = drop_scale4(drop, site, seq, local4 + c4);
        x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
      }
      if (r4) {
        const float4 a = r4[c4];
        x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
      }
      if (z4) z4[c4] = x;
      v[i] = x;
      sum += (x.x + x.y) + (x.z + x.w);
    }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)d;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)d + 1e-5f);
  float4* y4 = reinterpret_cast<float4*>(y + (long long)row * d);
  uint2* yb = y_bf ? reinterpret_cast<uint2*>(y_bf + (long long)row * d) : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per4) {
      const int c4 = lane + 32 * i;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + c4), bb = __ldg(reinterpret_cast<const float4*>(bt) + c4);
      const float4 o = make_float4((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
                                   (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
      y4[c4] = o;
      if (yb) yb[c4] = bf16x4(o);
    }
}

// LayerNorm backward over rows of width d (d % 32 == 0, d <= 1024): dz = rstd * (g - mean(g) - xhat * mean(g xhat)),
// g = dy * gamma; dgamma += sum_rows dy * xhat, dbeta += sum_rows dy (block partials, then atomics).
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ gamma, float* __restrict__ dz,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int d) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int per = d / 32;
  float dg[32], db[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) dg[i] = db[i] = 0.0f;
  const float inv_d = 1.0f / (float)d;
  for (int row = blockIdx.x * 8 + wib; row < M; row += gridDim.x * 8) {
    const float* zr = z + (long long)row * d;
    const float* gr = dy + (long long)row * d;
    float v[32], g[32];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) { v[i] = zr[lane + 32 * i]; s += v[i]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) { const float cdev = v[i] - mean; q += cdev * cdev; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_d + 1e-5f);
    float sg = 0.0f, sgx = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) {
        const int n = lane + 32 * i;
        const float xh = (v[i] - mean) * rstd, go = gr[n];
        v[i] = xh;
        g[i] = go * gamma[n];
        sg += g[i];
        sgx += g[i] * xh;
        dg[i] += go * xh;
        db[i] += go;
      }
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg * inv_d, mgx = sgx * inv_d;
    float* dr = dz + (long long)row * d;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) dr[lane + 32 * i] = rstd * (g[i] - mg - v[i] * mgx);
  }
  // block reduction of the per-warp column partials, one 32-column slab at a time
  // (fully unrolled: a run-time index into dg / db would put both arrays in local memory for the whole kernel)
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i < per) {  // uniform over the block
      red[wib][lane] = dg[i];
      red[wib][32 + lane] = db[i];
      __syncthreads();
      if (wib == 0) {
        float a = 0.0f, b2 = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += red[w][lane]; b2 += red[w][32 + lane]; }
        if (dgamma) atomicAdd(dgamma + lane + 32 * i, a);
        if (dbeta) atomicAdd(dbeta + lane + 32 * i, b2);
      }
      __syncthreads();
    }
  }
}

// The same for d = 128 * NC with 128-bit accesses: a lane owns the float4 chunks lane + 32 c of the row, so the dropout mask of
// the residual branch (drawn per float4, drop_scale4) can be applied here and the separate dropout launch of the backward goes
// away (dzm = dz * mask of `site`; null = not wanted).  The column partials of the 8 warps meet ONCE in shared memory (the
// loop above pays 2 * NC * 4 barriers and as many rounds of atomics), and every warp fetches its next row before it reduces the
// current one.
template <int NC>
__global__ void __launch_bounds__(256) layernorm_bwd4_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                             const float* __restrict__ gamma, float* __restrict__ dz,
                                                             float* __restrict__ dzm, __nv_bfloat16* __restrict__ dz_bf,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int S,
                                                             Drop drop, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  constexpr int d = 128 * NC;
  __shared__ float4 red[8][2 * NC * 32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float4 dg[NC], db[NC], gm[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    dg[c] = db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[c] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * c);
  }
  const float inv_d = 1.0f / (float)d;
  const int stride = gridDim.x * 8;
  int row = blockIdx.x * 8 + wib;
  float4 v[NC], go[NC];
  if (row < M) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      v[c] = __ldg(reinterpret_cast<const float4*>(z + (long long)row * d) + lane + 32 * c);
      go[c] = __ldg(reinterpret_cast<const float4*>(dy + (long long)row * d) + lane + 32 * c);
    }
  }
  for (; row < M; row += stride) {
    float4 vn[NC], gn[NC];
    const int next = row + stride;
    if (next < M) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        vn[c] = __ldg(reinterpret_cast<const float4*>(z + (long long)next * d) + lane + 32 * c);
        gn[c] = __ldg(reinterpret_cast<const float4*>(dy + (long long)next * d) + lane + 32 * c);
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) sum += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
      q += v[c].x * v[c].x + v[c].y * v[c].y + v[c].z * v[c].z + v[c].w * v[c].w;
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_d + 1e-5f);
    float sg = 0.0f, sgx = 0.0f;
    float4 g[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      v[c].x *= rstd; v[c].y *= rstd; v[c].z *= rstd; v[c].w *= rstd;  // xhat
      g[c] = make_float4(go[c].x * gm[c].x, go[c].y * gm[c].y, go[c].z * gm[c].z, go[c].w * gm[c].w);
      sg += (g[c].x + g[c].y) + (g[c].z + g[c].w);
      sgx += g[c].x * v[c].x + g[c].y * v[c].y + g[c].z * v[c].z + g[c].w * v[c].w;
      dg[c].x += go[c].x * v[c].x; dg[c].y += go[c].y * v[c].y; dg[c].z += go[c].z * v[c].z; dg[c].w += go[c].w * v[c].w;
      db[c].x += go[c].x; db[c].y += go[c].y; db[c].z += go[c].z; db[c].w += go[c].w;
    }
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg * inv_d, mgx = sgx * inv_d;
    const int seq = row / S;
    const long long local4 = (long long)(row - seq * S) * (d / 4);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float4 o4 = make_float4(rstd * (g[c].x - mg - v[c].x * mgx), rstd * (g[c].y - mg - v[c].y * mgx),
                              rstd * (g[c].z - mg - v[c].z * mgx), rstd * (g[c].w - mg - v[c].w * mgx));
      reinterpret_cast<float4*>(dz + (long long)row * d)[lane + 32 * c] = o4;
      if (dzm) {
        const float4 m = drop_scale4(drop, site, seq, local4 + lane + 32 * c);
        o4.x *= m.x; o4.y *= m.y; o4.z *= m.z; o4.w *= m.w;
        reinterpret_cast<float4*>(dzm + (long long)row * d)[lane + 32 * c] = o4;
      }
      if (dz_bf) {  // bf16 copy of what the GEMM branch sees (masked when dropout is on): A operand of the dX GEMM
        const __nv_bfloat162 lo = __floats2bfloat162_rn(o4.x, o4.y), hi = __floats2bfloat162_rn(o4.z, o4.w);
        reinterpret_cast<uint2*>(dz_bf + (long long)row * d)[lane + 32 * c] =
            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) { v[c] = vn[c]; go[c] = gn[c]; }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    red[wib][c * 32 + lane] = dg[c];
    red[wib][(NC + c) * 32 + lane] = db[c];
  }
  __syncthreads();
  // 2 * d column sums, 4 per float4 slot: thread t owns slots t, t + 256, ...
  for (int slot = threadIdx.x; slot < 2 * NC * 32; slot += 256) {
    float4 a = red[0][slot];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 b = red[w][slot];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    const bool is_b = slot >= NC * 32;
    float* out = is_b ? dbeta : dgamma;
    if (out) {
      const int col = 4 * (slot - (is_b ? NC * 32 : 0));  // chunk (c * 32 + lane) covers columns 4 * (lane + 32 c) ..
      atomicAdd(out + col, a.x); atomicAdd(out + col + 1, a.y); atomicAdd(out + col + 2, a.z); atomicAdd(out + col + 3, a.w);
    }
  }
}

// LayerNorm backward (+ the dropout-masked copy of dz for the GEMM branch when dzm is given): picks the 128-bit kernel when
// d is a multiple of 128.  Returns whether dzm was written (otherwise the caller runs the separate dropout kernel).
static int layernorm_bwd(const float* dy, const float* z, const float* gamma, float* dz, float* dzm, __nv_bfloat16* dz_bf,
                         float* dgamma, float* dbeta, int M, int d, int S, const Drop& drop, uint32_t site, bool* masked,
                         bool* staged, cudaStream_t s) {
  *masked = false;
  *staged = false;
  const int nc = d % 128 == 0 ? d / 128 : 0;
  if (nc == 1 || nc == 2 || nc == 4) {  // 8 would need 64 KB of static shared memory for the partials
    // rows per warp: 1 while that still fills the machine, at most ~4 (the column partials cost 2 d atomics per CTA)
    int blocks = ceil_div(M, 8);
    const int cap = 2 * sm_count();
    if (blocks > cap) blocks = ceil_div(M, 8 * ceil_div(blocks, cap));
    float* m = drop.on() ? dzm : nullptr;
    cudaError_t err;
    if (nc == 1) err = launch_pdl(layernorm_bwd4_kernel<1>, dim3(blocks), dim3(256), 0, s, dy, z, gamma, dz, m, dz_bf, dgamma, dbeta, M, S, drop, site);
    else if (nc == 2) err = launch_pdl(layernorm_bwd4_kernel<2>, dim3(blocks), dim3(256), 0, s, dy, z, gamma, dz, m, dz_bf, dgamma, dbeta, M, S, drop, site);
    else err = launch_pdl(layernorm_bwd4_kernel<4>, dim3(blocks), dim3(256), 0, s, dy, z, gamma, dz, m, dz_bf, dgamma, dbeta, M, S, drop, site);
    MST_CUDA_OK(err);
    *masked = m != nullptr;
    *staged = dz_bf != nullptr;
  } else {
    const int blocks = ceil_div(M, 8) < sm_count() ? ceil_div(M, 8) : sm_count();  // few blocks: 2 d atomics each at the end
    MST_CUDA_OK(launch_pdl(layernorm_bwd_kernel, dim3(blocks), dim3(256), 0, s, dy, z, gamma, dz, dgamma, dbeta, M, d));
  }
  MST_LAUNCHED("bwd_ln", s);
  return MST_OK;
}

// out[n] += sum_m x[m*ld + n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N, int ld,
                                                     int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.0f;
  if (n < N)
    for (int m = r0 + ty; m < r1; m += 8) s += x[(long long)m * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][tx];
    atomicAdd(out + n, t);
  }
}

static int colsum(const float* x, float* out, int M, int N, int ld, cudaStream_t s) {
  int rows_per_block = 256;
  dim3 grid(ceil_div(N, 32), ceil_div(M, rows_per_block));
  MST_CUDA_OK(launch_pdl(colsum_kernel, grid, dim3(256), 0, s, x, out, M, N, ld, rows_per_block));
  MST_LAUNCHED("colsum", s);
  return MST_OK;
}

static int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

static int dropout(const float* in, const float* add, float* out, long long n, const Drop& d, uint32_t site, cudaStream_t s) {
  if (n % (4ll * d.n_seqs) != 0)
    return fail(MST_ERR_UNSUPPORTED, "dropout: per-sequence tensor sizes must be multiples of 4");
  MST_CUDA_OK(launch_pdl(dropout_kernel, dim3(ew_blocks(n / 4)), dim3(256), 0, s, in, add, out, n / 4, n / 4 / d.n_seqs, d, site));
  MST_LAUNCHED("dropout", s);
  return MST_OK;
}
static inline uint32_t drop_site(int layer, int which) { return (uint32_t)(8 * (layer + 1) + which); }

// MotionEncoder tokens: rows 0 / 1 of every sequence = muQuery / sigmaQuery + pe[row]; rows >= 2 (already holding
// InputProcess(x)) += pe[row]
__global__ void __launch_bounds__(128) menc_tokens_kernel(const float* __restrict__ q0, const float* __restrict__ q1,
                                                          const float* __restrict__ pe, float* __restrict__ x, int S, int d) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  const int seq = blockIdx.x / S, r = blockIdx.x - seq * S;
  float* xr = x + ((long long)seq * S + r) * d;
  const float* per = pe + (long long)r * d;
  for (int n = threadIdx.x; n < d; n += blockDim.x) {
    if (r == 0) xr[n] = q0[n] + per[n];
    else if (r == 1) xr[n] = q1[n] + per[n];
    else xr[n] += per[n];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Linear layers of the encoder stack: y = x W^T (+ bias) (+ add);  dx = dy W (+ add);  dW += dy^T x.
// MST_PREC_FP32 engines run them on the fp32 SIMT GEMM above (parity mode); MST_PREC_BF16 engines convert the
// operands to bf16 (K-major copies; dW needs both operands transposed, dX the transposed weight packed at load time)
// and run them on the tcgen05 kernel with fp32 accumulation and an fp32 epilogue (csrc/tc_gemm.cu, TC_EPI_TRAIN_F32).
// ---------------------------------------------------------------------------------------------------------
struct Stage {          // bf16 staging for the tensor-core path (null in fp32 mode)
  __nv_bfloat16* a;     // [M, max(3d, ff)]
  __nv_bfloat16* at;    // [max(3d, ff), Mpad]
  __nv_bfloat16* bt;    // [max(d, ff), Mpad]
  int m_pad;
};

struct Lin {
  const float* w;              // [n_out, n_in] fp32
  const __nv_bfloat16* w_bf;   // [n_out, n_in]
  const __nv_bfloat16* wt_bf;  // [n_in, n_out]
  int n_out, n_in;
};

static size_t stage_elems(const mst_model_desc& d, int M, int* m_pad) {
  const int mp = (M + 63) / 64 * 64;
  if (m_pad) *m_pad = mp;
  const size_t wide = (size_t)(3 * d.d_model > d.d_ff ? 3 * d.d_model : d.d_ff);
  const size_t mid = (size_t)(d.d_model > d.d_ff ? d.d_model : d.d_ff);
  return (size_t)M * wide + 64 + wide * mp + mid * mp;
}

static void carve_stage(const mst_model_desc& d, int M, void* base, Stage* st) {
  int mp;
  stage_elems(d, M, &mp);
  const size_t wide = (size_t)(3 * d.d_model > d.d_ff ? 3 * d.d_model : d.d_ff);
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(base);
  st->a = p;
  st->at = p + (((size_t)M * wide + 63) / 64) * 64;
  st->bt = st->at + wide * mp;
  st->m_pad = mp;
}

// a_staged: the producer of x already left its bf16 copy in st.a (fused conversions of the forward chain)
// Fused GELU of linear1 (tensor-core mode): h = dropout(gelu(out)) as fp32 (tape) and bf16 (operand of linear2)
struct GeluOut {
  float* h = nullptr;
  __nv_bfloat16* h_bf = nullptr;
  Drop drop;
  uint32_t site = 0;
  int rows_per_seq = 1;
};

// a_bf: the bf16 A operand when it is staged somewhere else than st.a
static int linear_fwd(bool tc, const Stage& st, const float* x, const Lin& L, const float* bias, const float* add, float* out,
                      int M, cudaStream_t s, const char* name, bool a_staged = false, const GeluOut* gelu = nullptr,
                      const __nv_bfloat16* a_bf = nullptr) {
  if (!tc) {
    GemmEx g;
    g.a = x; g.lda = L.n_in; g.b = L.w; g.ldb = L.n_in; g.trans_b = 1; g.bias = bias; g.add = add; g.c = out; g.ldc = L.n_out;
    g.M = M; g.N = L.n_out; g.K = L.n_in;
    return gemm_ex(g, s, name);
  }
  int rc;
  if (!a_staged && (rc = cvt_bf16(x, M, L.n_in, L.n_in, st.a, nullptr, 0, s))) return rc;
  TcGemmParams p;
  p.a = a_bf ? a_bf : st.a; p.w = L.w_bf; p.bias = bias; p.add = add; p.out = out; p.ldo = L.n_out; p.M = M; p.N = L.n_out; p.K = L.n_in;
  p.epi = TC_EPI_TRAIN_F32;
  if (gelu) {
    p.gelu_out = gelu->h; p.gelu_bf = gelu->h_bf; p.drop = gelu->drop; p.drop_site = gelu->site; p.rows_per_seq = gelu->rows_per_seq;
  }
  return tc_gemm(p, s);
}

// dx [M, n_in] = dy W (+ add); dW [n_out, n_in] += dy^T x; db [n_out] += column sums of dy (null pointers skip)
// dy_staged: the producer of dy already left its bf16 copy in st.a (then frozen layers - no dW, no db - need no conversion
// launch at all)
static int linear_bwd(bool tc, const Stage& st, const float* dy, const float* x, const Lin& L, const float* add, float* dx,
                      float* dw, float* db, int M, cudaStream_t s, const char* name_dx, const char* name_dw,
                      bool dy_staged = false) {
  int rc;
  if (db && !tc && (rc = colsum(dy, db, M, L.n_out, L.n_out, s))) return rc;
  if (!tc) {
    if (dw) {
      GemmEx g;  // dW += dy^T x
      g.a = dy; g.lda = L.n_out; g.a_mode = AX_TRANS; g.b = x; g.ldb = L.n_in; g.c = dw; g.ldc = L.n_in;
      g.M = L.n_out; g.N = L.n_in; g.K = M; g.accumulate = 1; g.split_k = pick_split(L.n_out, L.n_in, M);
      if ((rc = gemm_ex(g, s, name_dw))) return rc;
    }
    if (dx) {
      GemmEx g;  // dx = dy W (+ add)
      g.a = dy; g.lda = L.n_out; g.b = L.w; g.ldb = L.n_in; g.add = add; g.c = dx; g.ldc = L.n_in;
      g.M = M; g.N = L.n_in; g.K = L.n_out;
      if ((rc = gemm_ex(g, s, name_dx))) return rc;
    }
    return MST_OK;
  }
  {  // one launch: dy -> bf16 (operand of dX), dy^T (operand of dW), db += column sums;  x^T (operand of dW)
    CvtJobs js;
    __nv_bfloat16* a_dst = dx && !dy_staged ? st.a : nullptr;
    if (a_dst || dw || db)
      cvt_jobs_add(js, dy, M, L.n_out, L.n_out, a_dst, nullptr, M, L.n_out, dw ? st.at : nullptr, st.m_pad, db);
    if (dw) cvt_jobs_add(js, x, M, L.n_in, L.n_in, nullptr, nullptr, 0, 0, st.bt, st.m_pad, nullptr);
    if ((rc = cvt_multi(js, s, "bwd_operands"))) return rc;
  }
  if (dw) {
    TcGemmParams p;
    p.a = st.at; p.w = st.bt; p.out = dw; p.ldo = L.n_in; p.M = L.n_out; p.N = L.n_in; p.K = st.m_pad;
    p.accumulate = 1; p.epi = TC_EPI_TRAIN_F32;
    if ((rc = tc_gemm(p, s))) return rc;
  }
  if (dx) {
    TcGemmParams p;
    p.a = st.a; p.w = L.wt_bf; p.add = add; p.out = dx; p.ldo = L.n_in; p.M = M; p.N = L.n_in; p.K = L.n_out;
    p.epi = TC_EPI_TRAIN_F32;
    if ((rc = tc_gemm(p, s))) return rc;
  }
  return MST_OK;
}

static void layer_lins(Engine* e, int l, Lin* qkv, Lin* o, Lin* f1, Lin* f2) {
  const mst_model_desc& d = e->desc;
  const LayerF32& L = e->lf[l];
  const bool tc = d.precision == MST_PREC_BF16;
  *qkv = Lin{L.qkv_w, tc ? e->lb[l].qkv_w : nullptr, tc ? e->lbt[l].qkv_w : nullptr, 3 * d.d_model, d.d_model};
  *o = Lin{L.o_w, tc ? e->lb[l].o_w : nullptr, tc ? e->lbt[l].o_w : nullptr, d.d_model, d.d_model};
  *f1 = Lin{L.w1, tc ? e->lb[l].w1 : nullptr, tc ? e->lbt[l].w1 : nullptr, d.d_ff, d.d_model};
  *f2 = Lin{L.w2, tc ? e->lb[l].w2 : nullptr, tc ? e->lbt[l].w2 : nullptr, d.d_model, d.d_ff};
}

// ---------------------------------------------------------------------------------------------------------
// Fused attention of the training path for short sequences (S <= 80: the 76-frame stylexia clips, 77 / 78 tokens):
// one CTA per (sequence, head) keeps Q, K, V (and P, dO, dP in the backward) in shared memory and runs the whole
// chain - scores, masked softmax, dropout, P V / dV, dP, softmax backward, dQ, dK - without the six batched GEMM
// launches over 77 x 77 x 128 problems (each padded to 128-wide tiles) that the general path needs.
// fp32; 256 threads as a 16 x 16 grid, every thread a TM x TN register tile with rows / columns interleaved by 16 so
// that shared-memory reads are conflict-free with the odd leading dimensions used below.
// ---------------------------------------------------------------------------------------------------------
constexpr int SA_MAXS = 80;          // padded sequence length the tiles cover (5 x 16)
constexpr int SA_DH = 128;
constexpr int SA_LDX = SA_DH + 4;    // Q, K, V, dO rows: 16-byte aligned, conflict-free for the 128-bit loads of smem_gemm_*2
constexpr int SA_LDP = SA_MAXS + 4;  // P, dP rows

// rows [0, S) x 128 columns of N global matrices (leading dimension ld, 16-byte aligned rows) -> smem [SA_MAXS][SA_LDX],
// rows >= S zero.  256 threads; every thread issues its 10 float4 loads per matrix before the first shared-memory store, so
// the CTA pays the global latency once instead of once per element (the scalar loop was ~60 % of the kernel).
template <int N, int NT = 256>
__device__ __forceinline__ void sa_load(float* const (&dst)[N], const float* const (&src)[N], int ld, int S) {
  constexpr int PER = SA_MAXS * (SA_DH / 4) / NT;  // 10 (256 threads) / 5 (512)
  float4 v[N][PER];
#pragma unroll
  for (int m = 0; m < N; ++m)
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = threadIdx.x + NT * i, r = idx >> 5, c4 = idx & 31;
      v[m][i] = r < S ? __ldg(reinterpret_cast<const float4*>(src[m] + (long long)r * ld) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
  for (int m = 0; m < N; ++m)
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = threadIdx.x + NT * i, r = idx >> 5, c4 = idx & 31;
      *reinterpret_cast<float4*>(dst[m] + r * SA_LDX + 4 * c4) = v[m][i];
    }
}
// [S][S] global -> smem [SA_MAXS][SA_LDP], zero padded (25 independent loads per thread)
template <int NT = 256>
__device__ __forceinline__ void sa_load_p(float* dst, const float* __restrict__ src, int S) {
  constexpr int PER = (SA_MAXS * SA_MAXS + NT - 1) / NT;  // 25 (256 threads) / 13 (512)
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = threadIdx.x + NT * i, r = idx / SA_MAXS, c = idx - r * SA_MAXS;
    v[i] = (r < S && c < S) ? __ldg(src + r * S + c) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = threadIdx.x + NT * i, r = idx / SA_MAXS, c = idx - r * SA_MAXS;
    if (idx < SA_MAXS * SA_MAXS) dst[r * SA_LDP + c] = v[i];
  }
}

__device__ __forceinline__ float drop_scale1(const Drop& d, uint32_t site, int seq, long long local) {
  const float4 m = drop_scale4(d, site, seq, local >> 2);
  const int r = (int)(local & 3);
  return r == 0 ? m.x : (r == 1 ? m.y : (r == 2 ? m.z : m.w));
}

// grid (heads, n_seqs, 5 / TM).  qkv [n_seqs*S, 3d]; p / pd [n_seqs][H][S][S]; ao [n_seqs*S, d]
// TM = 5: one CTA per (sequence, head), all 80 query rows.  TM = 1: five CTAs per (sequence, head), 16 query rows each -
// the B=1 steps of the finetune loop have only 4 (sequence, head) pairs, and one CTA's serial chain (30 us) was the longest
// kernel of their forward.
// Shared memory: Q and V share one buffer (V is fetched once the scores are done), so two CTAs fit on an SM - the kernel is
// latency-bound with 8 warps per SM (2 per scheduler), and at B=64 all 256 CTAs are then co-resident instead of 2 waves.
template <int TM, bool MMA>
__global__ void __launch_bounds__(256, 2) attn_small_fwd_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ key_valid,
                                                             float* __restrict__ p_out, float* __restrict__ pd_out,
                                                             float* __restrict__ ao, __nv_bfloat16* __restrict__ ao_bf, int S,
                                                             int d_model, int H, float scale, Drop drop, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + SA_MAXS * SA_LDX;
  float* Vs = Qs;  // overlays Q
  float* Ps = Ks + SA_MAXS * SA_LDX;
  constexpr int ROWS = 16 * TM;                    // query rows of this CTA
  const int head = blockIdx.x, seq = blockIdx.y, r0 = ROWS * (int)blockIdx.z;
  const float* base = qkv + (long long)seq * S * 3 * d_model + head * SA_DH;
  if (TM == 5) {
    float* const dst[2] = {Qs, Ks};
    const float* const src[2] = {base, base + d_model};
    sa_load<2>(dst, src, 3 * d_model, S);
  } else {
    float4 q[ROWS * (SA_DH / 4) / 256];  // this CTA's query rows, in flight together with K
#pragma unroll
    for (int i = 0; i < ROWS * (SA_DH / 4) / 256; ++i) {
      const int idx = threadIdx.x + 256 * i, r = r0 + (idx >> 5), c4 = idx & 31;
      q[i] = r < S ? __ldg(reinterpret_cast<const float4*>(base + (long long)r * 3 * d_model) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* const dst[1] = {Ks};
    const float* const src[1] = {base + d_model};
    sa_load<1>(dst, src, 3 * d_model, S);
#pragma unroll
    for (int i = 0; i < ROWS * (SA_DH / 4) / 256; ++i) {
      const int idx = threadIdx.x + 256 * i;
      *reinterpret_cast<float4*>(Qs + (idx >> 5) * SA_LDX + 4 * (idx & 31)) = q[i];
    }
  }
  __syncthreads();
  if constexpr (MMA)
    mma_gemm_nt80<TM>(Qs, SA_LDX, Ks, SA_LDX, SA_DH, [&](int i, int j, float v0, float v1) {
      *reinterpret_cast<float2*>(Ps + i * SA_LDP + j) = make_float2(v0 * scale, v1 * scale);
    });
  else
    smem_gemm_nt2<TM, 5>(Qs, SA_LDX, Ks, SA_LDX, SA_DH, [&](int i, int j, float v) { Ps[i * SA_LDP + j] = v * scale; });
  __syncthreads();
  {
    float* const dst[1] = {Vs};
    const float* const src[1] = {base + 2 * d_model};
    sa_load<1>(dst, src, 3 * d_model, S);  // visible to the P V product after the barrier that follows the softmax
  }
  // masked softmax, one warp per row; then dropout.  P (before dropout) and Pd go to the tape.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* kv = key_valid ? key_valid + (long long)seq * S : nullptr;
  float* pg = p_out + ((long long)seq * H + head) * S * S;
  float* pdg = pd_out + ((long long)seq * H + head) * S * S;
  for (int li = warp; li < ROWS; li += 8) {
    float* r = Ps + li * SA_LDP;
    const int i = r0 + li;  // row of the sequence
    if (i >= S) {
      for (int j = lane; j < SA_MAXS; j += 32) r[j] = 0.0f;
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32)
      if (!kv || kv[j]) mx = fmaxf(mx, r[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int j = lane; j < SA_MAXS; j += 32) {
      const float e = (j < S && (!kv || kv[j])) ? expf(r[j] - mx) : 0.0f;
      r[j] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    if (!drop.on()) {
      for (int j = lane; j < S; j += 32) {
        const float pv = r[j] * inv;
        pg[i * S + j] = pv;
        r[j] = pv;
      }
    } else {
      // the mask is drawn per group of 4 consecutive elements of the flattened [H, S, S] probabilities: a lane takes one
      // group (one Philox call instead of one per element) and the up to 4 elements of this row that fall into it
      const long long base = ((long long)head * S + i) * S, g0 = base >> 2;
      const int ng = (int)(((base + S - 1) >> 2) - g0) + 1;  // <= 21 for S <= 80
      for (int gi = lane; gi < ng; gi += 32) {
        const float4 m = drop_scale4(drop, site, seq, g0 + gi);
        const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = (int)(4 * (g0 + gi) + e - base);
          if (j >= 0 && j < S) {
            const float pv = r[j] * inv;
            pg[i * S + j] = pv;
            pdg[i * S + j] = pv * mm[e];
            r[j] = pv * mm[e];
          }
        }
      }
    }
  }
  __syncthreads();
  float* og = ao + (long long)seq * S * d_model + head * SA_DH;
  __nv_bfloat16* ob = ao_bf ? ao_bf + (long long)seq * S * d_model + head * SA_DH : nullptr;  // operand of the out-projection
  if constexpr (MMA)
    mma_gemm_n128<false, TM>(Ps, SA_LDP, Vs, SA_LDX, SA_MAXS, [&](int li, int c, float v0, float v1) {
      const int i = r0 + li;
      if (i < S) {
        *reinterpret_cast<float2*>(og + (long long)i * d_model + c) = make_float2(v0, v1);
        if (ob) *reinterpret_cast<__nv_bfloat162*>(ob + (long long)i * d_model + c) = __floats2bfloat162_rn(v0, v1);
      }
    });
  else
    smem_gemm_n128<false, TM>(Ps, SA_LDP, Vs, SA_LDX, SA_MAXS, [&](int li, int c, float v) {
      const int i = r0 + li;
      if (i < S) {
        og[(long long)i * d_model + c] = v;
        if (ob) ob[(long long)i * d_model + c] = __float2bfloat16_rn(v);
      }
    });
}

// grid (heads, n_seqs).  dao [n_seqs*S, d] -> dqkv [n_seqs*S, 3d] (Q | K | V column blocks)
// The tensor-core variant runs 16 warps (the kernel is a chain of latency-bound phases on ONE CTA per SM: twice the warps
// hide twice the latency); the fp32 variant keeps the 16 x 16 thread grid its register-tiled products are written for.
template <bool MMA>
__global__ void __launch_bounds__(MMA ? 512 : 256) attn_small_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ p_in,
                                                             const float* __restrict__ pd_in, const float* __restrict__ dao,
                                                             float* __restrict__ dqkv, __nv_bfloat16* __restrict__ dqkv_bf, int S,
                                                             int d_model, int H, float scale, Drop drop, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  constexpr int NT = MMA ? 512 : 256, NW = NT / 32;
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + SA_MAXS * SA_LDX;
  float* Vs = Ks + SA_MAXS * SA_LDX;
  float* Os = Vs + SA_MAXS * SA_LDX;   // dO
  float* Ps = Os + SA_MAXS * SA_LDX;   // P after dropout, then P
  float* Ds = Ps + SA_MAXS * SA_LDP;   // dP, then dS
  const int head = blockIdx.x, seq = blockIdx.y;
  const float* base = qkv + (long long)seq * S * 3 * d_model + head * SA_DH;
  float* dbase = dqkv + (long long)seq * S * 3 * d_model + head * SA_DH;
  __nv_bfloat16* dbf = dqkv_bf ? dqkv_bf + (long long)seq * S * 3 * d_model + head * SA_DH : nullptr;  // operand of the dX GEMM
  const long long pp = ((long long)seq * H + head) * S * S;
  {
    float* const dst[3] = {Qs, Ks, Vs};
    const float* const src[3] = {base, base + d_model, base + 2 * d_model};
    sa_load<3, NT>(dst, src, 3 * d_model, S);
    float* const dst_o[1] = {Os};
    const float* const src_o[1] = {dao + (long long)seq * S * d_model + head * SA_DH};
    sa_load<1, NT>(dst_o, src_o, d_model, S);
  }
  sa_load_p<NT>(Ps, (drop.on() ? pd_in : p_in) + pp, S);
  __syncthreads();
  // dV = Pd^T dO
  // results of the 128-column products: column block `blk` (0 Q, 1 K, 2 V) of dqkv, scaled
  auto store1 = [&](int blk, float sc) {
    return [=](int r, int c, float v) {
      if (r < S) {
        dbase[(long long)r * 3 * d_model + blk * d_model + c] = v * sc;
        if (dbf) dbf[(long long)r * 3 * d_model + blk * d_model + c] = __float2bfloat16_rn(v * sc);
      }
    };
  };
  auto store2 = [&](int blk, float sc) {
    return [=](int r, int c, float v0, float v1) {
      if (r < S) {
        *reinterpret_cast<float2*>(dbase + (long long)r * 3 * d_model + blk * d_model + c) = make_float2(v0 * sc, v1 * sc);
        if (dbf)
          *reinterpret_cast<__nv_bfloat162*>(dbf + (long long)r * 3 * d_model + blk * d_model + c) = __floats2bfloat162_rn(v0 * sc, v1 * sc);
      }
    };
  };
  if constexpr (MMA) mma_gemm_n128<true, 5, NW>(Ps, SA_LDP, Os, SA_LDX, SA_MAXS, store2(2, 1.0f));
  else smem_gemm_n128<true, 5>(Ps, SA_LDP, Os, SA_LDX, SA_MAXS, store1(2, 1.0f));
  // dP = dO V^T (the forward's dropout mask is applied row by row below)
  if constexpr (MMA)
    mma_gemm_nt80<5, NW>(Os, SA_LDX, Vs, SA_LDX, SA_DH, [&](int i, int j, float v0, float v1) {
      *reinterpret_cast<float2*>(Ds + i * SA_LDP + j) = make_float2(v0, v1);
    });
  else
    smem_gemm_nt2<5, 5>(Os, SA_LDX, Vs, SA_LDX, SA_DH, [&](int i, int j, float v) { Ds[i * SA_LDP + j] = v; });
  __syncthreads();
  if (drop.on()) {  // the softmax backward needs the probabilities BEFORE dropout
    sa_load_p<NT>(Ps, p_in + pp, S);
    __syncthreads();
  }
  // dS = P * (dP - sum_j P dP)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < SA_MAXS; i += NW) {
    const float* pr = Ps + i * SA_LDP;
    float* dr = Ds + i * SA_LDP;
    if (drop.on() && i < S) {  // dP *= mask: one Philox call per group of 4 consecutive elements (see the forward)
      const long long base = ((long long)head * S + i) * S, g0 = base >> 2;
      const int ng = (int)(((base + S - 1) >> 2) - g0) + 1;
      for (int gi = lane; gi < ng; gi += 32) {
        const float4 m = drop_scale4(drop, site, seq, g0 + gi);
        const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = (int)(4 * (g0 + gi) + e - base);
          if (j >= 0 && j < S) dr[j] *= mm[e];
        }
      }
      __syncwarp();
    }
    float dot = 0.0f;
    for (int j = lane; j < SA_MAXS; j += 32) dot = fmaf(pr[j], dr[j], dot);
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int j = lane; j < SA_MAXS; j += 32) dr[j] = pr[j] * (dr[j] - dot);
  }
  __syncthreads();
  // dQ = scale dS K ;  dK = scale dS^T Q
  if constexpr (MMA) {
    mma_gemm_n128<false, 5, NW>(Ds, SA_LDP, Ks, SA_LDX, SA_MAXS, store2(0, scale));
    mma_gemm_n128<true, 5, NW>(Ds, SA_LDP, Qs, SA_LDX, SA_MAXS, store2(1, scale));
  } else {
    smem_gemm_n128<false, 5>(Ds, SA_LDP, Ks, SA_LDX, SA_MAXS, store1(0, scale));
    smem_gemm_n128<true, 5>(Ds, SA_LDP, Qs, SA_LDX, SA_MAXS, store1(1, scale));
  }
}

constexpr size_t SA_FWD_SMEM = (size_t)(2 * SA_MAXS * SA_LDX + SA_MAXS * SA_LDP) * sizeof(float);  // 111 KB: 2 CTAs per SM
constexpr size_t SA_BWD_SMEM = (size_t)(4 * SA_MAXS * SA_LDX + 2 * SA_MAXS * SA_LDP) * sizeof(float);
static_assert(SA_BWD_SMEM <= 227 * 1024, "small-attention backward shared memory");

// the 16-bit training mode runs the attention products on mma.sync (tf32); MST_TRAIN_ATTN_MMA=0 keeps the fp32 FMAs
static bool attn_mma(bool tc) {
  static const bool on = [] {
    const char* e = getenv("MST_TRAIN_ATTN_MMA");
    return !(e && e[0] == '0');
  }();
  return tc && on;
}

// MST_TRAIN_GELU_EPI=1: the GELU (+ dropout) of the taped forward inside linear1's epilogue instead of its own kernel.
// One node less per layer, but measured SLOWER (10.14 vs 9.82 ms and 4.22 vs 4.13 ms per finetune step with / without
// semantic guidance on the same box): the row-per-thread fp32 epilogue then runs 64 erff + a Philox call per thread and
// tile on the critical chain of every layer.  Off by default, kept for the record.
static bool gelu_in_epilogue() {
  static const bool on = [] {
    const char* e = getenv("MST_TRAIN_GELU_EPI");
    return e && e[0] == '1';
  }();
  return on;
}

static bool attn_small_ok(const mst_model_desc& d, int S) {
  return S <= SA_MAXS && d.d_model / d.n_heads == SA_DH;
}

// The in-/out-projections of the B=64 batch on the fp32 SIMT GEMM cost 110-250 us each (K or N = 181 keeps them off the TMA
// path): in the 16-bit mode the motion goes through motion_to_tokens_bf16 (the sampler's [B*T, f_pad] operand) and the
// tcgen05 training GEMM instead.  MST_TRAIN_PROJ_TC=0 keeps the fp32 kernels.
static bool proj_on_tc(const Engine* e, int B, int T) {
  static const bool on = [] {
    const char* v = getenv("MST_TRAIN_PROJ_TC");
    return !(v && v[0] == '0');
  }();
  return on && e->desc.precision == MST_PREC_BF16 && e->in_w_bf && e->out_w_bf && (long long)B * T >= 512;
}

// tokens[(b, t + tok_off)] = x[b,:,t] in_w^T + in_b (+ pe[t + tok_off]) into out [B * (T + tok_off), d]; scratch_bf: B*T*f_pad bf16
static int inproj_tc(Engine* e, const float* x, int B, int T, int tok_off, const float* pe, float* out,
                     __nv_bfloat16* scratch_bf, cudaStream_t s) {
  const mst_model_desc& d = e->desc;
  int rc;
  if ((rc = motion_to_tokens_bf16(x, scratch_bf, B, d.n_feats, T, e->f_pad, nullptr, 0, s))) return rc;
  TcGemmParams p;
  p.a = scratch_bf; p.w = e->in_w_bf; p.bias = e->in_b; p.out = out; p.ldo = d.d_model;
  p.M = B * T; p.N = d.d_model; p.K = e->f_pad; p.epi = TC_EPI_TRAIN_F32;
  p.tok_T = T; p.tok_off = tok_off; p.pe = pe;
  return tc_gemm(p, s);
}

// g[(b, s)] = s < tok_off ? (untouched) : sum_f d_out[b][f][s - tok_off] out_w[f][:]  - backward of the out-projection
static int outproj_bwd_tc(Engine* e, const float* d_out, int B, int T, float* g, __nv_bfloat16* scratch_bf, cudaStream_t s) {
  const mst_model_desc& d = e->desc;
  int rc;
  if ((rc = motion_to_tokens_bf16(d_out, scratch_bf, B, d.n_feats, T, e->f_pad, nullptr, 0, s))) return rc;
  TcGemmParams p;
  p.a = scratch_bf; p.w = e->out_wt_bf; p.out = g; p.ldo = d.d_model;
  p.M = B * T; p.N = d.d_model; p.K = e->f_pad; p.epi = TC_EPI_TRAIN_F32;
  p.tok_T = T; p.tok_off = 1;
  return tc_gemm(p, s);
}

// d_x[b][f][t] = sum_n gx[(b, t + tok_off)][n] in_w[n][f]  - backward of the in-projection; gx_bf: bf16 copy of gx [M, d]
static int inproj_bwd_tc(Engine* e, const __nv_bfloat16* gx_bf, int B, int T, int tok_off, float* d_x, cudaStream_t s) {
  const mst_model_desc& d = e->desc;
  TcGemmParams p;
  p.a = gx_bf; p.w = e->in_wt_bf; p.bias = e->zero_pad; p.out = d_x;
  p.M = B * (T + tok_off); p.N = e->f_pad; p.K = d.d_model; p.epi = TC_EPI_OUTPROJ_F32;
  p.B = B; p.T = T; p.n_valid = d.n_feats; p.drop_tokens = tok_off;
  return tc_gemm(p, s);
}

// ---------------------------------------------------------------------------------------------------------
// tape: everything the backward needs, per layer
// ---------------------------------------------------------------------------------------------------------
struct LayerTape {
  float *x, *qkv, *p, *pd, *ao, *z1, *y, *u, *h, *z2;  // pd: attention probabilities after dropout (= p when dropout is off)
};
struct Tape {
  LayerTape l[MST_MAX_LAYERS];
  float* x_out;  // output of the last layer [M, d]
  Stage stage;   // bf16 staging of the forward (tensor-core mode)
};

static size_t carve_tape(const mst_model_desc& d, int n_seqs, int S, void* base, Tape* t) {
  const size_t M = (size_t)n_seqs * S, dm = d.d_model, ff = d.d_ff;
  const size_t pp = (size_t)n_seqs * d.n_heads * S * S;
  size_t off = 0;
  auto take = [&](size_t n) {
    off = align_up(off, 256);
    float* r = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
    off += n * sizeof(float);
    return r;
  };
  Tape tp;
  float* x = take(M * dm);
  for (int l = 0; l < d.n_layers; ++l) {
    LayerTape& L = tp.l[l];
    L.x = x;
    L.qkv = take(M * 3 * dm);
    L.p = take(pp);
    L.pd = take(pp);
    L.ao = take(M * dm);
    L.z1 = take(M * dm);
    L.y = take(M * dm);
    L.u = take(M * ff);
    L.h = take(M * ff);
    L.z2 = take(M * dm);
    x = take(M * dm);
  }
  tp.x_out = x;
  tp.stage = Stage{nullptr, nullptr, nullptr, 0};
  if (d.precision == MST_PREC_BF16) {
    off = align_up(off, 1024);
    if (base) carve_stage(d, (int)M, static_cast<char*>(base) + off, &tp.stage);
    off += stage_elems(d, (int)M, nullptr) * sizeof(__nv_bfloat16);
  }
  if (t) *t = tp;
  return align_up(off, 256);
}

// The tape of `tape_seqs` sequences seen from sequence `k` on: a forward of fewer sequences records into its slice, and
// one backward over the whole tape later back-propagates all of them together.
static Tape tape_view(const Tape& full, const mst_model_desc& d, int S, int k) {
  Tape v = full;
  const size_t dm = d.d_model, ff = d.d_ff, row = (size_t)k * S, pp = (size_t)k * d.n_heads * S * S;
  for (int l = 0; l < d.n_layers; ++l) {
    LayerTape& L = v.l[l];
    L.x += row * dm; L.qkv += row * 3 * dm; L.p += pp; L.pd += pp; L.ao += row * dm; L.z1 += row * dm; L.y += row * dm;
    L.u += row * ff; L.h += row * ff; L.z2 += row * dm;
  }
  v.x_out += row * dm;
  return v;
}

// stage_last: leave the bf16 copy of the stack's output in stage.a as well (operand of the tensor-core out-projection)
static int encoder_forward_tape(Engine* e, const Tape& tp, int NS, int S, const uint8_t* key_valid, const Drop& drop,
                                cudaStream_t s, bool stage_last = false) {
  const mst_model_desc& d = e->desc;
  const int M = NS * S, dm = d.d_model, H = d.n_heads, dh = dm / H, ff = d.d_ff;
  const float scale = 1.0f / sqrtf((float)dh);
  int rc;
  const bool tc = d.precision == MST_PREC_BF16;
  for (int l = 0; l < d.n_layers; ++l) {
    const LayerF32& L = e->lf[l];
    const LayerTape& t = tp.l[l];
    float* x_next = l + 1 < d.n_layers ? tp.l[l + 1].x : tp.x_out;
    Lin lqkv, lo, lf1, lf2;
    layer_lins(e, l, &lqkv, &lo, &lf1, &lf2);
    // fused tail kernels (residual + dropout + LayerNorm, GELU + dropout) that also leave the bf16 operand of the next
    // tensor-core GEMM in the staging buffer: 8 kernels per layer instead of 16 on the latency-bound B=1 chain
    const bool fused = dm % 128 == 0 && ff % 4 == 0;
    __nv_bfloat16* bf = tc ? tp.stage.a : nullptr;
    const bool x_staged = tc && fused && l > 0;  // the previous layer's LayerNorm kernel staged it
    if ((rc = linear_fwd(tc, tp.stage, t.x, lqkv, L.qkv_b, nullptr, t.qkv, M, s, "train_qkv", x_staged))) return rc;
    if (attn_small_ok(d, S)) {
      static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
      if (attr_set.first()) {
        MST_CUDA_OK(cudaFuncSetAttribute(attn_small_fwd_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_FWD_SMEM));
        MST_CUDA_OK(cudaFuncSetAttribute(attn_small_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_FWD_SMEM));
        MST_CUDA_OK(cudaFuncSetAttribute(attn_small_fwd_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_FWD_SMEM));
        MST_CUDA_OK(cudaFuncSetAttribute(attn_small_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_FWD_SMEM));
      }
      // 16-bit mode: the two products on mma.sync (tf32); fp32 parity mode: exact fp32 FMAs
      const bool few = NS * H <= 32;  // a handful of (sequence, head) pairs: five CTAs each
      auto kern = attn_mma(tc) ? (few ? attn_small_fwd_kernel<1, true> : attn_small_fwd_kernel<5, true>)
                     : (few ? attn_small_fwd_kernel<1, false> : attn_small_fwd_kernel<5, false>);
      MST_CUDA_OK(launch_pdl(kern, dim3(H, NS, few ? 5 : 1), dim3(256), SA_FWD_SMEM, s, (const float*)t.qkv, key_valid,
                             t.p, t.pd, t.ao, bf, S, dm, H, scale, drop, drop_site(l, 1)));
      MST_LAUNCHED("train_attn_small", s);
    } else {
      GemmEx sc;  // scores = scale * Q K^T per (seq, head)
      sc.a = t.qkv; sc.lda = 3 * dm; sc.b = t.qkv + dm; sc.ldb = 3 * dm; sc.trans_b = 1; sc.c = t.p; sc.ldc = S;
      sc.M = S; sc.N = S; sc.K = dh; sc.alpha = scale; sc.batch = NS; sc.heads = H;
      sc.a_bs = (long long)S * 3 * dm; sc.a_hs = dh; sc.b_bs = sc.a_bs; sc.b_hs = dh;
      sc.c_bs = (long long)H * S * S; sc.c_hs = (long long)S * S;
      if ((rc = gemm_ex(sc, s, "train_scores"))) return rc;
      const long long rows = (long long)NS * H * S;
      MST_CUDA_OK(launch_pdl(softmax_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, t.p, key_valid, rows, S, H * S));
      MST_LAUNCHED("train_softmax", s);
      const float* probs = t.p;
      if (drop.on()) {
        if ((rc = dropout(t.p, nullptr, t.pd, rows * S, drop, drop_site(l, 1), s))) return rc;
        probs = t.pd;
      }
      GemmEx pv;  // ao = P V
      pv.a = probs; pv.lda = S; pv.b = t.qkv + 2 * dm; pv.ldb = 3 * dm; pv.c = t.ao; pv.ldc = dm;
      pv.M = S; pv.N = dh; pv.K = S; pv.batch = NS; pv.heads = H;
      pv.a_bs = sc.c_bs; pv.a_hs = sc.c_hs; pv.b_bs = sc.a_bs; pv.b_hs = dh; pv.c_bs = (long long)S * dm; pv.c_hs = dh;
      if ((rc = gemm_ex(pv, s, "train_pv"))) return rc;
    }
    const bool ao_staged = tc && attn_small_ok(d, S);
    if (fused) {
      const int ln_grid = ceil_div(M, 8);
      // z1 = x + dropout1(ao Wo^T + bo); y = LN1(z1)
      if ((rc = linear_fwd(tc, tp.stage, t.ao, lo, L.o_b, drop.on() ? nullptr : t.x, t.z1, M, s, "train_outproj", ao_staged)))
        return rc;
      MST_CUDA_OK(launch_pdl(add_drop_ln_kernel, dim3(ln_grid), dim3(256), 0, s, (const float*)t.z1,
                             drop.on() ? (const float*)t.x : (const float*)nullptr, drop.on() ? t.z1 : (float*)nullptr, L.ln1_g,
                             L.ln1_b, t.y, bf, M, dm, S, drop, drop_site(l, 2)));
      MST_LAUNCHED("train_add_drop_ln", s);
      // h = dropout(gelu(u)); z2 = y + dropout2(h W2^T + b2); x' = LN2(z2)
      const __nv_bfloat16* h_bf = nullptr;
      if (tc && gelu_in_epilogue()) {
        // the GELU rides in linear1's epilogue; its bf16 copy goes to the (forward-idle) transpose staging, because the
        // GEMM is still reading its own A operand from stage.a
        GeluOut go;
        go.h = t.h; go.h_bf = tp.stage.at; go.drop = drop; go.site = drop_site(l, 3); go.rows_per_seq = S;
        if ((rc = linear_fwd(tc, tp.stage, t.y, lf1, L.b1, nullptr, t.u, M, s, "train_ffn1", tc, &go))) return rc;
        h_bf = tp.stage.at;
      } else {
        if ((rc = linear_fwd(tc, tp.stage, t.y, lf1, L.b1, nullptr, t.u, M, s, "train_ffn1", tc))) return rc;
        const long long n4 = (long long)M * ff / 4;
        MST_CUDA_OK(launch_pdl(gelu_drop_fwd_kernel, dim3(ew_blocks(n4)), dim3(256), 0, s, (const float*)t.u, t.h, bf, n4,
                               (long long)S * ff / 4, drop, drop_site(l, 3)));
        MST_LAUNCHED("train_gelu_drop", s);
      }
      if ((rc = linear_fwd(tc, tp.stage, t.h, lf2, L.b2, drop.on() ? nullptr : t.y, t.z2, M, s, "train_ffn2", tc, nullptr, h_bf))) return rc;
      MST_CUDA_OK(launch_pdl(add_drop_ln_kernel, dim3(ln_grid), dim3(256), 0, s, (const float*)t.z2,
                             drop.on() ? (const float*)t.y : (const float*)nullptr, drop.on() ? t.z2 : (float*)nullptr, L.ln2_g,
                             L.ln2_b, x_next, (l + 1 < d.n_layers || stage_last) ? bf : (__nv_bfloat16*)nullptr, M, dm, S, drop,
                             drop_site(l, 4)));
      MST_LAUNCHED("train_add_drop_ln", s);
      continue;
    }
    if (drop.on()) {  // z1 = x + dropout1(ao Wo^T + bo)
      if ((rc = linear_fwd(tc, tp.stage, t.ao, lo, L.o_b, nullptr, t.z1, M, s, "train_outproj"))) return rc;
      if ((rc = dropout(t.z1, t.x, t.z1, (long long)M * dm, drop, drop_site(l, 2), s))) return rc;
    } else if ((rc = linear_fwd(tc, tp.stage, t.ao, lo, L.o_b, t.x, t.z1, M, s, "train_outproj"))) {
      return rc;
    }
    if ((rc = layernorm_f32(t.z1, L.ln1_g, L.ln1_b, t.y, M, dm, s))) return rc;
    if ((rc = linear_fwd(tc, tp.stage, t.y, lf1, L.b1, nullptr, t.u, M, s, "train_ffn1"))) return rc;
    MST_CUDA_OK(launch_pdl(gelu_fwd_kernel, dim3(ew_blocks((long long)M * ff)), dim3(256), 0, s, (const float*)t.u, t.h, (long long)M * ff));
    MST_LAUNCHED("train_gelu", s);
    if (drop.on()) {  // h = dropout(gelu(u)); z2 = y + dropout2(h W2^T + b2)
      if ((rc = dropout(t.h, nullptr, t.h, (long long)M * ff, drop, drop_site(l, 3), s))) return rc;
      if ((rc = linear_fwd(tc, tp.stage, t.h, lf2, L.b2, nullptr, t.z2, M, s, "train_ffn2"))) return rc;
      if ((rc = dropout(t.z2, t.y, t.z2, (long long)M * dm, drop, drop_site(l, 4), s))) return rc;
    } else if ((rc = linear_fwd(tc, tp.stage, t.h, lf2, L.b2, t.y, t.z2, M, s, "train_ffn2"))) {
      return rc;
    }
    if ((rc = layernorm_f32(t.z2, L.ln2_g, L.ln2_b, x_next, M, dm, s))) return rc;
  }
  return MST_OK;
}

// Backward of the encoder stack.  g_out [M, d]: gradient w.r.t. the stack's output, overwritten with scratch;
// on return `g_x` [M, d] holds the gradient w.r.t. the stack's input.  Parameter gradients are ACCUMULATED.
struct BwdScratch {
  float *ga, *gb, *gm, *dqkv, *dp, *dh;  // [M,d] x3 (gm: dropout-masked gradient), [M,3d], [NS*H*S*S], [M,ff]
  Stage stage;
};

static size_t carve_bwd(const mst_model_desc& d, int n_seqs, int S, void* base, BwdScratch* o) {
  const size_t M = (size_t)n_seqs * S, dm = d.d_model;
  size_t off = 0;
  auto take = [&](size_t n) {
    off = align_up(off, 256);
    float* r = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
    off += n * sizeof(float);
    return r;
  };
  BwdScratch b;
  b.ga = take(M * dm);
  b.gb = take(M * dm);
  b.gm = take(M * dm);
  b.dqkv = take(M * 3 * dm);
  b.dp = take((size_t)n_seqs * d.n_heads * S * S);
  b.dh = take(M * d.d_ff);
  b.stage = Stage{nullptr, nullptr, nullptr, 0};
  if (d.precision == MST_PREC_BF16) {
    off = align_up(off, 1024);
    if (base) carve_stage(d, (int)M, static_cast<char*>(base) + off, &b.stage);
    off += stage_elems(d, (int)M, nullptr) * sizeof(__nv_bfloat16);
  }
  if (o) *o = b;
  return align_up(off, 256);
}

static int encoder_backward(Engine* e, const Tape& tp, const mst_layer_grads* grads, int NS, int S, float* g_in_out,
                            const BwdScratch& w, float** g_x, const Drop& drop, cudaStream_t s) {
  const mst_model_desc& d = e->desc;
  const int M = NS * S, dm = d.d_model, H = d.n_heads, dh = dm / H, ff = d.d_ff;
  const float scale = 1.0f / sqrtf((float)dh);
  int rc;
  float* gx = g_in_out;  // gradient w.r.t. the current layer's output
  float* spare = w.ga;
  float* spare2 = w.gb;
  const bool tc = d.precision == MST_PREC_BF16;
  for (int l = d.n_layers - 1; l >= 0; --l) {
    const LayerF32& L = e->lf[l];
    const LayerTape& t = tp.l[l];
    const mst_layer_grads& G = grads[l];
    Lin lqkv, lo, lf1, lf2;
    layer_lins(e, l, &lqkv, &lo, &lf1, &lf2);
    // LN2 (+ the dropout2 mask for the GEMM branch in the same kernel when the row width allows)
    float* dz2 = spare;
    bool masked, staged;
    __nv_bfloat16* bf = tc ? w.stage.a : nullptr;  // producers leave the bf16 operand of the next dX GEMM here
    if ((rc = layernorm_bwd(gx, t.z2, L.ln2_g, dz2, w.gm, bf, G.ln2_g, G.ln2_b, M, dm, S, drop, drop_site(l, 4), &masked, &staged, s))) return rc;
    // FFN2: dh = dz2 W2, dW2 += dz2^T h, db2 += sum dz2   (dropout2: the GEMM path sees the masked gradient)
    const float* dz2g = dz2;
    if (drop.on()) {
      if (!masked && (rc = dropout(dz2, nullptr, w.gm, (long long)M * dm, drop, drop_site(l, 4), s))) return rc;
      dz2g = w.gm;
    }
    if ((rc = linear_bwd(tc, w.stage, dz2g, t.h, lf2, nullptr, w.dh, G.w2, G.b2, M, s, "bwd_dh", "bwd_dw2", staged && (masked || !drop.on())))) return rc;
    if (drop.on())
      MST_CUDA_OK(launch_pdl(gelu_bwd_drop_kernel, dim3(ew_blocks((long long)M * ff / 4)), dim3(256), 0, s, (const float*)t.u, w.dh, bf,
                             (long long)M * ff / 4, (long long)S * ff / 4, drop, drop_site(l, 3)));
    else
      MST_CUDA_OK(launch_pdl(gelu_bwd_kernel, dim3(ew_blocks((long long)M * ff)), dim3(256), 0, s, (const float*)t.u, w.dh, bf, (long long)M * ff));
    MST_LAUNCHED("bwd_gelu", s);
    // FFN1: dy = dz2 + du W1 (into gx: the incoming gradient is no longer needed), dW1 += du^T y, db1 += sum du
    if ((rc = linear_bwd(tc, w.stage, w.dh, t.y, lf1, dz2, gx, G.w1, G.b1, M, s, "bwd_dy", "bwd_dw1", bf != nullptr))) return rc;
    // LN1
    float* dz1 = spare;  // dz2 is dead
    if ((rc = layernorm_bwd(gx, t.z1, L.ln1_g, dz1, w.gm, bf, G.ln1_g, G.ln1_b, M, dm, S, drop, drop_site(l, 2), &masked, &staged, s))) return rc;
    float* dao = spare2;  // out-proj: dao = dz1 Wo, dWo += dz1^T ao, dbo += sum dz1   (dropout1: masked gradient)
    const float* dz1g = dz1;
    if (drop.on()) {
      if (!masked && (rc = dropout(dz1, nullptr, w.gm, (long long)M * dm, drop, drop_site(l, 2), s))) return rc;
      dz1g = w.gm;
    }
    if ((rc = linear_bwd(tc, w.stage, dz1g, t.ao, lo, nullptr, dao, G.o_w, G.o_b, M, s, "bwd_dao", "bwd_dwo", staged && (masked || !drop.on())))) return rc;
    bool dqkv_staged = false;
    // attention
    if (attn_small_ok(d, S)) {
      static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
      if (attr_set.first()) {
        MST_CUDA_OK(cudaFuncSetAttribute(attn_small_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_BWD_SMEM));
        MST_CUDA_OK(cudaFuncSetAttribute(attn_small_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_BWD_SMEM));
      }
      MST_CUDA_OK(launch_pdl(attn_mma(tc) ? attn_small_bwd_kernel<true> : attn_small_bwd_kernel<false>, dim3(H, NS), dim3(attn_mma(tc) ? 512 : 256), SA_BWD_SMEM, s, (const float*)t.qkv, (const float*)t.p,
                             (const float*)t.pd, (const float*)dao, w.dqkv, bf, S, dm, H, scale, drop, drop_site(l, 1)));
      MST_LAUNCHED("bwd_attn_small", s);
      dqkv_staged = bf != nullptr;
    } else {
      const long long qkv_bs = (long long)S * 3 * dm, pp_bs = (long long)H * S * S, pp_hs = (long long)S * S;
      GemmEx dp;  // dP = dao V^T
      dp.a = dao; dp.lda = dm; dp.b = t.qkv + 2 * dm; dp.ldb = 3 * dm; dp.trans_b = 1; dp.c = w.dp; dp.ldc = S;
      dp.M = S; dp.N = S; dp.K = dh; dp.batch = NS; dp.heads = H;
      dp.a_bs = (long long)S * dm; dp.a_hs = dh; dp.b_bs = qkv_bs; dp.b_hs = dh; dp.c_bs = pp_bs; dp.c_hs = pp_hs;
      if ((rc = gemm_ex(dp, s, "bwd_dp"))) return rc;
      GemmEx dv;  // dV = P^T dao   (the probabilities the forward multiplied with: after dropout)
      dv.a = drop.on() ? t.pd : t.p; dv.lda = S; dv.a_mode = AX_TRANS; dv.b = dao; dv.ldb = dm; dv.c = w.dqkv + 2 * dm; dv.ldc = 3 * dm;
      dv.M = S; dv.N = dh; dv.K = S; dv.batch = NS; dv.heads = H;
      dv.a_bs = pp_bs; dv.a_hs = pp_hs; dv.b_bs = (long long)S * dm; dv.b_hs = dh; dv.c_bs = qkv_bs; dv.c_hs = dh;
      if ((rc = gemm_ex(dv, s, "bwd_dv"))) return rc;
      const long long rows = (long long)NS * H * S;
      if (drop.on() && (rc = dropout(w.dp, nullptr, w.dp, rows * S, drop, drop_site(l, 1), s))) return rc;
      MST_CUDA_OK(launch_pdl(softmax_bwd_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, (const float*)t.p, w.dp, rows, S));
      MST_LAUNCHED("bwd_softmax", s);
      GemmEx dq;  // dQ = scale dS K
      dq.a = w.dp; dq.lda = S; dq.b = t.qkv + dm; dq.ldb = 3 * dm; dq.c = w.dqkv; dq.ldc = 3 * dm; dq.alpha = scale;
      dq.M = S; dq.N = dh; dq.K = S; dq.batch = NS; dq.heads = H;
      dq.a_bs = pp_bs; dq.a_hs = pp_hs; dq.b_bs = qkv_bs; dq.b_hs = dh; dq.c_bs = qkv_bs; dq.c_hs = dh;
      if ((rc = gemm_ex(dq, s, "bwd_dq"))) return rc;
      GemmEx dk = dq;  // dK = scale dS^T Q
      dk.a_mode = AX_TRANS; dk.b = t.qkv; dk.c = w.dqkv + dm;
      if ((rc = gemm_ex(dk, s, "bwd_dk"))) return rc;
    }
    // QKV: gradient w.r.t. the layer input = dz1 + dqkv Wqkv, dWqkv += dqkv^T x, dbqkv += sum dqkv
    if ((rc = linear_bwd(tc, w.stage, w.dqkv, t.x, lqkv, dz1, gx, G.qkv_w, G.qkv_b, M, s, "bwd_dx", "bwd_dwqkv", dqkv_staged))) return rc;
  }
  *g_x = gx;
  return MST_OK;
}

// ---------------------------------------------------------------------------------------------------------
// masked L2 (gaussian_diffusion.py:223-235) and the gradient of the per-step update
// ---------------------------------------------------------------------------------------------------------
// loss[r] = sum_{f,t} (a - b)^2 mask[t] / (sum_t mask[t] * F);  rows r of a,b: [R, F, T]; a / mask rows are taken
// modulo a_rows / mask_rows (the reference expands a [1,...] target and mask over the stacked steps)
__global__ void __launch_bounds__(256) masked_l2_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask, float* __restrict__ loss,
                                                            int F, int T, int a_rows, int mask_rows) {
  __shared__ float red[2][8];
  const int r = blockIdx.x;
  const float* ar = a + (long long)(r % a_rows) * F * T;
  const float* br = b + (long long)r * F * T;
  const float* mr = mask + (long long)(r % mask_rows) * T;
  float s = 0.0f, ms = 0.0f;
  for (int i = threadIdx.x; i < F * T; i += blockDim.x) {
    const float dlt = ar[i] - br[i];
    s += dlt * dlt * mr[i % T];
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) ms += mr[t];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ms += __shfl_xor_sync(0xffffffffu, ms, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ms; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.0f, tm = 0.0f;
    for (int w = 0; w < 8; ++w) { ts += red[0][w]; tm += red[1][w]; }
    loss[r] = ts / (tm * (float)F);
  }
}

// db[r,f,t] = gl[r] * 2 (b - a) mask[t] / (sum_t mask * F)
__global__ void __launch_bounds__(256) masked_l2_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask, const float* __restrict__ gl,
                                                            float* __restrict__ db, int F, int T, int a_rows, int mask_rows) {
  __shared__ float red[8];
  __shared__ float coef;
  const int r = blockIdx.x;
  const float* ar = a + (long long)(r % a_rows) * F * T;
  const float* br = b + (long long)r * F * T;
  const float* mr = mask + (long long)(r % mask_rows) * T;
  float ms = 0.0f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) ms += mr[t];
  for (int o = 16; o > 0; o >>= 1) ms += __shfl_xor_sync(0xffffffffu, ms, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ms;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tm = 0.0f;
    for (int w = 0; w < 8; ++w) tm += red[w];
    coef = 2.0f * gl[r] / (tm * (float)F);
  }
  __syncthreads();
  float* dr = db + (long long)r * F * T;
  for (int i = threadIdx.x; i < F * T; i += blockDim.x) dr[i] = coef * (br[i] - ar[i]) * mr[i % T];
}

// d out = (d x0 + k[t_b] d sample) (1 - mask) [|x0| < 1 when clipped]
__global__ void __launch_bounds__(256) update_bwd_kernel(const float* __restrict__ d_x0, const float* __restrict__ d_sample,
                                                         const float* __restrict__ k_table, const int64_t* __restrict__ t_vec,
                                                         const float* __restrict__ mask, int mask_kind,
                                                         const float* __restrict__ x0, int clip, float* __restrict__ d_out,
                                                         int B, int F, int T) {
  const long long per = (long long)F * T, n = (long long)B * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const long long r = i - (long long)b * per;
    float g = d_x0 ? d_x0[i] : 0.0f;
    if (d_sample) g = fmaf(k_table[t_vec[b]], d_sample[i], g);
    float m = 0.0f;
    if (mask_kind == MST_MASK_FULL) m = mask[i];
    else if (mask_kind == MST_MASK_FT) m = mask[r];
    else if (mask_kind == MST_MASK_F) m = mask[r / T];
    g *= 1.0f - m;
    if (clip && x0 && fabsf(x0[i]) >= 1.0f && m == 0.0f) g = 0.0f;
    d_out[i] = g;
  }
}

// ---------------------------------------------------------------------------------------------------------
// optimizer: flat fp32 arenas (parameters, gradients, two moments)
// ---------------------------------------------------------------------------------------------------------
// torch.optim.AdamW semantics (decoupled decay, bias correction, eps added after the sqrt):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i];
    pi -= lr * wd * pi;
    const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// out[0] += sum x^2, out[1] += sum y^2 (y may be NULL): the grad / param norms of _compute_norms in one pass
__global__ void __launch_bounds__(256) sumsq2_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n,
                                                     double* __restrict__ out) {
  __shared__ double red[2][8];
  double sx = 0.0, sy = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = x[i];
    sx += (double)a * a;
    if (y) { const float b = y[i]; sy += (double)b * b; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sx; red[1][threadIdx.x >> 5] = sy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(out, a);
    if (y) atomicAdd(out + 1, b);
  }
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-graph replay of a whole taped forward / backward.  The B=1 style example of the finetune step is ~150 (forward) and
// ~300 (backward) dependent launches of a few microseconds each: launch-bound.  When the caller promises stable
// pointers (use_graph), the launch sequence is captured once per distinct argument set - on an internal stream, since
// torch's current stream is usually the legacy default stream, which cannot be captured - and replayed with one
// cudaGraphLaunch on the caller's stream.  Any capture problem falls back to the eager sequence.
// ---------------------------------------------------------------------------------------------------------
static uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
  return h;
}

static bool train_graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("MST_TRAIN_GRAPH");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename Body>
static int run_graphed(bool use_graph, uint64_t key, cudaStream_t user, Body&& body) {
  if (!use_graph || !train_graphs_enabled()) return body(user);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(user, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return body(user);  // already inside somebody else's capture: just record the launches there
  }
  static std::mutex mu;
  static std::unordered_map<uint64_t, cudaGraphExec_t> cache;
  static cudaStream_t cap_streams[64] = {};  // one capture stream per device (a stream belongs to the device it was made on)
  int dev_ = 0;
  cudaGetDevice(&dev_);
  cudaStream_t& cap_stream = cap_streams[dev_ & 63];
  key ^= 0x9e3779b97f4a7c15ull * (uint64_t)(dev_ + 1);
  cudaGraphExec_t exec = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) exec = it->second;
  }
  if (!exec) {
    std::lock_guard<std::mutex> lk(mu);
    if (!cap_stream && cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      return body(user);
    }
    if (cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return body(user);
    }
    const int rc = body(cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e1 = cudaStreamEndCapture(cap_stream, &graph);
    if (rc != MST_OK || e1 != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return body(user);  // capture failed (or the body reported an error while capturing): run - and report - eagerly
    }
    const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess || !exec) {
      cudaGetLastError();
      return body(user);
    }
    if (cache.size() >= 256) {  // bound the cache
      for (auto& kv : cache) cudaGraphExecDestroy(kv.second);
      cache.clear();
    }
    cache[key] = exec;
  }
  MST_CUDA_OK(cudaGraphLaunch(exec, user));
  note_launch("train_graph", user);
  return MST_OK;
}

}  // namespace mst

using namespace mst;

// ---------------------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------------------
static int check_train_engine(Engine* e) {
  if (!e->weights_loaded) return fail(MST_ERR_INVALID, "weights not loaded");
  if (e->desc.d_model % 32 != 0 || e->desc.d_model > 1024) return fail(MST_ERR_UNSUPPORTED, "d_model must be a multiple of 32, <= 1024");
  return MST_OK;
}

extern "C" int mst_abi_sizes_train(size_t* layer_grads, size_t* backward_args) {
  if (layer_grads) *layer_grads = sizeof(mst_layer_grads);
  if (backward_args) *backward_args = sizeof(mst_backward_args);
  return MST_OK;
}

extern "C" int mst_train_sizes(mst_engine_t h, int32_t n_seqs, int32_t seq_len, size_t* tape_bytes, size_t* scratch_bytes) {
  MST_CHECK_ARG(h != nullptr, "null engine");
  MST_CHECK_ARG(n_seqs > 0 && seq_len > 0, "non-positive size");
  Engine* e = reinterpret_cast<Engine*>(h);
  if (tape_bytes) *tape_bytes = carve_tape(e->desc, n_seqs, seq_len, nullptr, nullptr);
  if (scratch_bytes) *scratch_bytes = carve_bwd(e->desc, n_seqs, seq_len, nullptr, nullptr);
  return MST_OK;
}

extern "C" int mst_denoiser_forward_train(mst_engine_t h, const mst_forward_args* ap, void* tape, size_t tape_bytes,
                                          void* stream) {
  MST_CHECK_ARG(h && ap && tape, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  int rc;
  if ((rc = check_train_engine(e))) return rc;
  const mst_forward_args& a = *ap;
  MST_CHECK_ARG(a.batch > 0 && a.n_frames > 0, "empty batch");
  MST_CHECK_ARG(!a.cfg, "the training forward runs one pass (the reference trains the unwrapped model)");
  MST_CHECK_ARG(a.x && a.temb && a.out_cond, "null tensor pointer");
  MST_CHECK_ARG(a.n_frames + 1 <= e->desc.pe_len, "sequence longer than the positional table");
  const mst_model_desc& d = e->desc;
  const int B = a.batch, T = a.n_frames, S = T + 1, dm = d.d_model;
  const int tape_seqs = a.tape_seqs > 0 ? a.tape_seqs : B;
  MST_CHECK_ARG(a.tape_seq_offset >= 0 && a.tape_seq_offset + B <= tape_seqs, "tape_seq_offset + batch exceeds tape_seqs");
  Tape tp_full;
  MST_CHECK_ARG(tape_bytes >= carve_tape(d, tape_seqs, S, tape, &tp_full), "tape too small");
  const Tape tp = tape_view(tp_full, d, S, a.tape_seq_offset);
  MST_CHECK_ARG(a.dropout_p >= 0.0f && a.dropout_p < 1.0f && (a.dropout_p == 0.0f || a.dropout_seed), "bad dropout arguments");
  Drop drop;
  drop.p = a.dropout_p;
  drop.seed = reinterpret_cast<const unsigned long long*>(a.dropout_seed);
  drop.n_seqs = B;
  uint64_t key = fnv1a(&a, sizeof(a), 1469598103934665603ull);
  key = fnv1a(e, sizeof(Engine), key);
  key = fnv1a(&tape, sizeof(tape), key) ^ 0x66ull;
  return run_graphed(a.use_graph != 0, key, (cudaStream_t)stream, [&](cudaStream_t s) -> int {
    int rc;
    Token0Params t0;
    t0.temb = a.temb; t0.temb_row_dev = a.temb_row_dev; t0.temb_row_offset = a.temb_row_offset;
    t0.text_emb = a.text_emb; t0.txt_b = e->txt_b; t0.pe = e->pe; t0.x_f32 = tp.l[0].x;
    t0.B = B; t0.T = T; t0.d = dm; t0.cfg = 0; t0.uncond = a.uncond;
    if ((rc = token0(t0, B, s))) return rc;
    const bool ptc = proj_on_tc(e, B, T);
    if (ptc) {
      if ((rc = inproj_tc(e, a.x, B, T, 1, e->pe, tp.l[0].x, tp.stage.at, s))) return rc;
    } else {
      GemmF32Params p;
      p.a = a.x; p.a_mode = A_MOTION; p.w = e->in_w; p.ldw = d.n_feats; p.bias = e->in_b;
      p.c = tp.l[0].x; p.ldc = dm; p.M = B * T; p.N = dm; p.K = d.n_feats; p.epi = EPI_INPROJ;
      p.pe = e->pe; p.B = B; p.T = T; p.n_pass = 1;
      if ((rc = gemm_f32(p, s))) return rc;
    }
    if (drop.on() && (rc = dropout(tp.l[0].x, nullptr, tp.l[0].x, (long long)B * S * dm, drop, 0, s))) return rc;
    if ((rc = encoder_forward_tape(e, tp, B, S, nullptr, drop, s, ptc))) return rc;
    if (ptc) {  // the sampler's out-projection: bf16 tokens (staged by the last LayerNorm kernel) -> fp32 [B, F, T]
      TcGemmParams p;
      p.a = tp.stage.a; p.w = e->out_w_bf; p.bias = e->out_b_pad; p.out = a.out_cond;
      p.M = B * S; p.N = e->f_pad; p.K = dm; p.epi = TC_EPI_OUTPROJ_F32; p.B = B; p.T = T; p.n_valid = d.n_feats;
      if ((rc = tc_gemm(p, s))) return rc;
    } else {
      GemmF32Params p;
      p.a = tp.x_out; p.lda = dm; p.w = e->out_w; p.ldw = dm; p.bias = e->out_b;
      p.c = a.out_cond; p.M = B * S; p.N = d.n_feats; p.K = dm; p.epi = EPI_OUTPROJ; p.T = T; p.B = B;
      if ((rc = gemm_f32(p, s))) return rc;
    }
    return MST_OK;
  });
}

extern "C" int mst_denoiser_backward(mst_engine_t h, const mst_backward_args* ap, void* stream) {
  MST_CHECK_ARG(h && ap, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  int rc;
  if ((rc = check_train_engine(e))) return rc;
  const mst_backward_args& a = *ap;
  MST_CHECK_ARG(a.batch > 0 && a.n_frames > 0, "empty batch");
  MST_CHECK_ARG(a.d_out && a.tape && a.scratch && a.layer_grads, "null tensor pointer");
  const mst_model_desc& d = e->desc;
  const int B = a.batch, T = a.n_frames, S = T + 1, M = B * S, dm = d.d_model;
  const int tape_seqs = a.tape_seqs > 0 ? a.tape_seqs : B;
  MST_CHECK_ARG(a.tape_seq_offset >= 0 && a.tape_seq_offset + B <= tape_seqs, "tape_seq_offset + batch exceeds tape_seqs");
  Tape tp_full;
  BwdScratch w;
  MST_CHECK_ARG(a.tape_bytes >= carve_tape(d, tape_seqs, S, const_cast<void*>(a.tape), &tp_full), "tape too small");
  const Tape tp = tape_view(tp_full, d, S, a.tape_seq_offset);
  MST_CHECK_ARG(a.scratch_bytes >= carve_bwd(d, B, S, a.scratch, &w), "scratch too small");
  MST_CHECK_ARG(a.dropout_p >= 0.0f && a.dropout_p < 1.0f && (a.dropout_p == 0.0f || a.dropout_seed), "bad dropout arguments");
  Drop drop;
  drop.p = a.dropout_p;
  drop.seed = reinterpret_cast<const unsigned long long*>(a.dropout_seed);
  drop.n_seqs = B;
  // key: every device pointer / size of the call (not the host address of the layer_grads array, its contents)
  const void* kp[] = {a.d_out, a.d_x, a.tape, a.scratch, a.dropout_seed};
  uint32_t pbits;
  memcpy(&pbits, &a.dropout_p, 4);
  const size_t kn[] = {(size_t)a.batch, (size_t)a.n_frames, a.tape_bytes, a.scratch_bytes, (size_t)pbits,
                       (size_t)a.tape_seqs, (size_t)a.tape_seq_offset};
  uint64_t key = fnv1a(kp, sizeof(kp), 1469598103934665603ull);
  key = fnv1a(kn, sizeof(kn), key);
  key = fnv1a(e, sizeof(Engine), key);
  key = fnv1a(a.layer_grads, sizeof(mst_layer_grads) * d.n_layers, key) ^ 0xb7ull;
  return run_graphed(a.use_graph != 0, key, (cudaStream_t)stream, [&](cudaStream_t s) -> int {
    int rc;
    // OutputProcess backward: g[(b,s)][k] = s == 0 ? 0 : sum_f d_out[b][f][s-1] out_w[f][k]   (the tape's x_out slot is
    // dead after the forward: reuse it)
    float* g = tp.x_out;
    if (proj_on_tc(e, B, T) && e->out_wt_bf) {
      // token 0 of every sequence gets no gradient from the output projection
      MST_CUDA_OK(cudaMemset2DAsync(g, (size_t)S * dm * 4, 0, (size_t)dm * 4, B, s));
      if ((rc = outproj_bwd_tc(e, a.d_out, B, T, g, w.stage.at, s))) return rc;
    } else {
      GemmEx go;
      go.a = a.d_out; go.a_mode = AX_MOTION_TOK; go.T = T; go.tok_off = 1; go.b = e->out_w; go.ldb = dm; go.c = g; go.ldc = dm;
      go.M = M; go.N = dm; go.K = d.n_feats;
      if ((rc = gemm_ex(go, s, "bwd_outproj"))) return rc;
    }
    float* gx = nullptr;
    if ((rc = encoder_backward(e, tp, a.layer_grads, B, S, g, w, &gx, drop, s))) return rc;
    if (a.d_x) {  // InputProcess backward: d_x[b][f][t] = sum_n gx[(b,t+1)][n] in_w[n][f]
      if (drop.on() && (rc = dropout(gx, nullptr, gx, (long long)M * dm, drop, 0, s))) return rc;
      GemmEx gi;
      gi.a = gx; gi.lda = dm; gi.a_mode = AX_TOKROWS; gi.T = T; gi.tok_off = 1; gi.b = e->in_w; gi.ldb = d.n_feats;
      gi.c = a.d_x; gi.c_mode = CX_MOTION; gi.M = B * T; gi.N = d.n_feats; gi.K = dm;
      if ((rc = gemm_ex(gi, s, "bwd_inproj"))) return rc;
    }
    return MST_OK;
  });
}

// MotionEncoder (mdm_forstyledataset.py:89-124): tokens = [muQuery, sigmaQuery, InputProcess(x)] + pe, key-padding
// mask over the keys, the encoder stack of THIS engine; mu = output row 0 of every sequence.  The engine's in_w /
// in_b / pe are the frozen mdm_model's (the Python layer loads them so).
extern "C" int mst_motion_encoder_forward(mst_engine_t h, const float* x, const uint8_t* key_valid, const float* mu_query,
                                          const float* sigma_query, int32_t batch, int32_t n_frames, float* mu_out,
                                          void* tape, size_t tape_bytes, float dropout_p, const uint64_t* dropout_seed,
                                          int32_t use_graph, void* stream) {
  MST_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f && (dropout_p == 0.0f || dropout_seed), "bad dropout arguments");
  Drop drop;
  drop.p = dropout_p;
  drop.seed = reinterpret_cast<const unsigned long long*>(dropout_seed);
  drop.n_seqs = batch;
  MST_CHECK_ARG(h && x && mu_query && sigma_query && mu_out && tape, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  int rc;
  if ((rc = check_train_engine(e))) return rc;
  const mst_model_desc& d = e->desc;
  const int B = batch, T = n_frames, S = T + 2, dm = d.d_model;
  MST_CHECK_ARG(B > 0 && T > 0 && S <= d.pe_len, "bad geometry");
  Tape tp;
  MST_CHECK_ARG(tape_bytes >= carve_tape(d, B, S, tape, &tp), "tape too small");
  uint32_t pbits;
  memcpy(&pbits, &dropout_p, 4);
  const void* kp[] = {x, key_valid, mu_query, sigma_query, mu_out, tape, dropout_seed};
  const size_t kn[] = {(size_t)B, (size_t)T, tape_bytes, (size_t)pbits};
  uint64_t key = fnv1a(kp, sizeof(kp), 1469598103934665603ull);
  key = fnv1a(kn, sizeof(kn), key);
  key = fnv1a(e, sizeof(Engine), key) ^ 0x3eull;
  return run_graphed(use_graph != 0, key, (cudaStream_t)stream, [&](cudaStream_t s) -> int {
    int rc;
    if (proj_on_tc(e, B, T)) {
      if ((rc = inproj_tc(e, x, B, T, 2, nullptr, tp.l[0].x, tp.stage.at, s))) return rc;
    } else {
      GemmEx g;  // tokens[(b, t+2)] = x[b,:,t] in_w^T + in_b
      g.a = x; g.a_mode = AX_MOTION_TOK; g.T = T; g.tok_off = 2; g.b = e->in_w; g.ldb = d.n_feats; g.trans_b = 1;
      g.bias = e->in_b; g.c = tp.l[0].x; g.ldc = dm; g.M = B * S; g.N = dm; g.K = d.n_feats;
      if ((rc = gemm_ex(g, s, "menc_inproj"))) return rc;
    }
    MST_CUDA_OK(launch_pdl(menc_tokens_kernel, dim3(B * S), dim3(128), 0, s, mu_query, sigma_query, e->pe, tp.l[0].x, S, dm));
    MST_LAUNCHED("menc_tokens", s);
    if (drop.on() && (rc = dropout(tp.l[0].x, nullptr, tp.l[0].x, (long long)B * S * dm, drop, 0, s))) return rc;
    if ((rc = encoder_forward_tape(e, tp, B, S, key_valid, drop, s))) return rc;
    MST_CUDA_OK(cudaMemcpy2DAsync(mu_out, (size_t)dm * 4, tp.x_out, (size_t)S * dm * 4, (size_t)dm * 4, B,
                                  cudaMemcpyDeviceToDevice, s));
    return MST_OK;
  });
}

// d_x [B,F,T] = d mu / d x (all MotionEncoder parameters are frozen: the reference only needs the input gradient)
extern "C" int mst_motion_encoder_backward(mst_engine_t h, const float* d_mu, int32_t batch, int32_t n_frames, float* d_x,
                                           void* tape, size_t tape_bytes, void* scratch, size_t scratch_bytes,
                                           float dropout_p, const uint64_t* dropout_seed, int32_t use_graph,
                                           void* stream) {
  MST_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f && (dropout_p == 0.0f || dropout_seed), "bad dropout arguments");
  Drop drop;
  drop.p = dropout_p;
  drop.seed = reinterpret_cast<const unsigned long long*>(dropout_seed);
  drop.n_seqs = batch;
  MST_CHECK_ARG(h && d_mu && d_x && tape && scratch, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  int rc;
  if ((rc = check_train_engine(e))) return rc;
  const mst_model_desc& d = e->desc;
  const int B = batch, T = n_frames, S = T + 2, dm = d.d_model;
  Tape tp;
  BwdScratch w;
  MST_CHECK_ARG(tape_bytes >= carve_tape(d, B, S, tape, &tp), "tape too small");
  MST_CHECK_ARG(scratch_bytes >= carve_bwd(d, B, S, scratch, &w), "scratch too small");
  uint32_t pbits;
  memcpy(&pbits, &dropout_p, 4);
  const void* kp[] = {d_mu, d_x, tape, scratch, dropout_seed};
  const size_t kn[] = {(size_t)B, (size_t)T, tape_bytes, scratch_bytes, (size_t)pbits};
  uint64_t key = fnv1a(kp, sizeof(kp), 1469598103934665603ull);
  key = fnv1a(kn, sizeof(kn), key);
  key = fnv1a(e, sizeof(Engine), key) ^ 0x9dull;
  return run_graphed(use_graph != 0, key, (cudaStream_t)stream, [&](cudaStream_t s) -> int {
    int rc;
    float* g = tp.x_out;
    MST_CUDA_OK(cudaMemsetAsync(g, 0, (size_t)B * S * dm * 4, s));
    MST_CUDA_OK(cudaMemcpy2DAsync(g, (size_t)S * dm * 4, d_mu, (size_t)dm * 4, (size_t)dm * 4, B, cudaMemcpyDeviceToDevice, s));
    static const mst_layer_grads kNoGrads[MST_MAX_LAYERS] = {};
    float* gx = nullptr;
    if ((rc = encoder_backward(e, tp, kNoGrads, B, S, g, w, &gx, drop, s))) return rc;
    if (drop.on() && (rc = dropout(gx, nullptr, gx, (long long)B * S * dm, drop, 0, s))) return rc;
    if (proj_on_tc(e, B, T) && e->in_wt_bf) {
      CvtJobs js;  // bf16 operand of the tensor-core product
      cvt_jobs_add(js, gx, B * S, dm, dm, w.stage.a, nullptr, B * S, dm, nullptr, 0, nullptr);
      if ((rc = cvt_multi(js, s, "bwd_operands"))) return rc;
      return inproj_bwd_tc(e, w.stage.a, B, T, 2, d_x, s);
    }
    GemmEx gi;
    gi.a = gx; gi.lda = dm; gi.a_mode = AX_TOKROWS; gi.T = T; gi.tok_off = 2; gi.b = e->in_w; gi.ldb = d.n_feats;
    gi.c = d_x; gi.c_mode = CX_MOTION; gi.M = B * T; gi.N = d.n_feats; gi.K = dm;
    return gemm_ex(gi, s, "menc_bwd_inproj");
  });
}

__global__ void __launch_bounds__(256) dropout_scale_kernel(float* out, long long n4, mst::Drop d, uint32_t site) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(out)[v] = mst::drop_scale4(d, site, 0, v);
}

extern "C" int mst_test_dropout_scale(float* out, int64_t n, float p, const uint64_t* seed_dev, int32_t site, void* stream) {
  MST_CHECK_ARG(out && seed_dev && n > 0 && n % 4 == 0 && p > 0.0f && p < 1.0f, "bad argument");
  Drop d;
  d.p = p;
  d.seed = reinterpret_cast<const unsigned long long*>(seed_dev);
  dropout_scale_kernel<<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(out, n / 4, d, (uint32_t)site);
  MST_LAUNCHED("dropout_scale", (cudaStream_t)stream);
  return MST_OK;
}

extern "C" int mst_masked_l2(const float* a, const float* b, const float* mask, float* loss, const float* grad_loss,
                             float* grad_b, int32_t rows, int32_t a_rows, int32_t mask_rows, int32_t n_feats,
                             int32_t n_frames, void* stream) {
  MST_CHECK_ARG(a && b && mask, "null tensor pointer");
  MST_CHECK_ARG(rows > 0 && a_rows > 0 && mask_rows > 0 && n_feats > 0 && n_frames > 0, "non-positive size");
  MST_CHECK_ARG((loss != nullptr) != (grad_b != nullptr), "give either loss (forward) or grad_loss + grad_b (backward)");
  cudaStream_t s = (cudaStream_t)stream;
  if (loss) {
    masked_l2_fwd_kernel<<<rows, 256, 0, s>>>(a, b, mask, loss, n_feats, n_frames, a_rows, mask_rows);
    MST_LAUNCHED("masked_l2_fwd", s);
  } else {
    MST_CHECK_ARG(grad_loss != nullptr, "backward needs grad_loss");
    masked_l2_bwd_kernel<<<rows, 256, 0, s>>>(a, b, mask, grad_loss, grad_b, n_feats, n_frames, a_rows, mask_rows);
    MST_LAUNCHED("masked_l2_bwd", s);
  }
  return MST_OK;
}

extern "C" int mst_update_step_backward(const float* d_pred_xstart, const float* d_sample, const float* k_table,
                                        const int64_t* t_vec, int32_t mask_kind, const float* mask, const float* pred_xstart,
                                        int32_t clip_denoised, float* d_out, int32_t batch, int32_t n_feats,
                                        int32_t n_frames, void* stream) {
  MST_CHECK_ARG(d_out && (d_pred_xstart || d_sample), "null tensor pointer");
  MST_CHECK_ARG(!d_sample || (k_table && t_vec), "d_sample needs the coefficient table and t");
  MST_CHECK_ARG(mask_kind == MST_MASK_NONE || mask, "mask_kind without a mask");
  MST_CHECK_ARG(!clip_denoised || pred_xstart, "clip_denoised needs pred_xstart");
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)batch * n_feats * n_frames;
  update_bwd_kernel<<<ew_blocks(n), 256, 0, s>>>(d_pred_xstart, d_sample, k_table, t_vec, mask, mask_kind, pred_xstart,
                                                clip_denoised, d_out, batch, n_feats, n_frames);
  MST_LAUNCHED("update_bwd", s);
  return MST_OK;
}

extern "C" int mst_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                              void* stream) {
  MST_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "null tensor pointer");
  MST_CHECK_ARG(n > 0 && step > 0, "n and step must be positive");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  cudaStream_t s = (cudaStream_t)stream;
  adamw_kernel<<<ew_blocks(n), 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                           sqrtf(bc2), grad_scale);
  MST_LAUNCHED("adamw", s);
  return MST_OK;
}

extern "C" int mst_sumsq2(const float* x, const float* y, int64_t n, double* out2, void* stream) {
  MST_CHECK_ARG(x && out2 && n > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  MST_CUDA_OK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), s));
  sumsq2_kernel<<<ew_blocks(n), 256, 0, s>>>(x, y, n, out2);
  MST_LAUNCHED("sumsq2", s);
  return MST_OK;
}
