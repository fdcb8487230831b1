// CLIP text tower: the frozen text encoder MDM.encode_text calls once per trajectory
// (model/mdm_forstyledataset.py:298-313; loaded by load_and_freeze_clip :275-286).  The arithmetic is third-party,
// openai/CLIP @ a9b1bf59 (requirements.txt:26), clip/model.py:
//   CLIP.encode_text            x = token_embedding(text) + positional_embedding; x = transformer(x);
//                               x = ln_final(x); x[arange(B), text.argmax(-1)] @ text_projection
//   ResidualAttentionBlock      x = x + attn(ln_1(x), attn_mask = upper-triangular -inf); x = x + mlp(ln_2(x))
//   mlp                         c_proj(QuickGELU(c_fc(x))),  QuickGELU(x) = x * sigmoid(1.702 x)
//
// One forward = embedding kernel, per block {LN, QKV GEMM, causal attention (one CTA per sequence and head, all of
// Q K V P in shared memory), out-proj GEMM + residual, LN, c_fc GEMM, QuickGELU, c_proj GEMM + residual}, and a final
// kernel that picks the end-of-text row, normalises it and multiplies by text_projection.  The GEMMs are the fp32
// SIMT kernel (MST_PREC_FP32, the parity mode) or the tcgen05 bf16 kernel with fp32 accumulation and fp32 residual
// stream (MST_PREC_BF16); in bf16 mode the row kernels write the GEMM operands directly in bf16.
#include "common.cuh"
#include "simt.cuh"
#include "tc.cuh"
#include "smem_gemm.cuh"

#include <math.h>

namespace mst {

struct ClipText {
  mst_clip_text_desc desc;
  bool loaded = false;
  mst_clip_text_weights w;
  struct Pack { const __nv_bfloat16 *qkv_w, *o_w, *fc_w, *proj_w; } pk[MST_MAX_LAYERS];
};

constexpr int CT_MAXS = 80;          // padded context the attention tiles cover (5 x 16)
constexpr int CT_DH = 64;
constexpr int CT_LDX = CT_DH + 1;    // Q, K, V rows in shared memory
constexpr int CT_LDP = CT_MAXS + 1;  // P rows
constexpr size_t CT_ATTN_SMEM = (size_t)(3 * CT_MAXS * CT_LDX + CT_MAXS * CT_LDP) * sizeof(float);

__device__ __forceinline__ void ct_store(float* o, float v) { *o = v; }
__device__ __forceinline__ void ct_store(__nv_bfloat16* o, float v) { *o = __float2bfloat16_rn(v); }

// x[b, s, :] = token_embedding[tokens[b, s]] + positional_embedding[s]      grid (ctx, batch)
__global__ void __launch_bounds__(128) ct_embed_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ emb,
                                                       const float* __restrict__ pos, float* __restrict__ x, int ctx, int width,
                                                       int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const int s = blockIdx.x, b = blockIdx.y;
  int tok = tokens[(long long)b * ctx + s];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const float4* e4 = reinterpret_cast<const float4*>(emb + (long long)tok * width);
  const float4* p4 = reinterpret_cast<const float4*>(pos + (long long)s * width);
  float4* o4 = reinterpret_cast<float4*>(x + ((long long)b * ctx + s) * width);
  for (int i = threadIdx.x; i < width / 4; i += blockDim.x) {
    const float4 a = e4[i], c = p4[i];
    o4[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

// LayerNorm over the last dimension (eps 1e-5, biased variance, fp32 statistics), one warp per row
template <typename OutT>
__global__ void __launch_bounds__(256) ct_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ b, OutT* __restrict__ y, int M, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long long)row * d;
  float v[32];  // d <= 1024
  float sum = 0.0f;
  const int n = d / 32;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < n) { v[i] = xr[lane + 32 * i]; sum += v[i]; }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)d;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < n) { const float c = v[i] - mean; sq = fmaf(c, c, sq); }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)d + 1e-5f);
  OutT* yr = y + (long long)row * d;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < n) ct_store(yr + lane + 32 * i, (v[i] - mean) * rstd * g[lane + 32 * i] + b[lane + 32 * i]);
}

// QuickGELU: u * sigmoid(1.702 u)
template <typename OutT>
__global__ void __launch_bounds__(256) ct_quickgelu_kernel(const float* __restrict__ u, OutT* __restrict__ h, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = u[i];
    ct_store(h + i, v / (1.0f + expf(-1.702f * v)));
  }
}

// causal softmax(Q K^T / sqrt(64)) V for one (head, sequence): grid (heads, batch).  qkv [batch*S, 3w] -> ao [batch*S, w]
template <typename OutT>
__global__ void __launch_bounds__(256) ct_attention_kernel(const float* __restrict__ qkv, OutT* __restrict__ ao, int S,
                                                           int width) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + CT_MAXS * CT_LDX;
  float* Vs = Ks + CT_MAXS * CT_LDX;
  float* Ps = Vs + CT_MAXS * CT_LDX;
  const int head = blockIdx.x, seq = blockIdx.y;
  const float* base = qkv + (long long)seq * S * 3 * width + head * CT_DH;
  {  // 80 rows x 16 float4 per matrix = 5 per thread; all 15 loads in flight before the first shared-memory store
    constexpr int PER = CT_MAXS * (CT_DH / 4) / 256;
    float4 v[3][PER];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = threadIdx.x + 256 * i, r = idx >> 4, c4 = idx & 15;
        v[m][i] = r < S ? __ldg(reinterpret_cast<const float4*>(base + (long long)r * 3 * width + m * width) + c4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = threadIdx.x + 256 * i, r = idx >> 4, c4 = idx & 15;
        float* d = (m == 0 ? Qs : (m == 1 ? Ks : Vs)) + r * CT_LDX + 4 * c4;
        d[0] = v[m][i].x; d[1] = v[m][i].y; d[2] = v[m][i].z; d[3] = v[m][i].w;
      }
  }
  __syncthreads();
  smem_gemm<false, true, 5, 5>(Qs, CT_LDX, Ks, CT_LDX, CT_DH, [&](int i, int j, float v) { Ps[i * CT_LDP + j] = v * 0.125f; });
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < CT_MAXS; i += 8) {
    float* r = Ps + i * CT_LDP;
    if (i >= S) {
      for (int j = lane; j < CT_MAXS; j += 32) r[j] = 0.0f;
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) mx = fmaxf(mx, r[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int j = lane; j < CT_MAXS; j += 32) {
      const float e = j <= i ? expf(r[j] - mx) : 0.0f;
      r[j] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j <= i; j += 32) r[j] *= inv;
  }
  __syncthreads();
  OutT* og = ao + (long long)seq * S * width + head * CT_DH;
  smem_gemm<false, false, 5, 4>(Ps, CT_LDP, Vs, CT_LDX, CT_MAXS, [&](int i, int c, float v) {
    if (i < S) ct_store(og + (long long)i * width + c, v);
  });
}

// features[b, :] = ln_final(x[b, argmax_s tokens[b, s], :]) @ text_projection        one CTA per sequence
__global__ void __launch_bounds__(256) ct_final_kernel(const float* __restrict__ x, const int32_t* __restrict__ tokens,
                                                       const float* __restrict__ g, const float* __restrict__ bt,
                                                       const float* __restrict__ proj, float* __restrict__ out, int ctx,
                                                       int width, int d_out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float row[1024];
  __shared__ float red[2][8];
  __shared__ int eot;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {  // first position of the largest token id (torch.argmax over the context)
    int best = INT_MIN, pos = 0;
    for (int s = lane; s < ctx; s += 32) {
      const int t = tokens[(long long)b * ctx + s];
      if (t > best) { best = t; pos = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, pos, o);
      if (ob > best || (ob == best && op < pos)) { best = ob; pos = op; }
    }
    if (lane == 0) eot = pos;
  }
  __syncthreads();
  const float* xr = x + ((long long)b * ctx + eot) * width;
  float sum = 0.0f;
  for (int i = threadIdx.x; i < width; i += blockDim.x) { row[i] = xr[i]; sum += row[i]; }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[0][warp] = sum;
  __syncthreads();
  float mean = 0.0f;
  for (int i = 0; i < 8; ++i) mean += red[0][i];
  mean /= (float)width;
  float sq = 0.0f;
  for (int i = threadIdx.x; i < width; i += blockDim.x) { const float c = row[i] - mean; sq = fmaf(c, c, sq); }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) red[1][warp] = sq;
  __syncthreads();
  float var = 0.0f;
  for (int i = 0; i < 8; ++i) var += red[1][i];
  const float rstd = rsqrtf(var / (float)width + 1e-5f);
  for (int i = threadIdx.x; i < width; i += blockDim.x) row[i] = (row[i] - mean) * rstd * g[i] + bt[i];
  __syncthreads();
  for (int n = threadIdx.x; n < d_out; n += blockDim.x) {
    float acc = 0.0f;
    for (int k = 0; k < width; ++k) acc = fmaf(row[k], proj[(long long)k * d_out + n], acc);
    out[(long long)b * d_out + n] = acc;
  }
}

// ---- host side ------------------------------------------------------------------------------------------
struct CtWork {
  float *x, *x2, *qkv, *u;
  void *h, *ao, *g;  // GEMM operands: fp32 (MST_PREC_FP32) or bf16
};

static size_t ct_carve(const mst_clip_text_desc& d, int batch, void* base, CtWork* o) {
  const size_t M = (size_t)batch * d.ctx;
  const size_t esz = d.precision == MST_PREC_BF16 ? 2 : 4;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  float* x = static_cast<float*>(take(M * d.width * 4));
  float* x2 = static_cast<float*>(take(M * d.width * 4));
  float* qkv = static_cast<float*>(take(M * 3 * d.width * 4));
  float* u = static_cast<float*>(take(M * d.d_ff * 4));
  void* h = take(M * d.width * esz);
  void* ao = take(M * d.width * esz);
  void* g = take(M * d.d_ff * esz);
  if (o) *o = CtWork{x, x2, qkv, u, h, ao, g};
  return off;
}

static size_t ct_packed_bytes(const mst_clip_text_desc& d) {
  if (d.precision != MST_PREC_BF16) return 0;
  const size_t per = align_up((size_t)3 * d.width * d.width * 2, 256) + align_up((size_t)d.width * d.width * 2, 256) +
                     2 * align_up((size_t)d.d_ff * d.width * 2, 256);
  return per * d.n_layers;
}

// out [M, n_out] = a [M, n_in] W^T + bias (+ add)
static int ct_linear(const ClipText* e, const void* a, const float* w, const __nv_bfloat16* w_bf, const float* bias,
                     const float* add, float* out, int M, int n_out, int n_in, cudaStream_t s) {
  if (e->desc.precision == MST_PREC_BF16) {
    TcGemmParams p;
    p.a = static_cast<const __nv_bfloat16*>(a); p.w = w_bf; p.bias = bias; p.add = add; p.out = out; p.ldo = n_out;
    p.M = M; p.N = n_out; p.K = n_in; p.epi = TC_EPI_TRAIN_F32;
    return tc_gemm(p, s);
  }
  GemmF32Params p;
  p.a = static_cast<const float*>(a); p.w = w; p.bias = bias; p.c = out; p.M = M; p.N = n_out; p.K = n_in;
  p.lda = n_in; p.ldw = n_in; p.ldc = n_out;
  if (add) { p.epi = EPI_RESIDUAL; p.residual = add; }
  return gemm_f32(p, s);
}

template <typename OutT>
static int ct_forward(ClipText* e, const int32_t* tokens, int batch, float* features, const CtWork& wk, cudaStream_t s) {
  const mst_clip_text_desc& d = e->desc;
  const int M = batch * d.ctx;
  MST_CUDA_OK(launch_pdl(ct_embed_kernel, dim3(d.ctx, batch), dim3(128), 0, s, tokens, e->w.token_embedding,
                         e->w.positional_embedding, wk.x, d.ctx, d.width, d.vocab));
  MST_LAUNCHED("clip_text_embed", s);
  static PerDeviceOnce smem_set;  // cudaFuncSetAttribute is per device
  if (smem_set.first()) {
    MST_CUDA_OK(cudaFuncSetAttribute(ct_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_ATTN_SMEM));
    MST_CUDA_OK(cudaFuncSetAttribute(ct_attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_ATTN_SMEM));
  }
  float* x = wk.x;
  float* x2 = wk.x2;
  OutT* h = static_cast<OutT*>(wk.h);
  OutT* ao = static_cast<OutT*>(wk.ao);
  OutT* g = static_cast<OutT*>(wk.g);
  const long long n_ff = (long long)M * d.d_ff;
  const int ew = (int)((n_ff + 255) / 256 < 8 * sm_count() ? (n_ff + 255) / 256 : 8 * sm_count());
  int rc;
  for (int l = 0; l < d.n_layers; ++l) {
    const mst_clip_text_layer& L = e->w.layers[l];
    MST_CUDA_OK(launch_pdl(ct_layernorm_kernel<OutT>, dim3(ceil_div(M, 8)), dim3(256), 0, s, (const float*)x, L.ln1_g, L.ln1_b, h, M, d.width));
    MST_LAUNCHED("clip_text_ln", s);
    if ((rc = ct_linear(e, h, L.qkv_w, e->pk[l].qkv_w, L.qkv_b, nullptr, wk.qkv, M, 3 * d.width, d.width, s))) return rc;
    MST_CUDA_OK(launch_pdl(ct_attention_kernel<OutT>, dim3(d.n_heads, batch), dim3(256), CT_ATTN_SMEM, s, (const float*)wk.qkv, ao, d.ctx, d.width));
    MST_LAUNCHED("clip_text_attention", s);
    if ((rc = ct_linear(e, ao, L.o_w, e->pk[l].o_w, L.o_b, x, x2, M, d.width, d.width, s))) return rc;
    MST_CUDA_OK(launch_pdl(ct_layernorm_kernel<OutT>, dim3(ceil_div(M, 8)), dim3(256), 0, s, (const float*)x2, L.ln2_g, L.ln2_b, h, M, d.width));
    MST_LAUNCHED("clip_text_ln", s);
    if ((rc = ct_linear(e, h, L.fc_w, e->pk[l].fc_w, L.fc_b, nullptr, wk.u, M, d.d_ff, d.width, s))) return rc;
    MST_CUDA_OK(launch_pdl(ct_quickgelu_kernel<OutT>, dim3(ew), dim3(256), 0, s, (const float*)wk.u, g, n_ff));
    MST_LAUNCHED("clip_text_quickgelu", s);
    if ((rc = ct_linear(e, g, L.proj_w, e->pk[l].proj_w, L.proj_b, x2, x, M, d.width, d.d_ff, s))) return rc;
  }
  MST_CUDA_OK(launch_pdl(ct_final_kernel, dim3(batch), dim3(256), 0, s, (const float*)x, tokens, e->w.lnf_g, e->w.lnf_b,
                         e->w.text_projection, features, d.ctx, d.width, d.d_out));
  MST_LAUNCHED("clip_text_final", s);
  return MST_OK;
}

}  // namespace mst

using namespace mst;

extern "C" int mst_abi_sizes_clip_text(size_t* desc, size_t* weights) {
  if (desc) *desc = sizeof(mst_clip_text_desc);
  if (weights) *weights = sizeof(mst_clip_text_weights);
  return MST_OK;
}

extern "C" int mst_clip_text_create(const mst_clip_text_desc* desc, mst_clip_text_t* out) {
  MST_CHECK_ARG(desc && out, "null argument");
  const mst_clip_text_desc& d = *desc;
  MST_CHECK_ARG(d.vocab > 0 && d.ctx > 0 && d.width > 0 && d.n_heads > 0 && d.d_ff > 0 && d.d_out > 0, "non-positive dimension");
  MST_CHECK_ARG(d.n_layers > 0 && d.n_layers <= MST_MAX_LAYERS, "n_layers out of range");
  MST_CHECK_ARG(d.precision == MST_PREC_FP32 || d.precision == MST_PREC_BF16, "unknown precision");
  if (d.ctx > CT_MAXS) return fail(MST_ERR_UNSUPPORTED, "mst_clip_text_create: context length above 80 tokens");
  if (d.width != d.n_heads * CT_DH) return fail(MST_ERR_UNSUPPORTED, "mst_clip_text_create: head_dim must be 64");
  if (d.width % 64 != 0 || d.width > 1024 || d.d_ff % 64 != 0)
    return fail(MST_ERR_UNSUPPORTED, "mst_clip_text_create: width must be a multiple of 64 and <= 1024, d_ff a multiple of 64");
  ClipText* e = new ClipText();
  e->desc = d;
  *out = reinterpret_cast<mst_clip_text_t>(e);
  return MST_OK;
}

extern "C" int mst_clip_text_destroy(mst_clip_text_t h) {
  delete reinterpret_cast<ClipText*>(h);
  return MST_OK;
}

extern "C" int mst_clip_text_packed_weight_bytes(mst_clip_text_t h, size_t* bytes) {
  MST_CHECK_ARG(h && bytes, "null argument");
  *bytes = ct_packed_bytes(reinterpret_cast<ClipText*>(h)->desc);
  return MST_OK;
}

extern "C" int mst_clip_text_workspace_bytes(mst_clip_text_t h, int32_t batch, size_t* bytes) {
  MST_CHECK_ARG(h && bytes, "null argument");
  MST_CHECK_ARG(batch > 0, "batch must be positive");
  *bytes = ct_carve(reinterpret_cast<ClipText*>(h)->desc, batch, nullptr, nullptr);
  return MST_OK;
}

extern "C" int mst_clip_text_load_weights(mst_clip_text_t h, const mst_clip_text_weights* w, void* packed_dev,
                                          size_t packed_bytes, void* stream) {
  MST_CHECK_ARG(h && w, "null argument");
  ClipText* e = reinterpret_cast<ClipText*>(h);
  const mst_clip_text_desc& d = e->desc;
  MST_CHECK_ARG(w->token_embedding && w->positional_embedding && w->lnf_g && w->lnf_b && w->text_projection,
                "missing weight pointer");
  for (int l = 0; l < d.n_layers; ++l) {
    const mst_clip_text_layer& L = w->layers[l];
    MST_CHECK_ARG(L.ln1_g && L.ln1_b && L.qkv_w && L.qkv_b && L.o_w && L.o_b && L.ln2_g && L.ln2_b && L.fc_w && L.fc_b &&
                      L.proj_w && L.proj_b,
                  "missing layer weight pointer");
  }
  e->w = *w;
  if (d.precision == MST_PREC_BF16) {
    MST_CHECK_ARG(packed_dev && packed_bytes >= ct_packed_bytes(d), "packed weight buffer too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(packed_dev);
    int rc;
    auto pack = [&](const float* src, int rows, int cols, const __nv_bfloat16** dst) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p);
      p += align_up((size_t)rows * cols * 2, 256);
      *dst = o;
      return pack_bf16(src, o, rows, cols, rows, cols, s);
    };
    for (int l = 0; l < d.n_layers; ++l) {
      const mst_clip_text_layer& L = w->layers[l];
      if ((rc = pack(L.qkv_w, 3 * d.width, d.width, &e->pk[l].qkv_w))) return rc;
      if ((rc = pack(L.o_w, d.width, d.width, &e->pk[l].o_w))) return rc;
      if ((rc = pack(L.fc_w, d.d_ff, d.width, &e->pk[l].fc_w))) return rc;
      if ((rc = pack(L.proj_w, d.width, d.d_ff, &e->pk[l].proj_w))) return rc;
    }
  }
  e->loaded = true;
  return MST_OK;
}

extern "C" int mst_clip_text_encode(mst_clip_text_t h, const int32_t* tokens, int32_t batch, float* features,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  MST_CHECK_ARG(h && tokens && features && workspace, "null argument");
  ClipText* e = reinterpret_cast<ClipText*>(h);
  MST_CHECK_ARG(e->loaded, "weights not loaded");
  MST_CHECK_ARG(batch > 0, "batch must be positive");
  CtWork wk;
  const size_t need = ct_carve(e->desc, batch, workspace, &wk);
  MST_CHECK_ARG(workspace_bytes >= need, "workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (e->desc.precision == MST_PREC_BF16) return ct_forward<__nv_bfloat16>(e, tokens, batch, features, wk, s);
  return ct_forward<float>(e, tokens, batch, features, wk, s);
}
