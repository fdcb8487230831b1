// Persistent, warp-specialised tcgen05 GEMM for the denoiser's dense layers
// (MST_PREC_BF16):   D[M,N] = A[M,K] * W[N,K]^T  with fused epilogues.
//
//   A, W : bf16, K-major (row-major [rows, K]); TMA 128B-swizzled 128x64 / BNx64 tiles
//   D    : fp32 accumulators in TMEM, 2 stages of BN columns (epilogue of tile i
//          overlaps the MMAs of tile i+1)
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9  : epilogue, 8 warps; warp w may touch TMEM lanes 32*(w%4)..+31
//
// Layers served (reference model/mdm_forstyledataset.py):
//   InputProcess.poseEmbedding (:440) + positional add (:403)       TC_EPI_INPROJ
//   self_attn.in_proj (QKV)                                        TC_EPI_BIAS_BF16
//   self_attn.out_proj + residual + norm1                          TC_EPI_BIAS_RES_LN
//   linear1 + GELU                                                 TC_EPI_BIAS_GELU_BF16
//   linear2 + residual + norm2                                     TC_EPI_BIAS_RES_LN
//   OutputProcess.poseFinal + permute back to [B,F,1,T] (:467-477) TC_EPI_OUTPROJ_F32
#include "tc.cuh"
#include "tc_ptx.cuh"

#include <mutex>
#include <unordered_map>

namespace mst {

using namespace ptx;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int GEMM_THREADS = 64 + NUM_EPI_THREADS;  // 320
constexpr int LN_N = 512;                           // row width the LN epilogue is built for

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN >= 256 ? 4 : 6;
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN >= 512 ? 512 : (2 * BN >= 256 ? 256 : (2 * BN >= 128 ? 128 : 64));
  static constexpr int BAR_BYTES = 256;
  static constexpr int EXCH_BYTES = 2 * 2 * BLOCK_M * 8;  // [buf][group][row] float2
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES + EXCH_BYTES;
};

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---- epilogue of one 32-column chunk held by one thread (= one output row) ----
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const TcGemmParams& p, int row, int n, const uint32_t (&v)[32]) {
  if (row >= p.M) return;
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
    x[j] = __uint_as_float(v[j]) + b4.x;
    x[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
    x[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
    x[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
  }
  if constexpr (EPI == TC_EPI_BIAS_BF16 || EPI == TC_EPI_BIAS_GELU_BF16) {
    if constexpr (EPI == TC_EPI_BIAS_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = gelu_erf_f(x[j]);
    }
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                          pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
  } else if constexpr (EPI == TC_EPI_INPROJ) {
    const int b = row / p.T, t = row - b * p.T, S = p.T + 1;
    const float* pe = p.pe + (size_t)(t + 1) * p.N + n;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 e = __ldg(reinterpret_cast<const float4*>(pe + j));
      x[j] += e.x; x[j + 1] += e.y; x[j + 2] += e.z; x[j + 3] += e.w;
    }
    for (int pass = 0; pass < p.n_pass; ++pass) {
      uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) +
                                            ((size_t)(pass * p.B + b) * S + t + 1) * p.ldo + n);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                            pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
    }
  } else if constexpr (EPI == TC_EPI_OUTPROJ_F32) {
    const int S = p.T + 1;
    const int seq = row / S, s = row - seq * S;
    if (s == 0) return;
    float* base = (seq < p.B) ? static_cast<float*>(p.out) : p.out2;
    const int sb = (seq < p.B) ? seq : seq - p.B;
    float* dst = base + ((size_t)sb * p.n_valid + n) * p.T + (s - 1);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < p.n_valid) dst[(size_t)j * p.T] = x[j];
  } else if constexpr (EPI == TC_EPI_BIAS_F32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)row * p.ldo + n);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const TcGemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr bool IS_LN = (EPI == TC_EPI_BIAS_RES_LN);
  constexpr int NPT = IS_LN ? 2 : 1;  // accumulator sub-tiles per scheduled tile
  static_assert(!IS_LN || BN == 256, "LN epilogue is built for 2 x 256 columns");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto a_smem = [&](int st) { return base + st * Cfg::STAGE_BYTES; };
  auto b_smem = [&](int st) { return base + st * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto full_bar = [&](int st) { return bar_base + 8 * st; };
  auto empty_bar = [&](int st) { return bar_base + 8 * (Cfg::STAGES + st); };
  auto tfull_bar = [&](int i) { return bar_base + 8 * (2 * Cfg::STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8 * (2 * Cfg::STAGES + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8 * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));
  float2* exch = reinterpret_cast<float2*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_w);
    for (int st = 0; st < Cfg::STAGES; ++st) {
      mbar_init(full_bar(st), 1);
      mbar_init(empty_bar(st), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), IS_LN ? NUM_EPI_THREADS / 2 : NUM_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_groups = (p.N / BN) / NPT;
  const int m_blks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_tiles = m_blks * n_groups;
  const int k_blks = p.K / BLOCK_K;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_groups, g = tile - m_blk * n_groups;
        for (int sub = 0; sub < NPT; ++sub) {
          const int n_blk = g * NPT + sub;
          for (int kb = 0; kb < k_blks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            tma_load_2d(a_smem(stage), &tmap_a, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(b_smem(stage), &tmap_w, full_bar(stage), kb * BLOCK_K, n_blk * BN);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int sub = 0; sub < NPT; ++sub) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < k_blks; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(a_smem(stage) + k * (UMMA_K * 2), 0, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(b_smem(stage) + k * (UMMA_K * 2), 0, 1024);
              mma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            mma_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
          mma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // -------------------------------- epilogue --------------------------------
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp is allowed to access
    const int half = ew >> 2;   // column half (plain) / accumulator stage (LN)
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    if constexpr (!IS_LN) {
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_groups, n_blk = tile - m_blk * n_groups;
        const int row = m_blk * BLOCK_M + row_in_tile;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 32) {
          const int col = half * (BN / 2) + c;
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_addr + (uint32_t)(acc * BN + col), v);
          tmem_ld_wait();
          epilogue_chunk<EPI>(p, row, n_blk * BN + col, v);
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    } else {
      // LN(acc + bias + residual): group `half` owns accumulator stage `half`
      // (columns [256*half, 256*half+256) of the 512-wide row); row statistics are
      // exchanged through shared memory between the two groups.
      uint32_t tile_phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int row = tile * BLOCK_M + row_in_tile;  // n_groups == 1
        const bool valid = row < p.M;
        const int ncol0 = half * BN;
        const uint32_t tcol = tmem_base + lane_addr + (uint32_t)(half * BN);
        mbar_wait(tfull_bar(half), tile_phase);
        tc_fence_after();
        float sum = 0.0f, sq = 0.0f;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(tcol + c, v);
          uint4 r4[4];
          if (valid) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * LN_N + ncol0 + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) r4[q] = __ldg(rp + q);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) r4[q] = make_uint4(0, 0, 0, 0);
          }
          tmem_ld_wait();
          const uint32_t* ru = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 rr = unpack_bf16x2(ru[j >> 1]);
            float2 bb = __ldg(reinterpret_cast<const float2*>(p.bias + ncol0 + c + j));
            float x0 = __uint_as_float(v[j]) + bb.x + rr.x;
            float x1 = __uint_as_float(v[j + 1]) + bb.y + rr.y;
            sum += x0 + x1;
            sq = fmaf(x0, x0, fmaf(x1, x1, sq));
            v[j] = __float_as_uint(x0);
            v[j + 1] = __float_as_uint(x1);
          }
          tmem_st32(tcol + c, v);
        }
        tmem_st_wait();
        float2* ex = exch + (size_t)(it & 1) * 2 * BLOCK_M;
        ex[half * BLOCK_M + row_in_tile] = make_float2(sum, sq);
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");
        const float2 other = ex[(half ^ 1) * BLOCK_M + row_in_tile];
        const float mean = (sum + other.x) * (1.0f / LN_N);
        const float var = fmaxf((sq + other.y) * (1.0f / LN_N) - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(tcol + c, v);
          tmem_ld_wait();
          if (valid) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float2 g = __ldg(reinterpret_cast<const float2*>(p.ln_g + ncol0 + c + j));
              float2 b = __ldg(reinterpret_cast<const float2*>(p.ln_b + ncol0 + c + j));
              float y0 = fmaf((__uint_as_float(v[j]) - mean) * rstd, g.x, b.x);
              float y1 = fmaf((__uint_as_float(v[j + 1]) - mean) * rstd, g.y, b.y);
              o[j >> 1] = pack_bf16x2(y0, y1);
            }
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * LN_N + ncol0 + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(half));
        tile_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t rows, cols, stride;
  uint32_t box_rows, box_cols;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && stride == o.stride && box_rows == o.box_rows &&
           box_cols == o.box_cols;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    auto mix = [&](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.stride); mix(k.box_rows); mix(k.box_cols);
    return h;
  }
};

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                   uint32_t box_rows, uint32_t box_cols) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{base, rows, cols, row_stride_elems, box_rows, box_cols};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return MST_OK;
    }
  }
  EncodeTiledFn enc = get_encoder();
  if (!enc) return fail(MST_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * 2) % 16 != 0)
    return fail(MST_ERR_INVALID, "make_tmap_bf16: base and row stride must be 16-byte aligned");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MST_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  {
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = m;
  }
  *out = m;
  return MST_OK;
}

template <int BN, int EPI>
static int launch_gemm(const TcGemmParams& p, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tw;
  int rc;
  if ((rc = make_tmap_bf16(&ta, p.a, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K, BLOCK_M, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16(&tw, p.w, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, BN, BLOCK_K))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MST_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int npt = (EPI == TC_EPI_BIAS_RES_LN) ? 2 : 1;
  const int tiles = ceil_div(p.M, BLOCK_M) * ((p.N / BN) / npt);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  tc_gemm_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(ta, tw, p);
  static const char* const kNames[] = {"tc_gemm_qkv", "tc_gemm_ffn1_gelu", "tc_gemm_res_ln", "tc_gemm_inproj",
                                       "tc_gemm_outproj", "tc_gemm_f32"};
  MST_LAUNCHED(kNames[EPI], s);
  return MST_OK;
}

int tc_gemm(const TcGemmParams& p, cudaStream_t s) {
  MST_CHECK_ARG(p.a && p.w && p.bias && p.out, "null pointer");
  MST_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0, "empty problem");
  MST_CHECK_ARG(p.K % BLOCK_K == 0, "K must be a multiple of 64");
  switch (p.epi) {
    case TC_EPI_BIAS_BF16:
      MST_CHECK_ARG(p.N % 256 == 0 && p.ldo % 8 == 0, "N must be a multiple of 256");
      return launch_gemm<256, TC_EPI_BIAS_BF16>(p, s);
    case TC_EPI_BIAS_GELU_BF16:
      MST_CHECK_ARG(p.N % 256 == 0 && p.ldo % 8 == 0, "N must be a multiple of 256");
      return launch_gemm<256, TC_EPI_BIAS_GELU_BF16>(p, s);
    case TC_EPI_BIAS_RES_LN:
      MST_CHECK_ARG(p.N == LN_N && p.residual && p.ln_g && p.ln_b, "LN epilogue needs N == 512 and residual/gamma/beta");
      return launch_gemm<256, TC_EPI_BIAS_RES_LN>(p, s);
    case TC_EPI_INPROJ:
      MST_CHECK_ARG(p.N % 256 == 0 && p.pe && p.B > 0 && p.T > 0 && p.ldo % 8 == 0, "bad in-projection geometry");
      return launch_gemm<256, TC_EPI_INPROJ>(p, s);
    case TC_EPI_OUTPROJ_F32:
      MST_CHECK_ARG(p.N % 64 == 0 && p.B > 0 && p.T > 0 && p.n_valid > 0, "bad out-projection geometry");
      return launch_gemm<64, TC_EPI_OUTPROJ_F32>(p, s);
    case TC_EPI_BIAS_F32:
      MST_CHECK_ARG(p.N % 64 == 0 && p.ldo % 4 == 0, "N must be a multiple of 64");
      if (p.N % 256 == 0) return launch_gemm<256, TC_EPI_BIAS_F32>(p, s);
      return launch_gemm<64, TC_EPI_BIAS_F32>(p, s);
    default:
      return fail(MST_ERR_INVALID, "tc_gemm: unknown epilogue");
  }
}

// ---------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------
// x[b][f][t] fp32 -> a[(b*T + t)][f] bf16, columns [F, f_pad) zero.  32x32 smem transpose.
__global__ void __launch_bounds__(256) motion_to_tokens_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a,
                                                               int F, int T, int f_pad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int f = f0 + i, t = t0 + tx;
    tile[i][tx] = (f < F && t < T) ? x[((size_t)b * F + f) * T + t] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int t = t0 + i, f = f0 + tx;
    if (t < T && f < f_pad) a[((size_t)b * T + t) * f_pad + f] = __float2bfloat16_rn(tile[tx][i]);
  }
}

int motion_to_tokens_bf16(const float* x, __nv_bfloat16* a, int B, int F, int T, int f_pad, cudaStream_t s) {
  dim3 grid(ceil_div(T, 32), ceil_div(f_pad, 32), B);
  motion_to_tokens_kernel<<<grid, 256, 0, s>>>(x, a, F, T, f_pad);
  MST_LAUNCHED("motion_to_tokens", s);
  return MST_OK;
}

__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        int rows, int cols, int rows_pad, int cols_pad) {
  const size_t total = (size_t)rows_pad * cols_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / cols_pad), c = (int)(i - (size_t)r * cols_pad);
    dst[i] = __float2bfloat16_rn((r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.0f);
  }
}

int pack_bf16(const float* src, __nv_bfloat16* dst, int rows, int cols, int rows_pad, int cols_pad, cudaStream_t s) {
  size_t total = (size_t)rows_pad * cols_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_bf16_kernel<<<blocks, 256, 0, s>>>(src, dst, rows, cols, rows_pad, cols_pad);
  MST_LAUNCHED("pack_bf16", s);
  return MST_OK;
}

}  // namespace mst
