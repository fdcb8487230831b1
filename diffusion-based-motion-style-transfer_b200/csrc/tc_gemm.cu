// Persistent, warp-specialised tcgen05 GEMMs for the denoiser's dense layers
// (MST_PREC_BF16):   D[M,N] = A[M,K] * W[N,K]^T  with fused epilogues.
//
//   A, W : bf16, K-major (row-major [rows, K]); TMA 128B-swizzled 128x64 / 256x64 tiles, 3-4 stage ring
//   D    : fp32 accumulators in TMEM, 2 stages of 256 columns (the epilogue of tile i overlaps the
//          MMAs of tile i+1)
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9  : epilogue, 8 warps; warp w owns TMEM lanes 32*(w%4)..+31 (one output row per thread),
//                 warps 2-5 the low column half of the tile, warps 6-9 the high half
//
// Two kernels:
//   tc_gemm_kernel<BN,EPI>   one CTA per 128 x BN tile.  bf16 outputs are staged through shared memory in
//                            the TMA swizzle layout and written with cp.async.bulk.tensor stores (coalesced
//                            and asynchronous: a row-per-thread direct store was the bottleneck,
//                            profiles/r01a_*); the two small fp32 / row-remapping epilogues store directly.
//   tc_gemm_ln_kernel        out = LayerNorm(A W^T + bias + residual) for N = 512: a 2-CTA cluster owns a
//                            128-row block, each CTA 256 of the 512 columns, so every CTA has two
//                            accumulator stages and the MMAs never wait for the normalisation; the row
//                            statistics cross the CTA pair through distributed shared memory.
//
// Layers served (reference model/mdm_forstyledataset.py):
//   InputProcess.poseEmbedding (:440) + positional add (:403)       TC_EPI_INPROJ
//   self_attn.in_proj (QKV)                                        TC_EPI_BIAS_BF16
//   self_attn.out_proj + residual + norm1                          TC_EPI_BIAS_RES_LN
//   linear1 + GELU                                                 TC_EPI_BIAS_GELU_BF16
//   linear2 + residual + norm2                                     TC_EPI_BIAS_RES_LN
//   OutputProcess.poseFinal + permute back to [B,F,1,T] (:467-477) TC_EPI_OUTPROJ_F32
#include "simt.cuh"
#include "tc.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

namespace mst {

using namespace ptx;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int GEMM_THREADS = 64 + NUM_EPI_THREADS;  // 320
constexpr int LN_N = 512;                           // row width the LN epilogue is built for

#define MST_DBG_STAMP() do { if (dbg && di < 1000) dbg[di++] = clock64(); } while (0)
// wall-clock (ns) of CTA entry (slot 0), end of work (slot 1) and last instruction (slot 2), per CTA
#define MST_DBG_WALL(slot)                                                      \
  do {                                                                          \
    if (p.dbg && threadIdx.x == 0) {                                            \
      unsigned long long gt_;                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                   \
      p.dbg[6 * 1024 + 3 * blockIdx.x + (slot)] = (long long)gt_;               \
    }                                                                           \
  } while (0)
// true end of a warp's work (no barrier in between: a timer read right after BAR.SYNC sees the ISSUE time of the
// barrier, not its release) - atomicMax over the warps of the CTA
#define MST_DBG_WALL_END()                                                      \
  do {                                                                          \
    if (p.dbg && lane == 0) {                                                   \
      unsigned long long gt_;                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                   \
      atomicMax(reinterpret_cast<unsigned long long*>(p.dbg + 6 * 1024 + 3 * blockIdx.x + 1), gt_); \
    }                                                                           \
  } while (0)
#define MST_DBG_WALL_W1(slot)                                                   \
  do {                                                                          \
    if (p.dbg && threadIdx.x == 32) {                                           \
      unsigned long long gt_;                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                   \
      p.dbg[7 * 1024 + 2 * blockIdx.x + (slot)] = (long long)gt_;               \
    }                                                                           \
  } while (0)

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN >= 256 ? 4 : 6;
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN >= 512 ? 512 : (2 * BN >= 256 ? 256 : (2 * BN >= 128 ? 128 : 64));
  static constexpr int BAR_BYTES = 256;
  static constexpr int OBOX_BYTES = 32 * 64;  // in-projection epilogue: one [32 rows x 32 cols] transpose box per warp
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES + NUM_EPI_WARPS * OBOX_BYTES;
};

// ---- GELU for the FFN epilogue --------------------------------------------------------------------------
// 0.5 x (1 + erf(x / sqrt 2)) with erf(z) = z P(z^2) on |z| <= 3 (degree-8 minimax fit, |erf error| < 4.2e-5,
// |GELU error| < 9e-5 - two orders below the bf16 rounding of the output) and z clamped outside.
// No MUFU: an exp/rcp based erf needs 2 SFU ops per element = 4096 SFU cycles per 128x256 tile, as long as the
// tile's MMAs.  The polynomial runs on packed fma.rn.f32x2 (two elements per instruction).
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
// two bf16 (packed in 32 bits) -> two fp32 (packed in 64 bits): shift / mask, no conversion instruction
__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t u) { return pk2u(u << 16, u & 0xffff0000u); }
// gelu(x) = x * (0.5 + z q(z^2)),  z = clamp(x / sqrt 2, +-2.8),  q = degree-6 minimax fit of erf(z) / (2z)
// (|relative error| < 2e-4, an order below the bf16 rounding of the output).  Two elements per instruction.
__device__ __forceinline__ uint64_t gelu2(uint64_t x) {
  const float kC[7] = {5.6408941812e-01f, -1.8671857437e-01f, 5.3391947352e-02f, -1.0763958331e-02f,
                       1.4033610804e-03f, -1.0399474066e-04f, 3.2816237409e-06f};
  float x0, x1;
  upk2(x, x0, x1);
  const float z0 = fminf(fmaxf(x0 * 0.70710678118654752440f, -2.8f), 2.8f);
  const float z1 = fminf(fmaxf(x1 * 0.70710678118654752440f, -2.8f), 2.8f);
  const uint64_t z = pk2(z0, z1);
  const uint64_t t = mul2(z, z);
  uint64_t q = pk2(kC[6], kC[6]);
#pragma unroll
  for (int k = 5; k >= 0; --k) q = fma2(q, t, pk2(kC[k], kC[k]));
  return mul2(x, fma2(z, q, pk2(0.5f, 0.5f)));
}

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---- direct (row-per-thread) epilogue of one 32-column chunk: the small fp32 / row-remapping cases ----
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const TcGemmParams& p, int row, int n, const uint32_t (&v)[32]) {
  if (row >= p.M) return;
  float x[32];
  if constexpr (EPI == TC_EPI_TRAIN_F32) {
    int orow = row;
    const float* pe_row = nullptr;
    if (p.tok_T > 0) {  // in-projection of the taped forward: (b, t) -> token row (b, t + tok_off)
      const int b = row / p.tok_T, t = row - b * p.tok_T;
      orow = b * (p.tok_T + p.tok_off) + t + p.tok_off;
      if (p.pe) pe_row = p.pe + (size_t)(t + p.tok_off) * p.N + n;
    }
    float* dst = static_cast<float*>(p.out) + (size_t)orow * p.ldo + n;
    const float* addp = p.add ? p.add + (size_t)orow * p.ldo + n : nullptr;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 r = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      if (p.bias) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        r.x += b4.x; r.y += b4.y; r.z += b4.z; r.w += b4.w;
      }
      if (pe_row) {
        const float4 e4 = __ldg(reinterpret_cast<const float4*>(pe_row + j));
        r.x += e4.x; r.y += e4.y; r.z += e4.z; r.w += e4.w;
      }
      if (addp) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(addp + j));
        r.x += a4.x; r.y += a4.y; r.z += a4.z; r.w += a4.w;
      }
      if (p.accumulate) {
        const float4 o4 = *reinterpret_cast<const float4*>(dst + j);
        r.x += o4.x; r.y += o4.y; r.z += o4.z; r.w += o4.w;
      }
      *reinterpret_cast<float4*>(dst + j) = r;
      if (p.gelu_out) {  // h = dropout(gelu(u)): the same erf GELU and the same mask counters as gelu_drop_fwd_kernel
        float4 h = make_float4(gelu_erf_f(r.x), gelu_erf_f(r.y), gelu_erf_f(r.z), gelu_erf_f(r.w));
        if (p.drop.on()) {
          const int seq = row / p.rows_per_seq;
          const long long local4 = (long long)(row - seq * p.rows_per_seq) * (p.ldo >> 2) + ((n + j) >> 2);
          const float4 m = drop_scale4(p.drop, p.drop_site, seq, local4);
          h.x *= m.x; h.y *= m.y; h.z *= m.z; h.w *= m.w;
        }
        *reinterpret_cast<float4*>(p.gelu_out + (size_t)row * p.ldo + n + j) = h;
        if (p.gelu_bf) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
          *reinterpret_cast<uint2*>(p.gelu_bf + (size_t)row * p.ldo + n + j) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
    x[j] = __uint_as_float(v[j]) + b4.x;
    x[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
    x[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
    x[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
  }
  if constexpr (EPI == TC_EPI_INPROJ) {
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
        dst[q] = p.io_f16 ? make_uint4(pack_f16x2(x[8 * q], x[8 * q + 1]), pack_f16x2(x[8 * q + 2], x[8 * q + 3]),
                                       pack_f16x2(x[8 * q + 4], x[8 * q + 5]), pack_f16x2(x[8 * q + 6], x[8 * q + 7]))
                          : make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                                       pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
    }
  } else if constexpr (EPI == TC_EPI_OUTPROJ_F32) {
    const int tok = p.drop_tokens > 0 ? p.drop_tokens : 1;
    const int S = p.T + tok;
    const int seq = row / S, s = row - seq * S;
    if (s < tok) return;
    float* base = (seq < p.B) ? static_cast<float*>(p.out) : p.out2;
    const int sb = (seq < p.B) ? seq : seq - p.B;
    float* dst = base + ((size_t)sb * p.n_valid + n) * p.T + (s - tok);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < p.n_valid) dst[(size_t)j * p.T] = x[j];
  } else if constexpr (EPI == TC_EPI_BIAS_BF16 || EPI == TC_EPI_BIAS_GELU_BF16) {
    // small problems only (see tc_gemm: a handful of row blocks): row-per-thread 16-byte stores are fine there
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint64_t v2 = pk2(x[8 * q + 2 * e], x[8 * q + 2 * e + 1]);
        if constexpr (EPI == TC_EPI_BIAS_GELU_BF16) v2 = gelu2(v2);
        float y0, y1;
        upk2(v2, y0, y1);
        o[e] = pack_bf16x2(y0, y1);
      }
      dst[q] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else if constexpr (EPI == TC_EPI_BIAS_F32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)row * p.ldo + n);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  }
}

// ---- smem ring + barrier addresses shared by the kernels ---------------------------------------------
struct Ring {
  uint32_t base, bar_base;
  int stages, stage_bytes, a_bytes;
  __device__ __forceinline__ uint32_t a(int st) const { return base + st * stage_bytes; }
  __device__ __forceinline__ uint32_t b(int st) const { return base + st * stage_bytes + a_bytes; }
  __device__ __forceinline__ uint32_t full(int st) const { return bar_base + 8 * st; }
  __device__ __forceinline__ uint32_t empty(int st) const { return bar_base + 8 * (stages + st); }
  __device__ __forceinline__ uint32_t tfull(int i) const { return bar_base + 8 * (2 * stages + i); }
  __device__ __forceinline__ uint32_t tempty(int i) const { return bar_base + 8 * (2 * stages + 2 + i); }
  __device__ __forceinline__ uint32_t extra(int i) const { return bar_base + 8 * (2 * stages + 4 + i); }
};

template <int BN>
__device__ __forceinline__ void mma_tile(const Ring& r, uint32_t d_tmem, int k_blks, int& stage, uint32_t& phase, bool ab_f16) {
  const uint32_t idesc = ab_f16 ? make_idesc_f16(BLOCK_M, BN, 0) : make_idesc_bf16(BLOCK_M, BN, 0);
  for (int kb = 0; kb < k_blks; ++kb) {
    mbar_wait(r.full(stage), phase);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
      const uint64_t adesc = make_smem_desc_sw128(r.a(stage) + k * (UMMA_K * 2), 0, 1024);
      const uint64_t bdesc = make_smem_desc_sw128(r.b(stage) + k * (UMMA_K * 2), 0, 1024);
      mma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
    }
    mma_commit(r.empty(stage));  // frees the smem slot once these MMAs have read it
    if (++stage == r.stages) { stage = 0; phase ^= 1; }
  }
}

// ---------------------------------------------------------------------------
// Single-CTA kernel: the two small projections (row-remapping / fp32 epilogues) and the fp32 test hook.
// ---------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const TcGemmParams p) {
  using Cfg = GemmCfg<BN>;
  pdl_launch_dependents();
  MST_DBG_WALL(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  Ring ring{base, base + Cfg::STAGES * Cfg::STAGE_BYTES, Cfg::STAGES, Cfg::STAGE_BYTES, Cfg::A_BYTES};
  const uint32_t tmem_slot = ring.extra(0);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_w);
    for (int st = 0; st < Cfg::STAGES; ++st) {
      mbar_init(ring.full(st), 1);
      mbar_init(ring.empty(st), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(ring.tfull(i), 1);
      mbar_init(ring.tempty(i), NUM_EPI_THREADS);
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
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_blks = p.N / BN;
  const int m_blks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_tiles = m_blks * n_blks;
  const int k_blks = p.K / BLOCK_K;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blks, n_blk = tile - m_blk * n_blks;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(ring.empty(stage), phase ^ 1);
          mbar_expect_tx(ring.full(stage), Cfg::STAGE_BYTES);
          tma_load_2d(ring.a(stage), &tmap_a, ring.full(stage), kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(ring.b(stage), &tmap_w, ring.full(stage), kb * BLOCK_K, n_blk * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(ring.tempty(acc), acc_phase ^ 1);
        tc_fence_after();
        mma_tile<BN>(ring, tmem_base + (uint32_t)(acc * BN), k_blks, stage, phase, p.ab_f16 != 0);
        mma_commit(ring.tfull(acc));  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp is allowed to access
    const int half = ew >> 2;   // column half of the tile
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_blks, n_blk = tile - m_blk * n_blks;
      const int row = m_blk * BLOCK_M + row_in_tile;
      mbar_wait(ring.tfull(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        const int col = half * (BN / 2) + c;
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_addr + (uint32_t)(acc * BN + col), v);
        tmem_ld_wait();
        if constexpr (EPI == TC_EPI_INPROJ) {
          // acc + bias + positional row -> 16-bit -> the warp's swizzled [32 x 32] box -> read back 8 rows x 64 bytes per
          // instruction -> coalesced 16-byte stores to the token rows of every CFG pass (a row-per-thread store of the
          // same data wrote half sectors of 32 different lines per instruction: 29 us for a 4.6 GFLOP GEMM)
          const int n = n_blk * BN + col;
          const uint32_t box = ring.bar_base + Cfg::BAR_BYTES + (uint32_t)ew * Cfg::OBOX_BYTES;
          uint32_t o[16];
          {
            const int rb = row < p.M ? row / p.T : 0, rt = row < p.M ? row - rb * p.T : 0;
            const float* pe = p.pe + (size_t)(rt + 1) * p.N + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              const float4 e4 = __ldg(reinterpret_cast<const float4*>(pe + j));
              const float x0 = __uint_as_float(v[j]) + b4.x + e4.x, x1 = __uint_as_float(v[j + 1]) + b4.y + e4.y;
              const float x2 = __uint_as_float(v[j + 2]) + b4.z + e4.z, x3 = __uint_as_float(v[j + 3]) + b4.w + e4.w;
              o[j >> 1] = p.io_f16 ? pack_f16x2(x0, x1) : pack_bf16x2(x0, x1);
              o[(j >> 1) + 1] = p.io_f16 ? pack_f16x2(x2, x3) : pack_bf16x2(x2, x3);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(box + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
          __syncwarp();
          const int piece = lane & 3, S = p.T + 1;
          uint16_t* const out16 = static_cast<uint16_t*>(p.out);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = k * 8 + (lane >> 2);
            const uint4 val = lds128(box + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
            const int grow = m_blk * BLOCK_M + quad * 32 + rr;
            if (grow < p.M) {
              const int b = grow / p.T, t = grow - b * p.T;
              for (int pass = 0; pass < p.n_pass; ++pass)
                *reinterpret_cast<uint4*>(out16 + ((size_t)(pass * p.B + b) * S + t + 1) * p.ldo + n + piece * 8) = val;
            }
          }
          __syncwarp();
        } else {
          epilogue_chunk<EPI>(p, row, n_blk * BN + col, v);
        }
      }
      tc_fence_before();
      mbar_arrive(ring.tempty(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    MST_DBG_WALL_END();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    if (p.td_mode == 0) tc_fence_after();
    MST_DBG_WALL_W1(0);
    if (p.td_mode != 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    MST_DBG_WALL_W1(1);
  }
  if (p.td_mode == 2 && warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  __syncthreads();
  MST_DBG_WALL(2);
}

// ---------------------------------------------------------------------------
// CTA-pair GEMM (tcgen05 cta_group::2): one 256 x 256 output tile per cluster of two CTAs
// (QKV projection, FFN linear1 + GELU).  Each CTA stages its own 128 rows of A and HALF of the
// 256 weight rows, so it pulls 32 KB per k-block through L2 instead of 48 KB.  The leader CTA
// (rank 0) issues the MMAs for both; both CTAs run the epilogue of their 128 rows.
// Epilogue: every warp owns its 32 rows: TMEM -> registers -> (+bias, GELU) -> bf16 -> the warp's own
// 128B-swizzled [32 x 64] staging box -> one TMA store per box.  Two boxes per warp ping-pong, so the only
// synchronisation is __syncwarp and the bulk-group wait of lane 0 - no CTA-wide barrier in the loop.
// ---------------------------------------------------------------------------
template <int EPI, int BN_ = 256>
struct PairCfg {
  // Tile width (columns of the 256-row pair tile).  256 maximises operand re-use; a narrower tile trades it for less
  // wave quantisation (QKV at B=64 is 594 tiles of 256 columns on 74 clusters = 8.03 waves: two clusters run a ninth
  // tile while 72 idle, 12 % of the launch, profiles/r02h_ncu_full_step_kernels_summary.txt) - measured slower, see
  // launch_gemm_pair.
  static constexpr int BN = BN_;
  // The GELU epilogue is issue/latency-bound with two warps per scheduler (5.6k cycles per tile against 4.1k of MMAs):
  // it gets 16 epilogue warps (4 per scheduler) and pays with one ring stage; the plain-bias epilogue keeps 8 warps.
  static constexpr int EPI_WARPS = (EPI == TC_EPI_BIAS_GELU_BF16 && BN_ % 128 == 0) ? 16 : 8;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);
  static_assert(COLS_PER_WARP % 32 == 0, "a warp handles whole 32-column steps");
  static constexpr int A_BYTES = BLOCK_M * 128;       // 128 rows x 64 k
  static constexpr int B_BYTES = (BN / 2) * 128;      // this CTA's half of the BN weight rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OBOX_BYTES = 32 * 64;          // output staging: [32 rows x 32 bf16], 64B swizzle
  static constexpr int OUT_BYTES = EPI_WARPS * 2 * OBOX_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int MAX_SMEM = 232448;
  static constexpr int STAGES_FIT = (MAX_SMEM - 1024 - OUT_BYTES - BAR_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;  // 256: 6 (bias) / 5 (GELU); 192: 6; 128: 8 / 6
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES;
};


template <int EPI, int BN_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<EPI, BN_>::THREADS, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const TcGemmParams p) {
  using Cfg = PairCfg<EPI, BN_>;
  constexpr int BN = Cfg::BN;
  pdl_launch_dependents();
  MST_DBG_WALL(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t out_smem = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  Ring ring{base, out_smem + Cfg::OUT_BYTES, Cfg::STAGES, Cfg::STAGE_BYTES, Cfg::A_BYTES};
  const uint32_t tmem_slot = ring.extra(0);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_w);
    prefetch_tensormap(&tmap_out);
    for (int st = 0; st < Cfg::STAGES; ++st) {
      mbar_init(ring.full(st), 1);   // leader: its own arrive.expect_tx covers the bytes of both CTAs
      mbar_init(ring.empty(st), 1);  // multicast tcgen05.commit of the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(ring.tfull(i), 1);
      mbar_init(ring.tempty(i), 2 * Cfg::EPI_THREADS);  // leader: epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_blks = p.N / BN;
  const int m_blks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int m_pairs = (m_blks + 1) / 2;
  const int num_tiles = m_pairs * n_blks;
  const int k_blks = p.K / BLOCK_K;

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs; completion is signalled on the LEADER's full barrier) ----------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long* dbg = (p.dbg && cluster_id == 0) ? p.dbg + (0 * 2 + rank) * 1024 : nullptr;
      int di = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int mp = tile / n_blks, n_blk = tile - mp * n_blks;
        const int m_blk = 2 * mp + (int)rank;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(ring.empty(stage), phase ^ 1);
          MST_DBG_STAMP();
          const uint32_t full_leader = map_to_cta(ring.full(stage), 0);
          if (leader) mbar_expect_tx(ring.full(stage), 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2cta(ring.a(stage), &tmap_a, full_leader, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d_2cta(ring.b(stage), &tmap_w, full_leader, kb * BLOCK_K, n_blk * BN + (int)rank * (BN / 2));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: leader CTA only, one thread for the pair ----------------
    if (leader && elect_one()) {
      const uint32_t idesc = p.ab_f16 ? make_idesc_f16(2 * BLOCK_M, BN, 0) : make_idesc_bf16(2 * BLOCK_M, BN, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      long long* dbg = (p.dbg && cluster_id == 0) ? p.dbg + (1 * 2 + rank) * 1024 : nullptr;
      int di = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        MST_DBG_STAMP();
        mbar_wait_cluster(ring.tempty(acc), acc_phase ^ 1);
        tc_fence_after();
        MST_DBG_STAMP();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait_cluster(ring.full(stage), phase);
          tc_fence_after();
          MST_DBG_STAMP();
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(ring.a(stage) + k * (UMMA_K * 2), 0, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(ring.b(stage) + k * (UMMA_K * 2), 0, 1024);
            mma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          mma_commit_2cta(ring.empty(stage), 3);  // both CTAs' producers may refill the slot
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit_2cta(ring.tfull(acc), 3);  // both CTAs' epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------- epilogue (both CTAs, each for its own 128 rows) ----------------
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int cgrp = ew >> 2;  // column group of the tile: COLS_PER_WARP columns per warp
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t my_box = out_smem + ew * 2 * Cfg::OBOX_BYTES;  // this warp's two staging boxes (ping-pong)
    const uint32_t my_row = lane * 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long* dbg = (p.dbg && cluster_id == 0 && warp == 2 && lane == 0) ? p.dbg + (2 * 2 + rank) * 1024 : nullptr;
    int di = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int mp = tile / n_blks, n_blk = tile - mp * n_blks;
      const int m_blk = 2 * mp + (int)rank;
      MST_DBG_STAMP();
      mbar_wait(ring.tfull(acc), acc_phase);
      tc_fence_after();
      MST_DBG_STAMP();
      const uint32_t tempty_leader = map_to_cta(ring.tempty(acc), 0);
      constexpr int STEPS = Cfg::COLS_PER_WARP / 32;
      const uint32_t acc_addr = tmem_base + lane_addr + (uint32_t)(acc * BN + cgrp * Cfg::COLS_PER_WARP);
      // this warp's 32 rows x COLS_PER_WARP columns, 32 columns at a time; the next 32 are in flight while these are
      // processed
      uint32_t v[2][32];
      tmem_ld32(acc_addr, v[0]);
#pragma unroll
      for (int c4 = 0; c4 < STEPS; ++c4) {
        const int col = cgrp * Cfg::COLS_PER_WARP + c4 * 32;  // first tile column of this step
        // the box was handed to a TMA store two steps ago: at most the other box's store may still be reading
        if (lane == 0) bulk_wait_read_1();
        __syncwarp();
        tmem_ld_wait();
        if (c4 < STEPS - 1) tmem_ld32(acc_addr + (uint32_t)((c4 + 1) * 32), v[(c4 + 1) & 1]);
        if (c4 == STEPS - 1) {  // accumulator drained: the MMAs of the tile after next may overwrite it
          tc_fence_before();
          mbar_arrive_cluster_relaxed(tempty_leader);
          MST_DBG_STAMP();
        }
        const float* bias = p.bias + n_blk * BN + col;
        const uint32_t row_smem = my_box + (c4 & 1) * Cfg::OBOX_BYTES + my_row;
        const uint32_t* a = v[c4 & 1];
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // 8 columns -> one 16-byte piece of the 64-byte swizzled row
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 8 * q));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 8 * q + 4));
          uint64_t x[4];
          x[0] = add2(pk2u(a[8 * q + 0], a[8 * q + 1]), pk2(b0.x, b0.y));
          x[1] = add2(pk2u(a[8 * q + 2], a[8 * q + 3]), pk2(b0.z, b0.w));
          x[2] = add2(pk2u(a[8 * q + 4], a[8 * q + 5]), pk2(b1.x, b1.y));
          x[3] = add2(pk2u(a[8 * q + 6], a[8 * q + 7]), pk2(b1.z, b1.w));
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if constexpr (EPI == TC_EPI_BIAS_GELU_BF16) x[e] = gelu2(x[e]);
            float y0, y1;
            upk2(x[e], y0, y1);
            o[e] = pack_bf16x2(y0, y1);
          }
          sts128(row_smem + ((q ^ ((lane >> 1) & 3)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_out, my_box + (c4 & 1) * Cfg::OBOX_BYTES, n_blk * BN + col, m_blk * BLOCK_M + quad * 32, 0);
          bulk_commit_group();
        }
      }
      MST_DBG_STAMP();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_all();
    MST_DBG_WALL_END();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
  __syncthreads();
  MST_DBG_WALL(2);
}

// ---------------------------------------------------------------------------
// out = LayerNorm(A W^T + bias + residual) * gamma + beta,  N = 512, bf16 out
// (self_attn.out_proj + norm1, linear2 + norm2).
// Cluster of 4 CTAs per 256-row block: two cta_group::2 pairs, pair p = ranks {2p, 2p+1} owns columns
// [256p, 256p+256) of all 256 rows, so every CTA holds a 128 x 256 fp32 accumulator twice (two TMEM stages: the
// MMAs of the next block overlap the normalisation) and pulls only 32 KB per k-block through L2.
// Per epilogue warp: its 32 rows x 128 columns; the residual arrives by TMA in the warp's two [32 x 64] boxes,
// acc + bias + residual stays in registers (fp32), and the four partial (sum, sum of squares) of a row -
// 2 pairs x 2 column halves - meet through shared memory: local st.shared + mbarrier arrive, remote st.async with
// transaction bytes on the partner CTA's mbarrier (rank ^ 2 holds the same rows; no cluster-scope fence anywhere).
// ---------------------------------------------------------------------------
// two 16-bit floats (packed in 32 bits) -> two fp32 (packed in 64 bits)
template <bool F16>
__device__ __forceinline__ uint64_t h2_to_f32x2(uint32_t u) {
  if constexpr (F16) {
    const float2 f = unpack_f16x2(u);
    return pk2(f.x, f.y);
  } else {
    return bf16x2_to_f32x2(u);
  }
}
template <bool F16>
__device__ __forceinline__ uint32_t f32x2_to_h2(float lo, float hi) {
  if constexpr (F16) return pack_f16x2(lo, hi);
  else return pack_bf16x2(lo, hi);
}

constexpr int LN_LEAD = 2;  // k-blocks of the next tile that enter the ring before a tile's residual (see the producer)

struct LnCfg {
  static constexpr int BN = 256;
  static constexpr int STAGES = 5;  // ring slots shared by the k-blocks AND the two residual blocks of every tile
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EXCH_BYTES = 2 * 4 * BLOCK_M * 8;      // [parity][source = 2*pair + half][row] float2
  static constexpr int PARAM_BYTES = 3 * BN * 4;              // bias | gamma | beta of this CTA's 256 columns
  static constexpr int OBOX_BYTES = 32 * 64;                  // output staging: [32 rows x 32 bf16], 64B swizzle
  static constexpr int OUT_BYTES = NUM_EPI_WARPS * 2 * OBOX_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_BYTES + EXCH_BYTES + PARAM_BYTES + BAR_BYTES;
};

template <bool STREAM_F16>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                  const TcGemmParams p) {
  using Cfg = LnCfg;
  constexpr int BN = Cfg::BN;
  pdl_launch_dependents();
  MST_DBG_WALL(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t out_smem = base + Cfg::STAGES * Cfg::STAGE_BYTES;  // per-warp output boxes
  const uint32_t exch_smem = out_smem + Cfg::OUT_BYTES;
  const uint32_t par_smem = exch_smem + Cfg::EXCH_BYTES;
  Ring ring{base, par_smem + Cfg::PARAM_BYTES, Cfg::STAGES, Cfg::STAGE_BYTES, Cfg::A_BYTES};
  const uint32_t tmem_slot = ring.extra(0);
  // Row-statistics barriers, one per tile parity: 256 local arrivals + 2048 transaction bytes from the partner.
  // Two barriers (not one with two phases): the partner's st.async bytes of tile i+1 may be issued while a slow thread
  // of this CTA has not yet arrived for tile i; on a single barrier they would be booked on tile i's transaction
  // count and that phase could never complete (seen as a once-in-10^8-tiles cluster hang at B=2048).
  auto stats_bar = [&](uint32_t par_) { return par_ ? ring.extra(6) : ring.extra(1); };
  // The residual tile of a block (128 rows x 256 columns = two ring slots of [128 x 64 | 128 x 64]) travels through the
  // SAME ring as the k-blocks, right behind the block's last k-block: it is in flight while the MMAs finish and costs
  // no dedicated staging memory.  Slot h of the two holds the columns of epilogue half h.
  auto res_full = [&](int h) { return ring.extra(2 + h); };  // residual slot h has landed (local CTA, tx bytes)
  auto res_free = [&](int h) { return ring.extra(4 + h); };  // ... and has been read by the 128 threads of half h
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  const float2* exch_ptr = reinterpret_cast<const float2*>(base_ptr + (exch_smem - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1, mrank = rank & 1;  // column half / row half of the 256 x 512 block
  const uint32_t leader_rank = rank & ~1u;
  const bool leader = mrank == 0;
  const int cluster_id = blockIdx.x >> 2, n_clusters = gridDim.x >> 2;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_w);
    prefetch_tensormap(&tmap_res);
    prefetch_tensormap(&tmap_out);
    for (int st = 0; st < Cfg::STAGES; ++st) {
      mbar_init(ring.full(st), 1);
      mbar_init(ring.empty(st), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(ring.tfull(i), 1);
      mbar_init(ring.tempty(i), 2 * NUM_EPI_THREADS);
    }
    mbar_init(stats_bar(0), NUM_EPI_THREADS);
    mbar_init(stats_bar(1), NUM_EPI_THREADS);
    for (int h = 0; h < 2; ++h) {
      mbar_init(res_full(h), 1);
      mbar_init(res_free(h), NUM_EPI_THREADS / 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  if (threadIdx.x < BN) {  // the epilogue reads its per-column parameters from shared memory (broadcast LDS.128)
    float* par = reinterpret_cast<float*>(base_ptr + (par_smem - base));
    const int c = (int)pair * BN + (int)threadIdx.x;
    par[threadIdx.x] = p.bias[c];
    par[BN + threadIdx.x] = p.ln_g[c];
    par[2 * BN + threadIdx.x] = p.ln_b[c];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before anything is signalled on them remotely
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m_blks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int m_pairs = (m_blks + 1) / 2;
  const int k_blks = p.K / BLOCK_K;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int owner[Cfg::STAGES] = {};  // h+1 while the slot holds residual half h that has not been waited free yet
      int res_loads = 0;            // tiles whose residual has been issued
      auto acquire = [&]() {        // the slot is free for a new load
        if (owner[stage]) {
          // its last occupant was a residual block: the epilogue (not the MMA) consumed it; this thread completes the
          // slot's "empty" phase itself so that the ring's phase bookkeeping stays uniform
          mbar_wait(res_free(owner[stage] - 1), (uint32_t)((res_loads - 1) & 1));
          mbar_arrive(ring.empty(stage));
          owner[stage] = 0;
        }
        mbar_wait(ring.empty(stage), phase ^ 1);
      };
      auto load_kb = [&](int m_blk, int kb) {
        acquire();
        const uint32_t full_leader = map_to_cta(ring.full(stage), leader_rank);
        if (leader) mbar_expect_tx(ring.full(stage), 2 * Cfg::STAGE_BYTES);
        tma_load_2d_2cta(ring.a(stage), &tmap_a, full_leader, kb * BLOCK_K, m_blk * BLOCK_M);
        tma_load_2d_2cta(ring.b(stage), &tmap_w, full_leader, kb * BLOCK_K, (int)pair * BN + (int)mrank * (BN / 2));
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      };
      auto load_res = [&](int m_blk) {
        for (int h = 0; h < 2; ++h) {  // residual columns [256*pair + 128*h, +128) of this CTA's 128 rows
          acquire();
          mbar_expect_tx(res_full(h), Cfg::STAGE_BYTES);
          tma_load_2d(ring.a(stage), &tmap_res, res_full(h), (int)pair * BN + h * 128, m_blk * BLOCK_M);
          tma_load_2d(ring.b(stage), &tmap_res, res_full(h), (int)pair * BN + h * 128 + 64, m_blk * BLOCK_M);
          owner[stage] = h + 1;
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ++res_loads;
      };
      // Ring order: the residual of tile i goes in BEHIND the first LN_LEAD k-blocks of tile i+1 (directly behind its own
      // k-blocks for the last tile).  Loaded right behind tile i's last k-block it parks two of the five slots from ~3
      // k-blocks before the MMAs of tile i end until pass 1 of its epilogue is done, and the MMAs of tile i+1 start on a
      // three-slot ring with nothing prefetched: a 1.3 k cycle wait at k-block 0 and 1.7 k at k-block 3 of every tile
      // (profiles/r02o_ln_v4_probe_no_smem_epilogue.txt).  With the lead the wait at k-block 0 is gone and a shorter one
      // appears at k-block 2 (the two slots are still parked): 19.0 -> 18.7 us (K = 512), 28.6 -> 27.9 us (K = 1024); leads
      // of 1..5 all measure the same (profiles/r02o_ln_v4_residual_lead.txt).  MST_TEARDOWN=8 restores the old order.
      const int lead = (p.td_mode == 8 || k_blks <= LN_LEAD) ? 0 : LN_LEAD;
      int prev_blk = -1;
      for (int mp = cluster_id; mp < m_pairs; mp += n_clusters) {
        const int m_blk = 2 * mp + (int)mrank;
        for (int kb = 0; kb < k_blks; ++kb) {
          if (kb == lead && prev_blk >= 0) load_res(prev_blk);
          load_kb(m_blk, kb);
        }
        prev_blk = m_blk;
      }
      if (prev_blk >= 0) load_res(prev_blk);
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BN, 0);
      const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
      const int lead = (p.td_mode == 8 || k_blks <= LN_LEAD) ? 0 : LN_LEAD;
      int stage = 0, acc = 0;
      uint32_t full_phase = 0, acc_phase = 0;  // bit s of full_phase: parity the next k-block in slot s completes
      long long* dbg = (p.dbg && cluster_id == 0 && rank == 0) ? p.dbg + (1 * 2 + 0) * 1024 : nullptr;
      int di = 0;
      for (int mp = cluster_id; mp < m_pairs; mp += n_clusters) {
        MST_DBG_STAMP();
        mbar_wait_cluster(ring.tempty(acc), acc_phase ^ 1);
        tc_fence_after();
        MST_DBG_STAMP();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blks; ++kb) {
          if (kb == lead && mp != cluster_id)
            for (int h = 0; h < 2; ++h)  // the two residual slots of the previous tile: nothing to multiply
              if (++stage == Cfg::STAGES) stage = 0;
          mbar_wait_cluster(ring.full(stage), (full_phase >> stage) & 1u);
          full_phase ^= 1u << stage;
          tc_fence_after();
          MST_DBG_STAMP();
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(ring.a(stage) + k * (UMMA_K * 2), 0, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(ring.b(stage) + k * (UMMA_K * 2), 0, 1024);
            mma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          mma_commit_2cta(ring.empty(stage), pair_mask);
          if (++stage == Cfg::STAGES) stage = 0;
        }
        mma_commit_2cta(ring.tfull(acc), pair_mask);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;  // 128-column half of this CTA's 256 columns
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int col0 = (int)pair * BN + half * 128;  // first global column of this thread's 128
    const uint32_t my_obox = out_smem + ew * 2 * Cfg::OBOX_BYTES;  // output: two [32 x 32] boxes, ping-pong
    const uint32_t partner = rank ^ 2u;  // the CTA holding the other 256 columns of the same rows
    const uint32_t peer_exch = map_to_cta(exch_smem, partner);
    const uint32_t peer_stats_bar0 = map_to_cta(stats_bar(0), partner), peer_stats_bar1 = map_to_cta(stats_bar(1), partner);
    const int my_src = (int)pair * 2 + half;
    const uint32_t my_row = (uint32_t)r * 128;  // this thread's row inside a [128 x 64] residual box
    int acc = 0, it = 0;
    uint32_t acc_phase = 0;
    long long* dbg = (p.dbg && cluster_id == 0 && warp == 2 && lane == 0 && rank < 2) ? p.dbg + (2 * 2 + rank) * 1024 : nullptr;
    int di = 0;
    for (int mp = cluster_id; mp < m_pairs; mp += n_clusters, ++it) {
      const int m_blk = 2 * mp + (int)mrank;
      const uint32_t par = it & 1;
      MST_DBG_STAMP();

      // ring slot that holds this half's residual: the tile's k-blocks come first, then residual half 0, half 1
      // ring entries before this tile's residual: the k-blocks of tiles 0..it, the residuals of tiles 0..it-1 and - unless
      // this is the cluster's last tile - the first `lead` k-blocks of tile it+1
      const bool last_tile = mp + n_clusters >= m_pairs;
      const int lead = (p.td_mode == 8 || k_blks <= LN_LEAD) ? 0 : LN_LEAD;
      const uint32_t res_smem = ring.a(((it + 1) * k_blks + 2 * it + (last_tile ? 0 : lead) + half) % Cfg::STAGES);
      // The residual goes to registers BEFORE the accumulator is awaited (the epilogue warps are idle here while the MMAs of
      // the tile finish, and the 64 registers are free until pass 1 builds x): its two ring slots go back to the producer
      // ~2 k cycles earlier than when pass 1 read them from shared memory.
      mbar_wait(res_full(half), par);
      uint4 rres[16];
#pragma unroll
      for (int c = 0; c < 16; ++c)  // 16-byte piece c of this thread's 128 residual columns: box c / 8, piece c % 8 of the row
        rres[c] = lds128(res_smem + (c >> 3) * Cfg::A_BYTES + my_row + (((c & 7) ^ (r & 7)) << 4));
      mbar_arrive(res_free(half));  // the residual slot may be refilled
      MST_DBG_STAMP();
      mbar_wait(ring.tfull(acc), acc_phase);
      tc_fence_after();
      MST_DBG_STAMP();
      const uint32_t tempty_leader = map_to_cta(ring.tempty(acc), leader_rank);
      // pass 1: x = acc + bias + residual (kept in registers as packed fp32 pairs), row sum and sum of squares on
      // four independent packed accumulators (fma.rn.f32x2: two elements per instruction)
      uint64_t x2[64];
      uint64_t s2a = 0, s2b = 0, q2a = 0, q2b = 0;
      const uint32_t bias_smem = par_smem + (uint32_t)(half * 128) * 4;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {  // 32 accumulator columns at a time
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_addr + (uint32_t)(acc * BN + half * 128 + c4 * 32), v);
        uint4 rr[4], bb[8];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = (c4 & 1) * 4 + jj;  // 16-byte piece of the 128-byte staged row
          rr[jj] = rres[(c4 >> 1) * 8 + j];
          bb[2 * jj] = lds128(bias_smem + (uint32_t)(c4 * 32 + jj * 8) * 4);
          bb[2 * jj + 1] = lds128(bias_smem + (uint32_t)(c4 * 32 + jj * 8 + 4) * 4);
        }
        tmem_ld_wait();
        if (c4 == 3) {  // accumulator is in registers: release the TMEM stage before the normalisation
          tc_fence_before();
          mbar_arrive_cluster_relaxed(tempty_leader);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t* a = &v[jj * 8];
          uint64_t* xo = &x2[c4 * 16 + jj * 4];
          xo[0] = add2(add2(pk2u(a[0], a[1]), pk2u(bb[2 * jj].x, bb[2 * jj].y)), h2_to_f32x2<STREAM_F16>(rr[jj].x));
          xo[1] = add2(add2(pk2u(a[2], a[3]), pk2u(bb[2 * jj].z, bb[2 * jj].w)), h2_to_f32x2<STREAM_F16>(rr[jj].y));
          xo[2] = add2(add2(pk2u(a[4], a[5]), pk2u(bb[2 * jj + 1].x, bb[2 * jj + 1].y)), h2_to_f32x2<STREAM_F16>(rr[jj].z));
          xo[3] = add2(add2(pk2u(a[6], a[7]), pk2u(bb[2 * jj + 1].z, bb[2 * jj + 1].w)), h2_to_f32x2<STREAM_F16>(rr[jj].w));
          s2a = add2(s2a, xo[0]); q2a = fma2(xo[0], xo[0], q2a);
          s2b = add2(s2b, xo[1]); q2b = fma2(xo[1], xo[1], q2b);
          s2a = add2(s2a, xo[2]); q2a = fma2(xo[2], xo[2], q2a);
          s2b = add2(s2b, xo[3]); q2b = fma2(xo[3], xo[3], q2b);
        }
      }
      float sum, sq;
      {
        float a0, a1, b0, b1;
        upk2(add2(s2a, s2b), a0, a1);
        upk2(add2(q2a, q2b), b0, b1);
        sum = a0 + a1;
        sq = b0 + b1;
      }
      MST_DBG_STAMP();
      // row statistics: 4 partials per row (2 pairs x 2 halves)
      const uint32_t slot = (uint32_t)(((par * 4 + my_src) * BLOCK_M + r) * 8);
      const uint32_t sbar = stats_bar(par);
      st_async_f32x2(peer_exch + slot, sum, sq, par ? peer_stats_bar1 : peer_stats_bar0);
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(exch_smem + slot), "f"(sum), "f"(sq) : "memory");
      if (ew == 0 && lane == 0)
        mbar_expect_tx(sbar, NUM_EPI_THREADS * 8);  // arrive + the partner's 256 x 8 bytes of this tile
      else
        mbar_arrive(sbar);
      mbar_wait_cluster(sbar, (uint32_t)(it >> 1) & 1u);
      MST_DBG_STAMP();
      float tsum = 0.0f, tsq = 0.0f;
#pragma unroll
      for (int src = 0; src < 4; ++src) {
        const float2 e = exch_ptr[(par * 4 + src) * BLOCK_M + r];
        tsum += e.x;
        tsq += e.y;
      }
      const float mean = tsum * (1.0f / LN_N);
      const float var = fmaxf(tsq * (1.0f / LN_N) - mean * mean, 0.0f);
      const float rstd = rsqrtf(var + 1e-5f);
      // pass 2: y = x * (rstd*g) + (b - mean*rstd*g), packed; 32 columns at a time through the two output boxes
      const uint64_t rstd2 = pk2(rstd, rstd), nmean2 = pk2(-mean, -mean);
      const uint32_t g_smem = par_smem + (uint32_t)(BN + half * 128) * 4, b_smem = par_smem + (uint32_t)(2 * BN + half * 128) * 4;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t box = my_obox + (c4 & 1) * Cfg::OBOX_BYTES;
        if (lane == 0) bulk_wait_read_1();  // the store that last used this box has read it
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int col = c4 * 32 + jj * 8;
          const uint4 g0 = lds128(g_smem + (uint32_t)col * 4), g1 = lds128(g_smem + (uint32_t)(col + 4) * 4);
          const uint4 t0 = lds128(b_smem + (uint32_t)col * 4), t1 = lds128(b_smem + (uint32_t)(col + 4) * 4);
          const uint64_t* xi = &x2[c4 * 16 + jj * 4];
          const uint64_t gg[4] = {pk2u(g0.x, g0.y), pk2u(g0.z, g0.w), pk2u(g1.x, g1.y), pk2u(g1.z, g1.w)};
          const uint64_t tt[4] = {pk2u(t0.x, t0.y), pk2u(t0.z, t0.w), pk2u(t1.x, t1.y), pk2u(t1.z, t1.w)};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t a = mul2(gg[q], rstd2);
            const uint64_t y = fma2(xi[q], a, fma2(a, nmean2, tt[q]));
            float y0, y1;
            upk2(y, y0, y1);
            o[q] = f32x2_to_h2<STREAM_F16>(y0, y1);
          }
          sts128(box + lane * 64 + ((jj ^ ((lane >> 1) & 3)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_out, box, col0 + c4 * 32, m_blk * BLOCK_M + quad * 32, 0);
          bulk_commit_group();
        }
      }
      MST_DBG_STAMP();
      __syncwarp();
      MST_DBG_STAMP();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_all();
    MST_DBG_WALL_END();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits while a partner may still write statistics into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
  __syncthreads();
  MST_DBG_WALL(2);
}

// ---------------------------------------------------------------------------
// LN GEMM, version 5 (default; MST_LN_V=4 selects the kernel above).  Same cluster / MMA structure, but:
//  * the k-block ring carries ONLY k-blocks (4 stages).  In v4 the residual tile travelled through the ring and
//    its two slots stayed occupied until the epilogue had read them, so the producer could run no more than three
//    k-blocks ahead into the next tile: MMAs and epilogue ran back to back instead of overlapped
//    (profiles/r02a_timeline_ln_v4.txt: 6.5 k cycles per 8-k-block tile against 4.1 k of MMAs).
//  * every epilogue warp owns an 8 KB region (its 32 rows x 128 columns as two 128B-swizzled [32 x 64] boxes).  The
//    warp loads its own residual rows into it by TMA (own mbarrier; issued for tile i+1 as soon as the warp is done
//    with tile i), reads it in pass 1 and re-uses it in pass 2 as the transpose scratch of its output:
//    16-bit rows are written in the swizzled layout, read back 4 rows x 128 bytes per instruction and stored with
//    plain coalesced 16-byte st.global (full 128-byte lines).  No TMA store, no bulk-group waits in the loop: the
//    staged TMA stores of v4 cost ~800 cycles per 32-column chunk (3.2 k of the 6.1 k cycle epilogue).
//  * STREAM_F16: the residual read and the normalised output are IEEE fp16 instead of bf16.  LayerNorm outputs are
//    bounded (|y| <= |gamma| sqrt(d) + |beta|), so the 5-bit exponent is safe, and the 10-bit mantissa removes the
//    dominant error of the bf16 path (the residual stream rounded 16 times per forward: 1.7e-2 -> 4.6e-3 per-step x0
//    error under guidance scale 2.5; tools/rounding_experiment.py).
// ---------------------------------------------------------------------------
struct Ln5Cfg {
  static constexpr int BN = 256;
  static constexpr int STAGES = 4;
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RES_WARP_BYTES = 2 * 32 * 128;               // two [32 rows x 64 cols] boxes per epilogue warp
  static constexpr int RES_BYTES = NUM_EPI_WARPS * RES_WARP_BYTES;  // 64 KB
  static constexpr int EXCH_BYTES = 2 * 4 * BLOCK_M * 8;
  static constexpr int PARAM_BYTES = 3 * BN * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + RES_BYTES + EXCH_BYTES + PARAM_BYTES + BAR_BYTES;
};

__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void stg128(void* ptr, uint4 v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <bool STREAM_F16>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_ln5_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_res_pf,
                   const TcGemmParams p) {
  using Cfg = Ln5Cfg;
  constexpr int BN = Cfg::BN;
  pdl_launch_dependents();
  MST_DBG_WALL(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t res_smem = base + Cfg::STAGES * Cfg::STAGE_BYTES;  // per-warp residual / output-transpose regions
  const uint32_t exch_smem = res_smem + Cfg::RES_BYTES;
  const uint32_t par_smem = exch_smem + Cfg::EXCH_BYTES;
  Ring ring{base, par_smem + Cfg::PARAM_BYTES, Cfg::STAGES, Cfg::STAGE_BYTES, Cfg::A_BYTES};
  const uint32_t tmem_slot = ring.extra(0);
  auto stats_bar = [&](uint32_t par_) { return ring.extra(1 + par_); };  // one per tile parity (see v4)
  auto res_bar = [&](int ew_) { return ring.extra(3 + ew_); };           // residual of warp ew_ has landed
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  const float2* exch_ptr = reinterpret_cast<const float2*>(base_ptr + (exch_smem - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1, mrank = rank & 1;  // column half / row half of the 256 x 512 block
  const uint32_t leader_rank = rank & ~1u;
  const bool leader = mrank == 0;
  const int cluster_id = blockIdx.x >> 2, n_clusters = gridDim.x >> 2;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_w);
    prefetch_tensormap(&tmap_res);
    prefetch_tensormap(&tmap_res_pf);
    for (int st = 0; st < Cfg::STAGES; ++st) {
      mbar_init(ring.full(st), 1);
      mbar_init(ring.empty(st), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(ring.tfull(i), 1);
      mbar_init(ring.tempty(i), 2 * NUM_EPI_THREADS);
    }
    mbar_init(stats_bar(0), NUM_EPI_THREADS);
    mbar_init(stats_bar(1), NUM_EPI_THREADS);
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(res_bar(w), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  if (threadIdx.x < BN) {
    float* par = reinterpret_cast<float*>(base_ptr + (par_smem - base));
    const int c = (int)pair * BN + (int)threadIdx.x;
    par[threadIdx.x] = p.bias[c];
    par[BN + threadIdx.x] = p.ln_g[c];
    par[2 * BN + threadIdx.x] = p.ln_b[c];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m_blks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int m_pairs = (m_blks + 1) / 2;
  const int k_blks = p.K / BLOCK_K;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int mp = cluster_id; mp < m_pairs; mp += n_clusters) {
        const int m_blk = 2 * mp + (int)mrank;
        // pull the NEXT block's residual rows towards L2 while this block's operands stream in
        if (mp + n_clusters < m_pairs) {
          const int m_next = 2 * (mp + n_clusters) + (int)mrank;
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_prefetch_l2_2d(&tmap_res_pf, (int)pair * BN + c * 64, m_next * BLOCK_M);
        }
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(ring.empty(stage), phase ^ 1);
          const uint32_t full_leader = map_to_cta(ring.full(stage), leader_rank);
          if (leader) mbar_expect_tx(ring.full(stage), 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2cta(ring.a(stage), &tmap_a, full_leader, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d_2cta(ring.b(stage), &tmap_w, full_leader, kb * BLOCK_K, (int)pair * BN + (int)mrank * (BN / 2));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BN, 0);
      const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      long long* dbg = (p.dbg && cluster_id == 0 && rank == 0) ? p.dbg + (1 * 2 + 0) * 1024 : nullptr;
      int di = 0;
      for (int mp = cluster_id; mp < m_pairs; mp += n_clusters) {
        MST_DBG_STAMP();
        mbar_wait_cluster(ring.tempty(acc), acc_phase ^ 1);
        tc_fence_after();
        MST_DBG_STAMP();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait_cluster(ring.full(stage), phase);
          tc_fence_after();
          MST_DBG_STAMP();
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(ring.a(stage) + k * (UMMA_K * 2), 0, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(ring.b(stage) + k * (UMMA_K * 2), 0, 1024);
            mma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          mma_commit_2cta(ring.empty(stage), pair_mask);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit_2cta(ring.tfull(acc), pair_mask);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;  // 128-column half of this CTA's 256 columns
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int col0 = (int)pair * BN + half * 128;  // first global column of this thread's 128
    const uint32_t my_reg = res_smem + (uint32_t)ew * Cfg::RES_WARP_BYTES;  // this warp's residual / transpose region
    const uint32_t my_res_bar = res_bar(ew);
    const uint32_t partner = rank ^ 2u;
    const uint32_t peer_exch = map_to_cta(exch_smem, partner);
    const uint32_t peer_stats_bar0 = map_to_cta(stats_bar(0), partner), peer_stats_bar1 = map_to_cta(stats_bar(1), partner);
    const int my_src = (int)pair * 2 + half;
    const uint32_t my_row = (uint32_t)lane * 128;  // this thread's row inside a [32 x 64] box
    uint16_t* const out16 = static_cast<uint16_t*>(p.out);
    int acc = 0, it = 0;
    uint32_t acc_phase = 0;
    long long* dbg = (p.dbg && cluster_id == 0 && warp == 2 && lane == 0 && rank < 2) ? p.dbg + (2 * 2 + rank) * 1024 : nullptr;
    int di = 0;
    auto load_residual = [&](int mp_) {  // lane 0: this warp's 32 rows x 128 columns of block mp_
      const int row0 = (2 * mp_ + (int)mrank) * BLOCK_M + quad * 32;
      mbar_expect_tx(my_res_bar, Cfg::RES_WARP_BYTES);
      tma_load_2d(my_reg, &tmap_res, my_res_bar, col0, row0);
      tma_load_2d(my_reg + 4096, &tmap_res, my_res_bar, col0 + 64, row0);
    };
    if (lane == 0 && cluster_id < m_pairs) load_residual(cluster_id);
    for (int mp = cluster_id; mp < m_pairs; mp += n_clusters, ++it) {
      const int m_blk = 2 * mp + (int)mrank;
      const uint32_t par = it & 1;
      MST_DBG_STAMP();
      mbar_wait(ring.tfull(acc), acc_phase);
      tc_fence_after();
      MST_DBG_STAMP();
      mbar_wait(my_res_bar, par);
      MST_DBG_STAMP();
      const uint32_t tempty_leader = map_to_cta(ring.tempty(acc), leader_rank);
      // pass 1: x = acc + bias + residual (registers, packed fp32 pairs), row sum and sum of squares
      uint64_t x2[64];
      uint64_t s2a = 0, s2b = 0, q2a = 0, q2b = 0;
      const uint32_t bias_smem = par_smem + (uint32_t)(half * 128) * 4;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {  // 32 accumulator columns at a time
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_addr + (uint32_t)(acc * BN + half * 128 + c4 * 32), v);
        const uint32_t row_smem = my_reg + (uint32_t)(c4 >> 1) * 4096 + my_row;  // columns 0-63 | 64-127 of the half
        uint4 rr[4], bb[8];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = (c4 & 1) * 4 + jj;  // 16-byte piece of the 128-byte staged row
          rr[jj] = lds128(row_smem + ((j ^ (lane & 7)) << 4));
          bb[2 * jj] = lds128(bias_smem + (uint32_t)(c4 * 32 + jj * 8) * 4);
          bb[2 * jj + 1] = lds128(bias_smem + (uint32_t)(c4 * 32 + jj * 8 + 4) * 4);
        }
        tmem_ld_wait();
        if (c4 == 3) {  // accumulator is in registers: release the TMEM stage before the normalisation
          tc_fence_before();
          mbar_arrive_cluster_relaxed(tempty_leader);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t* a = &v[jj * 8];
          uint64_t* xo = &x2[c4 * 16 + jj * 4];
          xo[0] = add2(add2(pk2u(a[0], a[1]), pk2u(bb[2 * jj].x, bb[2 * jj].y)), h2_to_f32x2<STREAM_F16>(rr[jj].x));
          xo[1] = add2(add2(pk2u(a[2], a[3]), pk2u(bb[2 * jj].z, bb[2 * jj].w)), h2_to_f32x2<STREAM_F16>(rr[jj].y));
          xo[2] = add2(add2(pk2u(a[4], a[5]), pk2u(bb[2 * jj + 1].x, bb[2 * jj + 1].y)), h2_to_f32x2<STREAM_F16>(rr[jj].z));
          xo[3] = add2(add2(pk2u(a[6], a[7]), pk2u(bb[2 * jj + 1].z, bb[2 * jj + 1].w)), h2_to_f32x2<STREAM_F16>(rr[jj].w));
          s2a = add2(s2a, xo[0]); q2a = fma2(xo[0], xo[0], q2a);
          s2b = add2(s2b, xo[1]); q2b = fma2(xo[1], xo[1], q2b);
          s2a = add2(s2a, xo[2]); q2a = fma2(xo[2], xo[2], q2a);
          s2b = add2(s2b, xo[3]); q2b = fma2(xo[3], xo[3], q2b);
        }
      }
      float sum, sq;
      {
        float a0, a1, b0, b1;
        upk2(add2(s2a, s2b), a0, a1);
        upk2(add2(q2a, q2b), b0, b1);
        sum = a0 + a1;
        sq = b0 + b1;
      }
      MST_DBG_STAMP();
      // row statistics: 4 partials per row (2 pairs x 2 halves)
      const uint32_t slot = (uint32_t)(((par * 4 + my_src) * BLOCK_M + r) * 8);
      const uint32_t sbar = stats_bar(par);
      st_async_f32x2(peer_exch + slot, sum, sq, par ? peer_stats_bar1 : peer_stats_bar0);
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(exch_smem + slot), "f"(sum), "f"(sq) : "memory");
      if (ew == 0 && lane == 0)
        mbar_expect_tx(sbar, NUM_EPI_THREADS * 8);  // arrive + the partner's 256 x 8 bytes of this tile
      else
        mbar_arrive(sbar);
      mbar_wait_cluster(sbar, (uint32_t)(it >> 1) & 1u);
      MST_DBG_STAMP();
      float tsum = 0.0f, tsq = 0.0f;
#pragma unroll
      for (int src = 0; src < 4; ++src) {
        const float2 e = exch_ptr[(par * 4 + src) * BLOCK_M + r];
        tsum += e.x;
        tsq += e.y;
      }
      const float mean = tsum * (1.0f / LN_N);
      const float var = fmaxf(tsq * (1.0f / LN_N) - mean * mean, 0.0f);
      const float rstd = rsqrtf(var + 1e-5f);
      // pass 2: y = x * (rstd*g) + (b - mean*rstd*g); 64 columns at a time: own swizzled row of the box, then the box
      // leaves as 8 coalesced stores of 4 rows x 128 bytes
      const uint64_t rstd2 = pk2(rstd, rstd), nmean2 = pk2(-mean, -mean);
      const uint32_t g_smem = par_smem + (uint32_t)(BN + half * 128) * 4, b_smem = par_smem + (uint32_t)(2 * BN + half * 128) * 4;
      __syncwarp();  // every lane has read its residual row: the region becomes the output scratch
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t box = my_reg + (uint32_t)c * 4096;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int col = c * 64 + jj * 8;
          const uint4 g0 = lds128(g_smem + (uint32_t)col * 4), g1 = lds128(g_smem + (uint32_t)(col + 4) * 4);
          const uint4 t0 = lds128(b_smem + (uint32_t)col * 4), t1 = lds128(b_smem + (uint32_t)(col + 4) * 4);
          const uint64_t* xi = &x2[c * 32 + jj * 4];
          const uint64_t gg[4] = {pk2u(g0.x, g0.y), pk2u(g0.z, g0.w), pk2u(g1.x, g1.y), pk2u(g1.z, g1.w)};
          const uint64_t tt[4] = {pk2u(t0.x, t0.y), pk2u(t0.z, t0.w), pk2u(t1.x, t1.y), pk2u(t1.z, t1.w)};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t a = mul2(gg[q], rstd2);
            const uint64_t y = fma2(xi[q], a, fma2(a, nmean2, tt[q]));
            float y0, y1;
            upk2(y, y0, y1);
            o[q] = f32x2_to_h2<STREAM_F16>(y0, y1);
          }
          sts128(box + my_row + ((jj ^ (lane & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
        __syncwarp();
        const int prow = lane >> 3, piece = lane & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr_ = k * 4 + prow;  // row of the box
          const uint4 val = lds128(box + (uint32_t)rr_ * 128 + ((piece ^ (rr_ & 7)) << 4));
          const int grow = m_blk * BLOCK_M + quad * 32 + rr_;
          if (grow < p.M) stg128(out16 + (size_t)grow * LN_N + col0 + c * 64 + piece * 8, val);
        }
      }
      MST_DBG_STAMP();
      __syncwarp();  // the region has been read back: it may receive the next block's residual
      if (lane == 0 && mp + n_clusters < m_pairs) {
        fence_proxy_async_smem();  // generic-proxy accesses above -> before the async-proxy (TMA) write
        load_residual(mp + n_clusters);
      }
      MST_DBG_STAMP();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    MST_DBG_WALL_END();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits while a partner may still write statistics into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
  __syncthreads();
  MST_DBG_WALL(2);
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
    if (cache.size() > 8192) cache.clear();  // training tapes come and go: bound the cache
    cache[key] = m;
  }
  *out = m;
  return MST_OK;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t n_outer, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint64_t outer_stride_elems, uint32_t box_rows, uint32_t box_cols,
                      int swizzle_bytes) {
  static std::mutex mu;
  struct Key3 {
    const void* base;
    uint64_t v[5];
    uint32_t b[3];
    bool operator==(const Key3& o) const {
      return base == o.base && !memcmp(v, o.v, sizeof(v)) && !memcmp(b, o.b, sizeof(b));
    }
  };
  static std::vector<std::pair<Key3, CUtensorMap>> cache;  // a handful of entries per process
  Key3 key{base, {n_outer, rows, cols, row_stride_elems, outer_stride_elems}, {box_rows, box_cols, (uint32_t)swizzle_bytes}};
  {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : cache)
      if (kv.first == key) {
        *out = kv.second;
        return MST_OK;
      }
  }
  EncodeTiledFn enc = get_encoder();
  if (!enc) return fail(MST_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * 2) % 16 != 0 || (outer_stride_elems * 2) % 16 != 0)
    return fail(MST_ERR_INVALID, "make_tmap_bf16_3d: base and strides must be 16-byte aligned");
  if (!((swizzle_bytes == 128 && box_cols == 64) || (swizzle_bytes == 64 && box_cols == 32)))
    return fail(MST_ERR_INVALID, "make_tmap_bf16_3d: box width must equal the swizzle span");
  cuuint64_t dims[3] = {cols, rows, n_outer};
  cuuint64_t strides[2] = {row_stride_elems * 2, outer_stride_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MST_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string((int)r));
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 256) cache.clear();
    cache.emplace_back(key, m);
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
  static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
  if (attr_set.first()) {
    MST_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int tiles = ceil_div(p.M, BLOCK_M) * (p.N / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  MST_CUDA_OK(launch_pdl(tc_gemm_kernel<BN, EPI>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, ta, tw, p));
  static const char* const kNames[] = {"tc_gemm_qkv", "tc_gemm_ffn1_gelu", "tc_gemm_res_ln", "tc_gemm_inproj",
                                       "tc_gemm_outproj", "tc_gemm_f32", "tc_gemm_train"};
  MST_LAUNCHED(kNames[EPI], s);
  return MST_OK;
}

// Largest number of clusters of `cluster_size` CTAs that can be co-resident: with 4-CTA clusters only 132 of the
// 148 SMs can be covered (GPC granularity), and a persistent kernel launched with more clusters than that runs
// the surplus as a second wave - which doubled the LN GEMM's time before this cap (profiles/r01c_*).
template <typename Kernel>
static int max_active_clusters(Kernel kernel, int cluster_size, int threads, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster_size * (sm_count() / cluster_size));
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cluster_size;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = (sm_count() / cluster_size) * 7 / 8;  // conservative fallback
  }
  return n;
}

template <int EPI, int BN_>
static int max_pair_clusters() {
  using Cfg = PairCfg<EPI, BN_>;
  static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
  if (attr_set.first())
    cudaFuncSetAttribute(tc_gemm_pair_kernel<EPI, BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  static const int n = max_active_clusters(tc_gemm_pair_kernel<EPI, BN_>, 2, Cfg::THREADS, Cfg::SMEM_BYTES);
  return n;
}

template <int EPI, int BN_>
static int launch_gemm_pair_bn(const TcGemmParams& p, cudaStream_t s) {
  using Cfg = PairCfg<EPI, BN_>;
  CUtensorMap ta, tw, to;
  int rc;
  if ((rc = make_tmap_bf16(&ta, p.a, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K, BLOCK_M, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16(&tw, p.w, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, Cfg::BN / 2, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16_3d(&to, p.out, 1, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldo, (uint64_t)p.M * p.ldo, 32, 32, 64)))
    return rc;
  const int max_clusters = max_pair_clusters<EPI, BN_>();
  const int tiles = ceil_div(ceil_div(p.M, BLOCK_M), 2) * (p.N / Cfg::BN);
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  MST_CUDA_OK(launch_pdl(tc_gemm_pair_kernel<EPI, BN_>, dim3(2 * clusters), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, ta, tw, to, p));
  MST_LAUNCHED(EPI == TC_EPI_BIAS_BF16 ? "tc_gemm_qkv" : "tc_gemm_ffn1_gelu", s);
  return MST_OK;
}

// Tile width: 256 columns unless MST_PAIR_BN=128 asks for the narrow tile.  Measured at B=64 (round 2, graph-replayed,
// tools/gemm_timeline.py): QKV 33.7 us at 256 (8.03 waves), 35.2 at 192 (10.7 waves), 38.7 at 128 (16.05 waves); FFN1
// 25.1 at 256 (5.35 waves), 28.5 at 128 (10.7 waves) - the better wave balance of a narrow tile does not pay for its
// lower operand re-use and the extra per-tile hand-overs, so the wave model that picked 192 for QKV was dropped
// (as was the 192-column instantiation).
template <int EPI>
static int launch_gemm_pair(const TcGemmParams& p, cudaStream_t s) {
  static const int pin = getenv("MST_PAIR_BN") ? atoi(getenv("MST_PAIR_BN")) : 256;
  if (pin == 128 && p.N % 128 == 0) return launch_gemm_pair_bn<EPI, 128>(p, s);
  return launch_gemm_pair_bn<EPI, 256>(p, s);
}

template <bool F16>
static int launch_gemm_ln(const TcGemmParams& p, cudaStream_t s) {
  using Cfg = LnCfg;
  CUtensorMap ta, tw, tr, to;
  int rc;
  if ((rc = make_tmap_bf16(&ta, p.a, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K, BLOCK_M, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16(&tw, p.w, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, Cfg::BN / 2, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16(&tr, p.residual, (uint64_t)p.M, (uint64_t)LN_N, (uint64_t)LN_N, BLOCK_M, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&to, p.out, 1, (uint64_t)p.M, (uint64_t)LN_N, (uint64_t)LN_N, (uint64_t)p.M * LN_N, 32, 32, 64)))
    return rc;
  static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
  if (attr_set.first()) {
    MST_CUDA_OK(cudaFuncSetAttribute(tc_gemm_ln_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int m_pairs = ceil_div(ceil_div(p.M, BLOCK_M), 2);
  static const int max_clusters = max_active_clusters(tc_gemm_ln_kernel<F16>, 4, GEMM_THREADS, Cfg::SMEM_BYTES);
  const int clusters = m_pairs < max_clusters ? m_pairs : max_clusters;
  MST_CUDA_OK(launch_pdl(tc_gemm_ln_kernel<F16>, dim3(4 * clusters), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, ta, tw, tr, to, p));
  MST_LAUNCHED("tc_gemm_res_ln", s);
  return MST_OK;
}

template <bool F16>
static int launch_gemm_ln5(const TcGemmParams& p, cudaStream_t s) {
  using Cfg = Ln5Cfg;
  CUtensorMap ta, tw, tr, tp;
  int rc;
  if ((rc = make_tmap_bf16(&ta, p.a, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K, BLOCK_M, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16(&tw, p.w, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, Cfg::BN / 2, BLOCK_K))) return rc;
  if ((rc = make_tmap_bf16(&tr, p.residual, (uint64_t)p.M, (uint64_t)LN_N, (uint64_t)LN_N, 32, 64))) return rc;
  if ((rc = make_tmap_bf16(&tp, p.residual, (uint64_t)p.M, (uint64_t)LN_N, (uint64_t)LN_N, BLOCK_M, 64))) return rc;
  static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
  if (attr_set.first()) {
    MST_CUDA_OK(cudaFuncSetAttribute(tc_gemm_ln5_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int m_pairs = ceil_div(ceil_div(p.M, BLOCK_M), 2);
  static const int max_clusters = max_active_clusters(tc_gemm_ln5_kernel<F16>, 4, GEMM_THREADS, Cfg::SMEM_BYTES);
  const int clusters = m_pairs < max_clusters ? m_pairs : max_clusters;
  MST_CUDA_OK(launch_pdl(tc_gemm_ln5_kernel<F16>, dim3(4 * clusters), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, ta, tw, tr, tp, p));
  MST_LAUNCHED("tc_gemm_res_ln", s);
  return MST_OK;
}

static long long* g_gemm_dbg = nullptr;
void set_gemm_debug(long long* dev_buf) { g_gemm_dbg = dev_buf; }
long long* attn_debug_ptr() { return g_gemm_dbg; }

// Latency regime (B = 1 ... 8 trajectories: M of a few hundred rows): a 256 x 256 pair tile is one serial chain of
// K/64 k-blocks x 512 cycles on ONE SM pair while the rest of the chip idles (9 us per launch at B=1,
// profiles/r02i_*).  128 x 64 tiles on single CTAs cut the chain to K/64 x 128 cycles and spread it over N/64 times
// more SMs.  Threshold (measured, 50-step sweep at T=196): up to 48 wide tiles (B <= 4): 0.352 -> 0.312 ms per step at B=1,
// 0.372 -> 0.346 at B=4; from B=8 on the wide tiles win again.  MST_SMALL_GEMM=0 disables.
static bool small_problem(const TcGemmParams& p) {
  static const bool on = !(getenv("MST_SMALL_GEMM") && getenv("MST_SMALL_GEMM")[0] == '0');
  if (!on) return false;
  static const int max_tiles = getenv("MST_SMALL_GEMM_TILES") ? atoi(getenv("MST_SMALL_GEMM_TILES")) : 48;
  const int tiles = ceil_div(ceil_div(p.M, BLOCK_M), 2) * (p.N / 256);
  return tiles <= max_tiles;
}

int tc_gemm(const TcGemmParams& p_in, cudaStream_t s) {
  TcGemmParams p = p_in;
  p.dbg = g_gemm_dbg;
  {
    const char* e = getenv("MST_TEARDOWN");
    p.td_mode = e ? atoi(e) : 0;
  }
  MST_CHECK_ARG(p.a && p.w && p.out && (p.bias || p.epi == TC_EPI_TRAIN_F32), "null pointer");
  MST_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0, "empty problem");
  MST_CHECK_ARG(p.K % BLOCK_K == 0, "K must be a multiple of 64");
  switch (p.epi) {
    case TC_EPI_BIAS_BF16:
      MST_CHECK_ARG(p.N % 256 == 0 && p.ldo % 8 == 0, "N must be a multiple of 256");
      if (small_problem(p)) return launch_gemm<64, TC_EPI_BIAS_BF16>(p, s);
      return launch_gemm_pair<TC_EPI_BIAS_BF16>(p, s);
    case TC_EPI_BIAS_GELU_BF16:
      MST_CHECK_ARG(p.N % 256 == 0 && p.ldo % 8 == 0, "N must be a multiple of 256");
      if (small_problem(p)) return launch_gemm<64, TC_EPI_BIAS_GELU_BF16>(p, s);
      return launch_gemm_pair<TC_EPI_BIAS_GELU_BF16>(p, s);
    case TC_EPI_BIAS_RES_LN:
      MST_CHECK_ARG(p.N == LN_N && p.residual && p.ln_g && p.ln_b, "LN epilogue needs N == 512 and residual/gamma/beta");
      MST_CHECK_ARG(!p.ab_f16, "the LN GEMM multiplies bf16 operands");
      {
        static const int ln_v = getenv("MST_LN_V") ? atoi(getenv("MST_LN_V")) : 4;
        if (ln_v == 5) return p.io_f16 ? launch_gemm_ln5<true>(p, s) : launch_gemm_ln5<false>(p, s);
      }
      return p.io_f16 ? launch_gemm_ln<true>(p, s) : launch_gemm_ln<false>(p, s);
    case TC_EPI_INPROJ:
      MST_CHECK_ARG(p.N % 256 == 0 && p.pe && p.B > 0 && p.T > 0 && p.ldo % 8 == 0, "bad in-projection geometry");
      // K is only 3 k-blocks, so a tile is all epilogue: 128-wide tiles (392 of them at B=64, 2.65 waves of half the
      // work) beat 256-wide ones (196 tiles, 2 waves): 34 -> 29 us.  MST_INPROJ_BN=256 restores the wide tile.
      {
        static const int bn = getenv("MST_INPROJ_BN") ? atoi(getenv("MST_INPROJ_BN")) : 128;
        if (bn == 256) return launch_gemm<256, TC_EPI_INPROJ>(p, s);
      }
      return launch_gemm<128, TC_EPI_INPROJ>(p, s);
    case TC_EPI_OUTPROJ_F32:
      MST_CHECK_ARG(p.N % 64 == 0 && p.B > 0 && p.T > 0 && p.n_valid > 0, "bad out-projection geometry");
      return launch_gemm<64, TC_EPI_OUTPROJ_F32>(p, s);
    case TC_EPI_BIAS_F32:
      MST_CHECK_ARG(p.N % 64 == 0 && p.ldo % 4 == 0, "N must be a multiple of 64");
      if (p.N % 256 == 0) return launch_gemm<256, TC_EPI_BIAS_F32>(p, s);
      return launch_gemm<64, TC_EPI_BIAS_F32>(p, s);
    case TC_EPI_TRAIN_F32:
      MST_CHECK_ARG(p.N % 64 == 0 && p.ldo % 4 == 0, "N must be a multiple of 64");
      {
        // Tile width by a wave model: a tile of width BN costs BN/256 + 0.1 (fixed TMA -> MMA -> epilogue latency) and the
        // kernel takes ceil(tiles / SMs) of them.  Small problems (the 77-row B=1 steps of the finetune loop, their batched
        // backward, the weight-gradient GEMMs with 12 x 2 wide tiles and a 77-block reduction) get 4x more, 4x shorter
        // CTAs; problems of many waves keep the 256-wide tile and its operand reuse.  MST_TRAIN_BN pins a width.
        static const int pin = getenv("MST_TRAIN_BN") ? atoi(getenv("MST_TRAIN_BN")) : 0;
        int best = 64;
        float best_cost = 1e30f;
        for (int bn = 256; bn >= 64; bn >>= 1) {
          if (p.N % bn != 0) continue;
          const int tiles = ceil_div(p.M, BLOCK_M) * (p.N / bn);
          const float cost = (float)ceil_div(tiles, sm_count()) * ((float)bn / 256.0f + 0.1f);
          if (cost < best_cost) { best_cost = cost; best = bn; }
        }
        if (pin == 256 || pin == 128 || pin == 64) best = (p.N % pin == 0) ? pin : 64;
        if (best == 256) return launch_gemm<256, TC_EPI_TRAIN_F32>(p, s);
        if (best == 128) return launch_gemm<128, TC_EPI_TRAIN_F32>(p, s);
      }
      return launch_gemm<64, TC_EPI_TRAIN_F32>(p, s);
    default:
      return fail(MST_ERR_INVALID, "tc_gemm: unknown epilogue");
  }
}

// ---------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------
// x[b][f][t] fp32 -> a[(b*T + t)][f] bf16, columns [F, f_pad) zero.  32x32 smem transpose.
// The same launch also writes token 0 of every sequence (blockIdx.y == gridDim.y - 1, one block per (pass, b)):
// temb[row] + (uncond ? txt_b : text_emb[b]) + pe[0]   (reference mdm_forstyledataset.py:322-327, :344-345) -
// one launch less per denoise step than a separate token-0 kernel.
__global__ void __launch_bounds__(256) motion_to_tokens_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a,
                                                               int F, int T, int f_pad, const Token0Params t0, int n_seqs) {
  pdl_launch_dependents();
  pdl_wait();  // *temb_row_dev is decremented by the previous step's update kernel
  if (blockIdx.y == gridDim.y - 1) {
    const int seq = blockIdx.z * gridDim.x + blockIdx.x;
    if (seq >= n_seqs || (!t0.x_bf16 && !t0.x_f16 && !t0.x_f32)) return;
    const int b = seq % t0.B;
    const bool uncond = t0.cfg ? (seq >= t0.B) : (t0.uncond != 0);
    const int row = t0.temb_row_dev ? (*t0.temb_row_dev + t0.temb_row_offset) : (b + t0.temb_row_offset);
    const int S = t0.T + 1;
    for (int n = threadIdx.x; n < t0.d; n += blockDim.x) {
      float v = t0.temb[(int64_t)row * t0.d + n];
      if (t0.txt_b) v += (uncond || !t0.text_emb) ? t0.txt_b[n] : t0.text_emb[(int64_t)b * t0.d + n];
      v += t0.pe[n];
      const int64_t o = (int64_t)seq * S * t0.d + n;
      if (t0.x_f32) t0.x_f32[o] = v;
      if (t0.x_bf16) t0.x_bf16[o] = __float2bfloat16_rn(v);
      if (t0.x_f16) t0.x_f16[o] = __float2half_rn(v);
    }
    return;
  }
  __shared__ float tile[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0_ = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int f = f0 + i, t = t0_ + tx;
    tile[i][tx] = (f < F && t < T) ? x[((size_t)b * F + f) * T + t] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int t = t0_ + i, f = f0 + tx;
    if (t < T && f < f_pad) a[((size_t)b * T + t) * f_pad + f] = __float2bfloat16_rn(tile[tx][i]);
  }
}

int motion_to_tokens_bf16(const float* x, __nv_bfloat16* a, int B, int F, int T, int f_pad, const Token0Params* t0,
                          int n_seqs, cudaStream_t s) {
  Token0Params tp;
  if (t0) tp = *t0;
  int gx = ceil_div(T, 32);
  if (t0 && gx * B < n_seqs) gx = ceil_div(n_seqs, B);  // enough token-0 blocks (surplus transpose blocks fall outside T)
  dim3 grid(gx, ceil_div(f_pad, 32) + 1, B);
  MST_CUDA_OK(launch_pdl(motion_to_tokens_kernel, grid, dim3(256), 0, s, x, a, F, T, f_pad, tp, t0 ? n_seqs : 0));
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

// src fp32 [rows, cols] -> dst bf16 [rows, cols] and / or dst_t bf16 [cols, rows_pad] (zero padded), 32 x 32 smem transpose
__global__ void __launch_bounds__(256) cvt_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                                       __nv_bfloat16* __restrict__ dst, __nv_bfloat16* __restrict__ dst_t,
                                                       int rows_pad) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    const float v = (r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.0f;
    tile[i][tx] = v;
    if (dst && r < rows && c < cols) dst[(size_t)r * cols + c] = __float2bfloat16_rn(v);
  }
  if (!dst_t) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows_pad) dst_t[(size_t)c * rows_pad + r] = __float2bfloat16_rn(tile[tx][i]);
  }
}

int cvt_bf16(const float* src, int rows, int cols, int ld, __nv_bfloat16* dst, __nv_bfloat16* dst_t, int rows_pad,
             cudaStream_t s) {
  const int rgrid = dst_t ? ceil_div(rows_pad, 32) : ceil_div(rows, 32);
  MST_CUDA_OK(launch_pdl(cvt_bf16_kernel, dim3(ceil_div(cols, 32), rgrid), dim3(256), 0, s, src, rows, cols, ld, dst, dst_t, rows_pad));
  MST_LAUNCHED("cvt_bf16", s);
  return MST_OK;
}

// Several conversions in ONE launch (the per-weight / per-operand launches were pure latency on the finetune step: 83
// launches to re-pack the weights after every optimizer step, 3 per backward linear): every job turns an fp32 matrix into any
// of a zero-padded bf16 copy, a transposed bf16 copy, a zero-padded fp16 copy, and column sums; one CTA per 32 x 32 tile.
constexpr int CVT_TILE = 64;  // rows and columns of one CTA's tile

__global__ void __launch_bounds__(256) cvt_multi_kernel(const __grid_constant__ CvtJobs jobs) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel of the stream (programmatic dependent launch)
  // 64 x 64 tile: 16 values per thread in flight (128-bit loads when the rows allow), transposed rows leave as 128-byte runs
  __shared__ float tile[CVT_TILE][CVT_TILE + 1];
  __shared__ float red[8][CVT_TILE];
  int k = 0;
  while (k + 1 < jobs.n && (int)blockIdx.x >= jobs.j[k + 1].tile0) ++k;
  const CvtJob& J = jobs.j[k];
  const int t = (int)blockIdx.x - J.tile0;
  const int r0 = (t / J.tiles_x) * CVT_TILE, c0 = (t % J.tiles_x) * CVT_TILE;
  const int rows = J.rows, cols = J.cols;
  const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;  // float4 column of the tile, row phase (16 rows per pass)
  const bool vec = (J.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(J.src) & 15) == 0;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + rr + 16 * i, c = c0 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const float* sp = J.src + (size_t)r * J.ld + c;
      if (vec && c + 3 < cols) {
        v = __ldg(reinterpret_cast<const float4*>(sp));
      } else {
        if (c < cols) v.x = __ldg(sp);
        if (c + 1 < cols) v.y = __ldg(sp + 1);
        if (c + 2 < cols) v.z = __ldg(sp + 2);
        if (c + 3 < cols) v.w = __ldg(sp + 3);
      }
    }
    float* tr = &tile[rr + 16 * i][4 * q];
    tr[0] = v.x; tr[1] = v.y; tr[2] = v.z; tr[3] = v.w;
    part[0] += v.x; part[1] += v.y; part[2] += v.z; part[3] += v.w;
    if (r < J.rows_pad) {
      const float e[4] = {v.x, v.y, v.z, v.w};
      if ((J.cols_pad & 3) == 0 && c + 3 < J.cols_pad) {  // whole group inside the padded row: 8-byte stores
        if (J.dst) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
          *reinterpret_cast<uint2*>(J.dst + (size_t)r * J.cols_pad + c) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
        if (J.dst_h) {
          const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
          *reinterpret_cast<uint2*>(J.dst_h + (size_t)r * J.cols_pad + c) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c + u < J.cols_pad) {
            if (J.dst) J.dst[(size_t)r * J.cols_pad + c + u] = __float2bfloat16_rn(e[u]);
            if (J.dst_h) J.dst_h[(size_t)r * J.cols_pad + c + u] = __float2half_rn(e[u]);
          }
      }
    }
  }
  if (!J.dst_t && !J.colsum) return;
  if (J.colsum) {  // the 16 row phases of a column group meet through the quarter-warp shuffles, then shared memory
#pragma unroll
    for (int u = 0; u < 4; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], 16);
    if ((threadIdx.x & 16) == 0) {
      float* rd = &red[0][0] + (threadIdx.x >> 5) * CVT_TILE + 4 * q;  // 8 warps x 64 columns
      rd[0] = part[0]; rd[1] = part[1]; rd[2] = part[2]; rd[3] = part[3];
    }
  }
  __syncthreads();
  if (J.colsum && threadIdx.x < CVT_TILE && c0 + (int)threadIdx.x < cols && r0 < rows) {
    const float* rd = &red[0][0];
    float a = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += rd[w * CVT_TILE + threadIdx.x];
    atomicAdd(J.colsum + c0 + threadIdx.x, a);
  }
  if (J.dst_t) {
    // warp w writes transposed rows c0 + w, c0 + w + 8, ...: lane l the source rows r0 + 2 l, r0 + 2 l + 1 (4 bytes)
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int r = r0 + 2 * l;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cc = w + 8 * i, c = c0 + cc;
      if (c >= cols) continue;
      const float x0 = tile[2 * l][cc], x1 = tile[2 * l + 1][cc];
      __nv_bfloat16* dp = J.dst_t + (size_t)c * J.t_ld + r;
      if ((J.t_ld & 1) == 0 && r + 1 < J.t_ld) {
        *reinterpret_cast<__nv_bfloat162*>(dp) = __floats2bfloat162_rn(x0, x1);
      } else {
        if (r < J.t_ld) dp[0] = __float2bfloat16_rn(x0);
        if (r + 1 < J.t_ld) dp[1] = __float2bfloat16_rn(x1);
      }
    }
  }
}

void cvt_jobs_add(CvtJobs& js, const float* src, int rows, int cols, int ld, __nv_bfloat16* dst, __half* dst_h, int rows_pad,
                  int cols_pad, __nv_bfloat16* dst_t, int t_ld, float* colsum) {
  CvtJob& J = js.j[js.n];
  J.src = src; J.dst = dst; J.dst_t = dst_t; J.dst_h = dst_h; J.colsum = colsum;
  J.rows = rows; J.cols = cols; J.ld = ld;
  J.rows_pad = (dst || dst_h) ? rows_pad : 0;
  J.cols_pad = (dst || dst_h) ? cols_pad : 0;
  J.t_ld = dst_t ? t_ld : 0;
  int r_ext = J.rows_pad > J.t_ld ? J.rows_pad : J.t_ld, c_ext = J.cols_pad > cols ? J.cols_pad : cols;
  if (colsum && r_ext < rows) r_ext = rows;
  if (!dst_t && !colsum) c_ext = J.cols_pad;
  J.tiles_x = ceil_div(c_ext, CVT_TILE);
  J.tile0 = js.n ? js.j[js.n - 1].tile0 + js.tiles_of_last : 0;
  js.tiles_of_last = J.tiles_x * ceil_div(r_ext, CVT_TILE);
  ++js.n;
}

int cvt_multi(const CvtJobs& js, cudaStream_t s, const char* name) {
  if (js.n <= 0) return MST_OK;
  const int tiles = js.j[js.n - 1].tile0 + js.tiles_of_last;
  if (tiles <= 0) return MST_OK;
  MST_CUDA_OK(launch_pdl(cvt_multi_kernel, dim3(tiles), dim3(256), 0, s, js));
  MST_LAUNCHED(name, s);
  return MST_OK;
}

__global__ void __launch_bounds__(256) pack_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int rows,
                                                       int cols, int rows_pad, int cols_pad) {
  const size_t total = (size_t)rows_pad * cols_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / cols_pad), c = (int)(i - (size_t)r * cols_pad);
    dst[i] = __float2half_rn((r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.0f);
  }
}

int pack_f16(const float* src, __half* dst, int rows, int cols, int rows_pad, int cols_pad, cudaStream_t s) {
  size_t total = (size_t)rows_pad * cols_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_f16_kernel<<<blocks, 256, 0, s>>>(src, dst, rows, cols, rows_pad, cols_pad);
  MST_LAUNCHED("pack_f16", s);
  return MST_OK;
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
