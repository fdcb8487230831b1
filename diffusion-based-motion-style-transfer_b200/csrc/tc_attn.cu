// Self-attention of one (sequence, head) per CTA on tcgen05 tensor cores:
//   O = softmax(Q K^T / sqrt(dh)) V,   S = T+1 <= 208 tokens, dh = 128, no mask
// (torch.nn.MultiheadAttention as used by nn.TransformerEncoderLayer in the
// reference, model/mdm_forstyledataset.py:231-238 / :346).
//
// The whole key range fits one tile, so there is no online-softmax rescaling:
//   1. TMA: K, V ([s_pad x 128] each) and one 128-row Q tile -> 128B-swizzled smem
//   2. tcgen05.mma  S[128 x s_pad] = Q K^T            (fp32 in TMEM cols [0, s_pad))
//   3. 128 threads, one query row each: tcgen05.ld the row, max, exp2, sum;
//      write P (bf16) to shared memory in the K-major swizzled operand layout
//   4. tcgen05.mma  O[128 x 128] = P V   (V consumed MN-major straight from the
//      QKV buffer's layout - no transpose pass)        (TMEM cols [256, 384))
//   5. O * (1/rowsum) -> bf16 -> global
// Steps 2-5 repeat for the second Q tile (rows 128..S-1) with K/V resident.
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace mst {

using namespace ptx;

constexpr int ATT_DH = 128;
constexpr int ATT_MAX_SPAD = 208;
constexpr int ATT_THREADS = 160;  // warps 0-3: softmax/epilogue (one row per thread); warp 4: TMA + MMA issue
constexpr int ATT_Q_BYTES = 2 * 128 * 128;
constexpr int ATT_KV_BYTES = 2 * ATT_MAX_SPAD * 128;
constexpr int ATT_P_BYTES = 4 * 128 * 128;
constexpr int ATT_SMEM_BYTES = 1024 + ATT_Q_BYTES + 2 * ATT_KV_BYTES + ATT_P_BYTES + 64;
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_O_COL = 256;

__global__ void __launch_bounds__(ATT_THREADS, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    __nv_bfloat16* __restrict__ out, int S, int s_pad, int d_model, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t kv_box = (uint32_t)s_pad * 128u;  // bytes of one [s_pad x 64] box
  const uint32_t q_smem = base;
  const uint32_t k_smem = q_smem + ATT_Q_BYTES;
  const uint32_t v_smem = k_smem + ATT_KV_BYTES;
  const uint32_t p_smem = v_smem + ATT_KV_BYTES;
  const uint32_t bar_base = p_smem + ATT_P_BYTES;
  const uint32_t kv_bar = bar_base, q_bar = bar_base + 8, s_full = bar_base + 16, p_ready = bar_base + 24,
                 o_full = bar_base + 32, tmem_slot = bar_base + 40;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + ATT_Q_BYTES + 2 * ATT_KV_BYTES + ATT_P_BYTES + 40);
  uint8_t* p_ptr = base_ptr + ATT_Q_BYTES + 2 * ATT_KV_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x, seq = blockIdx.y;
  const int row0 = seq * S;
  const int n_qt = (S + 127) / 128;

  if (warp == 4 && elect_one()) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_kv);
    mbar_init(kv_bar, 1);
    mbar_init(q_bar, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 4) {
    if (elect_one()) {
      // K and V: two 64-column boxes each
      mbar_expect_tx(kv_bar, 4 * kv_box);
      for (int kc = 0; kc < 2; ++kc) {
        tma_load_2d(k_smem + kc * kv_box, &tmap_kv, kv_bar, d_model + head * ATT_DH + kc * 64, row0);
        tma_load_2d(v_smem + kc * kv_box, &tmap_kv, kv_bar, 2 * d_model + head * ATT_DH + kc * 64, row0);
      }
      const uint32_t idesc_qk = make_idesc_bf16(128, s_pad, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, ATT_DH, 1);
      for (int qt = 0; qt < n_qt; ++qt) {
        const uint32_t par = qt & 1;
        // the previous tile's QK^T has been consumed (p_ready implies s_full), Q smem is free
        mbar_expect_tx(q_bar, ATT_Q_BYTES);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_2d(q_smem + kc * 16384, &tmap_q, q_bar, head * ATT_DH + kc * 64, row0 + qt * 128);
        if (qt == 0) mbar_wait(kv_bar, 0);
        mbar_wait(q_bar, par);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < ATT_DH / 16; ++ks) {
          const int kc = ks >> 2, k4 = ks & 3;
          const uint64_t adesc = make_smem_desc_sw128(q_smem + kc * 16384 + k4 * 32, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(k_smem + kc * kv_box + k4 * 32, 0, 1024);
          mma_bf16_ss(tmem_base, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
        }
        mma_commit(s_full);
        // P V once the softmax threads have published P
        mbar_wait(p_ready, par);
        tc_fence_after();
        const int n_ks = s_pad / 16;
        for (int ks = 0; ks < n_ks; ++ks) {
          const int kc = ks >> 2, k4 = ks & 3;
          const uint64_t adesc = make_smem_desc_sw128(p_smem + kc * 16384 + k4 * 32, 0, 1024);
          // V is [keys][dh]: N (=dh) contiguous -> MN-major B; 16 keys = 2 groups of 8 rows (SBO),
          // the second 64 dh columns live in the next box (LBO)
          const uint64_t bdesc = make_smem_desc_sw128(v_smem + ks * 2048, kv_box, 1024);
          mma_bf16_ss(tmem_base + ATT_O_COL, adesc, bdesc, idesc_pv, ks != 0 ? 1u : 0u);
        }
        mma_commit(o_full);
        // Q smem is rewritten next iteration: QK^T of this tile completed long ago
        // (s_full fired before p_ready), so no extra wait is needed.
      }
    }
  } else {
    const int r = threadIdx.x;  // query row inside the tile, also the TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int qt = 0; qt < n_qt; ++qt) {
      const uint32_t par = qt & 1;
      mbar_wait(s_full, par);
      tc_fence_after();
      float mx = -INFINITY;
      for (int c = 0; c < s_pad; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < S) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      float sum = 0.0f;
      for (int c = 0; c < s_pad; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + lane_addr + c, v);
        tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          e[j] = (c + j < S) ? exp2f((__uint_as_float(v[j]) - mx) * scale_log2e) : 0.0f;
          sum += e[j];
        }
        // two 16-byte chunks (8 keys each) of row r in key-chunk buffer kc, 128B-swizzled
        const int kc = c >> 6, chunk = (c & 63) >> 3;
        uint8_t* rowp = p_ptr + kc * 16384 + r * 128;
        uint4 lo = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        uint4 hi = make_uint4(pack_bf16x2(e[8], e[9]), pack_bf16x2(e[10], e[11]), pack_bf16x2(e[12], e[13]), pack_bf16x2(e[14], e[15]));
        *reinterpret_cast<uint4*>(rowp + (((chunk) ^ (r & 7)) << 4)) = lo;
        *reinterpret_cast<uint4*>(rowp + (((chunk + 1) ^ (r & 7)) << 4)) = hi;
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_ready);
      mbar_wait(o_full, par);
      tc_fence_after();
      const float inv = 1.0f / sum;
      const int q = qt * 128 + r;
      for (int c = 0; c < ATT_DH; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_addr + ATT_O_COL + c, v);
        tmem_ld_wait();
        if (q < S) {
          uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(row0 + q) * d_model + head * ATT_DH + c);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[g] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv));
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int tc_attention(const TcAttnParams& p, cudaStream_t s) {
  MST_CHECK_ARG(p.qkv && p.out, "null pointer");
  MST_CHECK_ARG(p.n_seqs > 0 && p.S > 0, "empty problem");
  MST_CHECK_ARG(p.d_model == p.n_heads * ATT_DH, "attention kernel is built for head_dim 128");
  const int s_pad = (p.S + 15) / 16 * 16;
  if (s_pad > ATT_MAX_SPAD)
    return fail(MST_ERR_UNSUPPORTED, "tc_attention: sequences longer than 208 tokens (T > 207) are not supported");
  const uint64_t M = (uint64_t)p.n_seqs * p.S;
  CUtensorMap tq, tkv;
  int rc;
  if ((rc = make_tmap_bf16(&tq, p.qkv, M, 3 * (uint64_t)p.d_model, 3 * (uint64_t)p.d_model, 128, 64))) return rc;
  if ((rc = make_tmap_bf16(&tkv, p.qkv, M, 3 * (uint64_t)p.d_model, 3 * (uint64_t)p.d_model, (uint32_t)s_pad, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MST_CUDA_OK(cudaFuncSetAttribute(tc_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    attr_set = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATT_DH);
  dim3 grid(p.n_heads, p.n_seqs);
  tc_attention_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, s>>>(tq, tkv, p.out, p.S, s_pad, p.d_model, scale_log2e);
  MST_LAUNCHED("tc_attention", s);
  return MST_OK;
}

}  // namespace mst
