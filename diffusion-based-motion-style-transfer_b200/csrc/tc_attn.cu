// Self-attention on tcgen05 tensor cores, persistent and software-pipelined:
//   O = softmax(Q K^T / sqrt(dh)) V,   S = T+1 <= 208 tokens, dh = 128, no mask
// (torch.nn.MultiheadAttention as used by nn.TransformerEncoderLayer in the reference,
// model/mdm_forstyledataset.py:231-238 / :346).
//
// Work item = (sequence, head, 128-query tile).  The whole key range of a sequence fits one tile, so there is
// no online-softmax rescaling.  Every CTA owns a contiguous range of items and runs three roles:
//   warp 0          TMA producer: K and V of a (sequence, head) once, Q per item (two Q buffers), straight out of
//                   the packed QKV activation through a 3-D tensor map [sequence][token][column] - rows past
//                   the end of a sequence are out of bounds (zero-filled), never the next sequence's data
//   warp 1          tcgen05.mma issuer:  S = Q K^T  (fp32, TMEM)  and  O = P V  with P read from TMEM
//   warps 4-7, 8-11 two softmax groups of 128 threads (one query row each) that alternate items, so the
//                   row max / exp2 / sum of item i overlaps the Q K^T of item i+1 and the P V of item i-1.
//                   P (bf16) is written back into TMEM over the S columns it was computed from - it never
//                   touches shared memory - and O leaves through per-warp [32 x 32] staging boxes and TMA stores.
// TMEM slot (2 slots, 256 columns apart): S fp32 in [0, s_pad), P bf16x2 in [0, s_pad/2),
// O fp32 in [s_pad/2, s_pad/2 + 128) (the tail of S is dead once the softmax has read it).
#include "tc.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace mst {

using namespace ptx;

constexpr int ATT_DH = 128;
constexpr int ATT_MAX_SPAD = 208;
constexpr int ATT_THREADS = 384;
constexpr int ATT_Q_BYTES = 2 * 128 * 128;             // one Q tile: two 64-column halves of 128 rows
constexpr int ATT_KV_BYTES = 2 * ATT_MAX_SPAD * 128;   // K (or V): two 64-column halves of s_pad rows
constexpr int ATT_OBOX_BYTES = 32 * 64;                // [32 rows x 32 bf16], 64B swizzle
constexpr int ATT_STAGING_BYTES = 8 * 2 * ATT_OBOX_BYTES;
constexpr int ATT_SMEM_BYTES = 1024 + 2 * ATT_Q_BYTES + 2 * ATT_KV_BYTES + ATT_STAGING_BYTES + 256;
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_SLOT_COLS = 256;

struct AttnGeom {
  int S, s_pad, n_qt, n_heads, n_items, d_model;
  float scale_log2e;
  long long* dbg;  // developer hook (mst_test_set_gemm_debug): clock64 timeline of CTA 0
};

#define ATT_STAMP() do { if (dbg && di < 1000) dbg[di++] = clock64(); } while (0)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_o, const AttnGeom g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t kv_box = (uint32_t)g.s_pad * 128u;  // bytes of one [s_pad x 64] box
  const uint32_t q_smem = base;                       // 2 buffers
  const uint32_t k_smem = q_smem + 2 * ATT_Q_BYTES;
  const uint32_t v_smem = k_smem + ATT_KV_BYTES;
  const uint32_t o_smem = v_smem + ATT_KV_BYTES;
  const uint32_t bar_base = o_smem + ATT_STAGING_BYTES;
  const uint32_t k_full = bar_base, k_free = bar_base + 8, v_full = bar_base + 16, v_free = bar_base + 24;
  auto q_full = [&](int b) { return bar_base + 32 + 8 * b; };
  auto q_free = [&](int b) { return bar_base + 48 + 8 * b; };
  auto s_full = [&](int s) { return bar_base + 64 + 8 * s; };
  auto p_ready = [&](int s) { return bar_base + 80 + 8 * s; };
  auto o_full = [&](int s) { return bar_base + 96 + 8 * s; };
  auto slot_free = [&](int s) { return bar_base + 112 + 8 * s; };
  const uint32_t tmem_slot = bar_base + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // contiguous, balanced range of work items of this CTA
  const int per = g.n_items / (int)gridDim.x, extra = g.n_items % (int)gridDim.x;
  const int i0 = (int)blockIdx.x * per + min((int)blockIdx.x, extra);
  const int i1 = i0 + per + ((int)blockIdx.x < extra ? 1 : 0);

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_kv);
    prefetch_tensormap(&tmap_o);
    mbar_init(k_full, 1);
    mbar_init(k_free, 1);
    mbar_init(v_full, 1);
    mbar_init(v_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(q_full(b), 1);
      mbar_init(q_free(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(p_ready(b), 128);
      mbar_init(o_full(b), 1);
      mbar_init(slot_free(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // the QKV projection has completed
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int p_cols = g.s_pad / 2;  // P: two bf16 per 32-bit column; O starts right behind it

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (elect_one()) {
      int n_units = 0;
      auto load_q = [&](int it) {
        const int j = it - i0, qb = j & 1;
        const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
        const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
        mbar_wait(q_free(qb), ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(q_full(qb), ATT_Q_BYTES);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_3d(q_smem + qb * ATT_Q_BYTES + kc * 16384, &tmap_q, q_full(qb), head * ATT_DH + kc * 64, qt * 128, seq);
      };
      int q_loaded = i0;  // items < q_loaded have their Q load issued
      for (int it = i0; it < i1; ++it) {
        const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
        const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
        const bool new_unit = (it == i0) || qt == 0;
        if (new_unit) {
          mbar_wait(k_free, (n_units & 1) ^ 1);
          mbar_expect_tx(k_full, 2 * kv_box);
          for (int kc = 0; kc < 2; ++kc)
            tma_load_3d(k_smem + kc * kv_box, &tmap_kv, k_full, g.d_model + head * ATT_DH + kc * 64, 0, seq);
        }
        if (q_loaded <= it) load_q(q_loaded++);
        if (new_unit) {
          // the second query tile of this unit does not depend on V's buffer: issue it before waiting for v_free
          if (q_loaded == it + 1 && it + 1 < i1 && qt + 1 < g.n_qt) load_q(q_loaded++);
          mbar_wait(v_free, (n_units & 1) ^ 1);
          mbar_expect_tx(v_full, 2 * kv_box);
          for (int kc = 0; kc < 2; ++kc)
            tma_load_3d(v_smem + kc * kv_box, &tmap_kv, v_full, 2 * g.d_model + head * ATT_DH + kc * 64, 0, seq);
          ++n_units;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc_bf16(128, g.s_pad, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, ATT_DH, 1);
      int k_units = 0, v_units = 0;  // units whose K / V has been waited for
      long long* dbg = (g.dbg && blockIdx.x == 0) ? g.dbg + 1024 : nullptr;
      int di = 0;
      auto is_new_unit = [&](int it) { return it == i0 || (it % g.n_qt) == 0; };
      auto is_last_of_unit = [&](int it) { return it == i1 - 1 || ((it + 1) % g.n_qt) == 0; };
      auto issue_qk = [&](int it) {
        const int j = it - i0, slot = j & 1, qb = j & 1;
        if (is_new_unit(it)) {
          mbar_wait(k_full, k_units & 1);
          ++k_units;
        }
        mbar_wait(q_full(qb), (j >> 1) & 1);
        mbar_wait(slot_free(slot), ((j >> 1) & 1) ^ 1);  // O of item j-2 has been read out of this slot
        tc_fence_after();
        ATT_STAMP();
        const uint32_t d = tmem_base + (uint32_t)(slot * ATT_SLOT_COLS);
#pragma unroll
        for (int ks = 0; ks < ATT_DH / 16; ++ks) {
          const int kc = ks >> 2, k4 = ks & 3;
          const uint64_t adesc = make_smem_desc_sw128(q_smem + qb * ATT_Q_BYTES + kc * 16384 + k4 * 32, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(k_smem + kc * kv_box + k4 * 32, 0, 1024);
          mma_bf16_ss(d, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
        }
        mma_commit(s_full(slot));
        mma_commit(q_free(qb));
        if (is_last_of_unit(it)) mma_commit(k_free);
      };
      auto issue_pv = [&](int it) {
        const int j = it - i0, slot = j & 1;
        mbar_wait(p_ready(slot), (j >> 1) & 1);
        if (is_new_unit(it)) {
          mbar_wait(v_full, v_units & 1);
          ++v_units;
        }
        tc_fence_after();
        ATT_STAMP();
        const uint32_t slot_base = tmem_base + (uint32_t)(slot * ATT_SLOT_COLS);
        const int n_ks = g.s_pad / 16;
        for (int ks = 0; ks < n_ks; ++ks) {
          // V is [keys][dh]: N (= dh) contiguous -> MN-major B; 16 keys = 2 groups of 8 rows (SBO), the second
          // 64 dh columns live in the next box (LBO).  A = P: 16 keys = 8 packed columns per step.
          const uint64_t bdesc = make_smem_desc_sw128(v_smem + ks * 2048, kv_box, 1024);
          mma_bf16_ts(slot_base + (uint32_t)p_cols, slot_base + (uint32_t)(ks * 8), bdesc, idesc_pv, ks != 0 ? 1u : 0u);
        }
        mma_commit(o_full(slot));
        if (is_last_of_unit(it)) mma_commit(v_free);
      };
      if (i0 < i1) issue_qk(i0);
      for (int it = i0; it < i1; ++it) {
        if (it + 1 < i1) issue_qk(it + 1);
        issue_pv(it);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ softmax / output groups
    const int grp = (warp - 4) >> 2;  // 0: even items, 1: odd items
    const int quad = warp & 3;
    const int ew = warp - 4;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t my_box = o_smem + ew * 2 * ATT_OBOX_BYTES;
    long long* dbg = (g.dbg && blockIdx.x == 0 && (warp & 3) == 0 && lane == 0) ? g.dbg + (2 + grp) * 1024 : nullptr;
    int di = 0;
    for (int it = i0 + grp; it < i1; it += 2) {
      const int j = it - i0, slot = j & 1;
      const uint32_t par = (j >> 1) & 1;
      const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
      const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
      const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(slot * ATT_SLOT_COLS);
      const bool row_valid = qt * 128 + r < g.S;
      const bool warp_valid = qt * 128 + quad * 32 < g.S;  // any valid row in this warp
      ATT_STAMP();
      mbar_wait(s_full(slot), par);
      tc_fence_after();
      ATT_STAMP();
      float sum = 0.0f;
      if (warp_valid) {
        // Row softmax in two passes over the S row held in TMEM.  The chain LDTM -> wait -> math is latency-bound with
        // one warp per scheduler, so: loads are double-buffered (the next 32 columns are in flight while the current
        // 32 are processed), the running max / sum use four independent accumulators, and only the last (partial)
        // chunk pays for the key mask.
        const int n_full = g.S >> 5;                 // chunks of 32 keys that are entirely valid
        const int tail0 = n_full << 5;               // first key of the masked tail
        const int tail_w = g.s_pad - tail0;          // 0, 16 or 32 columns
        uint32_t va[32], vb[32];
        // ---- pass 1: row maximum
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        auto max32 = [&](const uint32_t (&v)[32]) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[k]));
            m1 = fmaxf(m1, __uint_as_float(v[k + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[k + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[k + 3]));
          }
        };
        if (n_full > 0) tmem_ld32(s_addr, va);
        for (int i = 0; i < n_full; i += 2) {
          tmem_ld_wait();
          if (i + 1 < n_full) tmem_ld32(s_addr + (i + 1) * 32, vb);
          max32(va);
          if (i + 1 < n_full) {
            tmem_ld_wait();
            if (i + 2 < n_full) tmem_ld32(s_addr + (i + 2) * 32, va);
            max32(vb);
          }
        }
        if (tail_w == 32) {
          tmem_ld32(s_addr + tail0, va);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (tail0 + k < g.S) m0 = fmaxf(m0, __uint_as_float(va[k]));
        } else if (tail_w == 16) {
          uint32_t vt[16];
          tmem_ld16(s_addr + tail0, vt);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (tail0 + k < g.S) m0 = fmaxf(m0, __uint_as_float(vt[k]));
        }
        ATT_STAMP();
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        const float moff = mx * g.scale_log2e;
        // ---- pass 2: p = exp2((s - max) * scale), row sum, bf16 P written back over S
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        auto exp32 = [&](const uint32_t (&v)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float e0 = fast_exp2(fmaf(__uint_as_float(v[k]), g.scale_log2e, -moff));
            const float e1 = fast_exp2(fmaf(__uint_as_float(v[k + 1]), g.scale_log2e, -moff));
            const float e2 = fast_exp2(fmaf(__uint_as_float(v[k + 2]), g.scale_log2e, -moff));
            const float e3 = fast_exp2(fmaf(__uint_as_float(v[k + 3]), g.scale_log2e, -moff));
            s0 += e0; s1 += e1; s2 += e2; s3 += e3;
            pk[k >> 1] = pack_bf16x2(e0, e1);
            pk[(k >> 1) + 1] = pack_bf16x2(e2, e3);
          }
          tmem_st16(s_addr + (c >> 1), pk);
        };
        if (n_full > 0) tmem_ld32(s_addr, va);
        for (int i = 0; i < n_full; i += 2) {
          tmem_ld_wait();
          if (i + 1 < n_full) tmem_ld32(s_addr + (i + 1) * 32, vb);
          exp32(va, i * 32);
          if (i + 1 < n_full) {
            tmem_ld_wait();
            if (i + 2 < n_full) tmem_ld32(s_addr + (i + 2) * 32, va);
            exp32(vb, (i + 1) * 32);
          }
        }
        if (tail_w == 32) {
          tmem_ld32(s_addr + tail0, va);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const float e0 = (tail0 + k < g.S) ? fast_exp2(fmaf(__uint_as_float(va[k]), g.scale_log2e, -moff)) : 0.0f;
            const float e1 = (tail0 + k + 1 < g.S) ? fast_exp2(fmaf(__uint_as_float(va[k + 1]), g.scale_log2e, -moff)) : 0.0f;
            s0 += e0; s1 += e1;
            pk[k >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st16(s_addr + (tail0 >> 1), pk);
        } else if (tail_w == 16) {
          uint32_t vt[16];
          tmem_ld16(s_addr + tail0, vt);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 16; k += 2) {
            const float e0 = (tail0 + k < g.S) ? fast_exp2(fmaf(__uint_as_float(vt[k]), g.scale_log2e, -moff)) : 0.0f;
            const float e1 = (tail0 + k + 1 < g.S) ? fast_exp2(fmaf(__uint_as_float(vt[k + 1]), g.scale_log2e, -moff)) : 0.0f;
            s0 += e0; s1 += e1;
            pk[k >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st8(s_addr + (tail0 >> 1), pk);
        }
        sum = (s0 + s1) + (s2 + s3);
        tmem_st_wait();
      }
      // a warp without any valid query row leaves S (all zeros: its Q rows are out of bounds) as "P"; its O rows are
      // never stored
      tc_fence_before();
      mbar_arrive(p_ready(slot));
      ATT_STAMP();
      mbar_wait(o_full(slot), par);
      tc_fence_after();
      ATT_STAMP();
      // O drain in two phases: (1) TMEM -> registers (scaled by 1/sum, packed to bf16: 64 registers for the 128 columns),
      // double-buffered loads; the TMEM slot is released as soon as the last load has landed, so the Q K^T of the item
      // after next starts while (2) the rows are staged and stored.  The slot's critical chain
      // QK -> softmax -> PV -> drain loses the whole store phase.
      uint32_t ob[64];
      if (warp_valid) {
        const float inv = row_valid ? 1.0f / sum : 0.0f;
        uint32_t va[32], vb[32];
        auto pack32 = [&](const uint32_t (&v)[32], int c) {
#pragma unroll
          for (int k = 0; k < 32; k += 2)
            ob[(c >> 1) + (k >> 1)] = pack_bf16x2(__uint_as_float(v[k]) * inv, __uint_as_float(v[k + 1]) * inv);
        };
        tmem_ld32(s_addr + (uint32_t)p_cols, va);
        tmem_ld_wait();
        tmem_ld32(s_addr + (uint32_t)(p_cols + 32), vb);
        pack32(va, 0);
        tmem_ld_wait();
        tmem_ld32(s_addr + (uint32_t)(p_cols + 64), va);
        pack32(vb, 32);
        tmem_ld_wait();
        tmem_ld32(s_addr + (uint32_t)(p_cols + 96), vb);
        pack32(va, 64);
        tmem_ld_wait();
        pack32(vb, 96);
      }
      tc_fence_before();
      mbar_arrive(slot_free(slot));
      ATT_STAMP();
      if (warp_valid) {
#pragma unroll
        for (int c = 0; c < ATT_DH; c += 32) {
          const int bx = (c >> 5) & 1;
          if (lane == 0) bulk_wait_read_1();  // the store that last used this box has read it
          __syncwarp();
          const uint32_t row_smem = my_box + bx * ATT_OBOX_BYTES + lane * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // 8 columns -> one 16-byte piece; 64B swizzle: piece ^ ((row >> 1) & 3)
            const uint32_t* o = &ob[(c >> 1) + q * 4];
            sts128(row_smem + ((q ^ ((lane >> 1) & 3)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmap_o, my_box + bx * ATT_OBOX_BYTES, head * ATT_DH + c, qt * 128 + quad * 32, seq);
            bulk_commit_group();
          }
        }
      }
      ATT_STAMP();
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// Version 3 (default; MST_ATTN_V=2 selects the kernel above): SIXTEEN softmax warps.
// The v2 timeline (profiles/r02a_attention_v2_timeline.txt) is one serial chain per TMEM slot of ~9 k cycles:
// row max 1.0 k, exp2 + P write-back 3.6 k, wait for P V 0.6-1.6 k, O drain + store 1.8 k - with only four warps
// (one per TMEM lane quadrant) working on an item, and two items in flight.  Here every lane quadrant of an item
// is served by TWO warps that split the key columns (A: [0, a), B: [a, s_pad), a = 16 ceil(s_pad / 32)) and, for
// the drain, the 128 output columns; the two exchange the row maximum and the row sum through shared memory
// (named barrier of 64 threads).  Each phase of the chain is half as long and the SM's four schedulers see four
// warps each instead of two.
// TMEM slot (256 columns): S fp32 in [0, s_pad); P of warp A (bf16 pairs) over its own S columns [0, a/2), P of
// warp B in the slot's spare columns [256 - (s_pad - a)/2, 256) - B must not overwrite S columns A may still be
// reading - and O fp32 in [a/2, a/2 + 128), dead S columns once both warps have arrived on p_ready.
// The output leaves through a per-warp [32 x 32] transpose box with plain coalesced 16-byte stores (8 rows x 64
// bytes per instruction): no async-proxy fence, no bulk-group wait.
// ---------------------------------------------------------------------------
constexpr int ATT3_SM_WARPS = 16;
constexpr int ATT3_THREADS = (ATT3_SM_WARPS + 2) * 32;  // warps 0-15 softmax, 16 TMA producer, 17 MMA issuer
constexpr int ATT3_STAGING_BYTES = ATT3_SM_WARPS * ATT_OBOX_BYTES;
constexpr int ATT3_XCH_BYTES = 2 * 2 * 2 * 128 * 4;  // [max | sum][group][column half][row]
constexpr int ATT3_SMEM_BYTES = 1024 + 2 * ATT_Q_BYTES + 2 * ATT_KV_BYTES + ATT3_STAGING_BYTES + ATT3_XCH_BYTES + 256;

__device__ __forceinline__ void stg128_attn(void* ptr, uint4 v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(ATT3_THREADS, 1)
tc_attention3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     __nv_bfloat16* __restrict__ out, const AttnGeom g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t kv_box = (uint32_t)g.s_pad * 128u;  // bytes of one [s_pad x 64] box
  const uint32_t q_smem = base;                       // 2 buffers
  const uint32_t k_smem = q_smem + 2 * ATT_Q_BYTES;
  const uint32_t v_smem = k_smem + ATT_KV_BYTES;
  const uint32_t o_smem = v_smem + ATT_KV_BYTES;
  const uint32_t xch_smem = o_smem + ATT3_STAGING_BYTES;
  const uint32_t bar_base = xch_smem + ATT3_XCH_BYTES;
  const uint32_t k_full = bar_base, k_free = bar_base + 8, v_full = bar_base + 16, v_free = bar_base + 24;
  auto q_full = [&](int b) { return bar_base + 32 + 8 * b; };
  auto q_free = [&](int b) { return bar_base + 48 + 8 * b; };
  auto s_full = [&](int s) { return bar_base + 64 + 8 * s; };
  auto p_ready = [&](int s) { return bar_base + 80 + 8 * s; };
  auto o_full = [&](int s) { return bar_base + 96 + 8 * s; };
  auto slot_free = [&](int s) { return bar_base + 112 + 8 * s; };
  const uint32_t tmem_slot = bar_base + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  float* const xch = reinterpret_cast<float*>(base_ptr + (xch_smem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // contiguous, balanced range of work items of this CTA
  const int per = g.n_items / (int)gridDim.x, extra = g.n_items % (int)gridDim.x;
  const int i0 = (int)blockIdx.x * per + min((int)blockIdx.x, extra);
  const int i1 = i0 + per + ((int)blockIdx.x < extra ? 1 : 0);

  if (warp == ATT3_SM_WARPS && elect_one()) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_kv);
    mbar_init(k_full, 1);
    mbar_init(k_free, 1);
    mbar_init(v_full, 1);
    mbar_init(v_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(q_full(b), 1);
      mbar_init(q_free(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(p_ready(b), 256);
      mbar_init(o_full(b), 1);
      mbar_init(slot_free(b), 256);
    }
    fence_barrier_init();
  }
  if (warp == ATT3_SM_WARPS + 1) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // the QKV projection has completed
  const uint32_t tmem_base = *tmem_slot_ptr;
  // column split of an S row between the two warps of a lane quadrant
  const int a_cols = ((g.s_pad + 31) >> 5) << 4;               // warp A: [0, a_cols), warp B: [a_cols, s_pad)
  const int pb_col = ATT_SLOT_COLS - ((g.s_pad - a_cols) >> 1);  // first TMEM column of B's packed P
  const int o_col = a_cols >> 1;                                // first TMEM column of O

  if (warp == ATT3_SM_WARPS) {
    // ------------------------------------------------ TMA producer (as in v2)
    if (elect_one()) {
      int n_units = 0;
      auto load_q = [&](int it) {
        const int j = it - i0, qb = j & 1;
        const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
        const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
        mbar_wait(q_free(qb), ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(q_full(qb), ATT_Q_BYTES);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_3d(q_smem + qb * ATT_Q_BYTES + kc * 16384, &tmap_q, q_full(qb), head * ATT_DH + kc * 64, qt * 128, seq);
      };
      int q_loaded = i0;
      for (int it = i0; it < i1; ++it) {
        const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
        const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
        const bool new_unit = (it == i0) || qt == 0;
        if (new_unit) {
          mbar_wait(k_free, (n_units & 1) ^ 1);
          mbar_expect_tx(k_full, 2 * kv_box);
          for (int kc = 0; kc < 2; ++kc)
            tma_load_3d(k_smem + kc * kv_box, &tmap_kv, k_full, g.d_model + head * ATT_DH + kc * 64, 0, seq);
        }
        if (q_loaded <= it) load_q(q_loaded++);
        if (new_unit) {
          if (q_loaded == it + 1 && it + 1 < i1 && qt + 1 < g.n_qt) load_q(q_loaded++);
          mbar_wait(v_free, (n_units & 1) ^ 1);
          mbar_expect_tx(v_full, 2 * kv_box);
          for (int kc = 0; kc < 2; ++kc)
            tma_load_3d(v_smem + kc * kv_box, &tmap_kv, v_full, 2 * g.d_model + head * ATT_DH + kc * 64, 0, seq);
          ++n_units;
        }
      }
    }
  } else if (warp == ATT3_SM_WARPS + 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc_bf16(128, g.s_pad, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, ATT_DH, 1);
      int k_units = 0, v_units = 0;
      long long* dbg = (g.dbg && blockIdx.x == 0) ? g.dbg + 1024 : nullptr;
      int di = 0;
      auto is_new_unit = [&](int it) { return it == i0 || (it % g.n_qt) == 0; };
      auto is_last_of_unit = [&](int it) { return it == i1 - 1 || ((it + 1) % g.n_qt) == 0; };
      auto issue_qk = [&](int it) {
        const int j = it - i0, slot = j & 1, qb = j & 1;
        if (is_new_unit(it)) {
          mbar_wait(k_full, k_units & 1);
          ++k_units;
        }
        mbar_wait(q_full(qb), (j >> 1) & 1);
        mbar_wait(slot_free(slot), ((j >> 1) & 1) ^ 1);  // O of item j-2 has been read out of this slot
        tc_fence_after();
        ATT_STAMP();
        const uint32_t d = tmem_base + (uint32_t)(slot * ATT_SLOT_COLS);
#pragma unroll
        for (int ks = 0; ks < ATT_DH / 16; ++ks) {
          const int kc = ks >> 2, k4 = ks & 3;
          const uint64_t adesc = make_smem_desc_sw128(q_smem + qb * ATT_Q_BYTES + kc * 16384 + k4 * 32, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(k_smem + kc * kv_box + k4 * 32, 0, 1024);
          mma_bf16_ss(d, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
        }
        mma_commit(s_full(slot));
        mma_commit(q_free(qb));
        if (is_last_of_unit(it)) mma_commit(k_free);
      };
      auto issue_pv = [&](int it) {
        const int j = it - i0, slot = j & 1;
        mbar_wait(p_ready(slot), (j >> 1) & 1);
        if (is_new_unit(it)) {
          mbar_wait(v_full, v_units & 1);
          ++v_units;
        }
        tc_fence_after();
        ATT_STAMP();
        const uint32_t slot_base = tmem_base + (uint32_t)(slot * ATT_SLOT_COLS);
        const int n_ks = g.s_pad / 16, ks_a = a_cols / 16;
        for (int ks = 0; ks < n_ks; ++ks) {
          // A = P: 16 keys = 8 packed columns per step, warp A's block first, then warp B's
          const uint32_t p_addr = ks < ks_a ? slot_base + (uint32_t)(ks * 8) : slot_base + (uint32_t)(pb_col + (ks - ks_a) * 8);
          const uint64_t bdesc = make_smem_desc_sw128(v_smem + ks * 2048, kv_box, 1024);
          mma_bf16_ts(slot_base + (uint32_t)o_col, p_addr, bdesc, idesc_pv, ks != 0 ? 1u : 0u);
        }
        mma_commit(o_full(slot));
        if (is_last_of_unit(it)) mma_commit(v_free);
      };
      if (i0 < i1) issue_qk(i0);
      for (int it = i0; it < i1; ++it) {
        if (it + 1 < i1) issue_qk(it + 1);
        issue_pv(it);
      }
    }
  } else {
    // ------------------------------------------------ softmax / output warps
    const int grp = warp >> 3;        // 0: even items, 1: odd items
    const int hcol = (warp >> 2) & 1;  // 0: warp A (columns [0, a)), 1: warp B
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t my_box = o_smem + (uint32_t)warp * ATT_OBOX_BYTES;
    const int pair_bar = 1 + grp * 4 + quad;  // named barrier of this quadrant's two warps
    float* const xch_max = xch + (grp * 2) * 128;            // [hcol][row]
    float* const xch_sum = xch + 512 + (grp * 2) * 128;
    const int c0 = hcol ? a_cols : 0, c1 = hcol ? g.s_pad : a_cols;
    const int n_ch = (c1 - c0) >> 4;                         // chunks of 16 columns (at most 7)
    const int p_col0 = hcol ? pb_col : 0;                    // first packed-P column of this warp
    long long* dbg = (g.dbg && blockIdx.x == 0 && quad == 0 && hcol == 0 && lane == 0) ? g.dbg + (2 + grp) * 1024 : nullptr;
    int di = 0;
    for (int it = i0 + grp; it < i1; it += 2) {
      const int j = it - i0, slot = j & 1;
      const uint32_t par = (j >> 1) & 1;
      const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
      const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
      const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(slot * ATT_SLOT_COLS);
      const bool row_valid = qt * 128 + r < g.S;
      const bool warp_valid = qt * 128 + quad * 32 < g.S;  // any valid row in this warp (same for both warps of a quadrant)
      ATT_STAMP();
      mbar_wait(s_full(slot), par);
      tc_fence_after();
      ATT_STAMP();
      float sum = 0.0f;
      if (warp_valid) {
        // 16 columns at a time, double-buffered TMEM loads (with 18 warps a thread has 96 registers: 32-column chunks
        // spill); every chunk but the one that holds key S is entirely valid or entirely padding
        uint32_t buf[2][16];
        auto chunk_nv = [&](int i) { return max(0, min(16, g.S - (c0 + 16 * i))); };
        // ---- pass 1: row maximum over this warp's columns
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        tmem_ld16(s_addr + (uint32_t)c0, buf[0]);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          if (i < n_ch) {
            tmem_ld_wait();
            if (i + 1 < n_ch) tmem_ld16(s_addr + (uint32_t)(c0 + 16 * (i + 1)), buf[(i + 1) & 1]);
            const uint32_t(&v)[16] = buf[i & 1];
            const int nv = chunk_nv(i);
            if (nv == 16) {
#pragma unroll
              for (int k = 0; k < 16; k += 4) {
                m0 = fmaxf(m0, __uint_as_float(v[k]));
                m1 = fmaxf(m1, __uint_as_float(v[k + 1]));
                m2 = fmaxf(m2, __uint_as_float(v[k + 2]));
                m3 = fmaxf(m3, __uint_as_float(v[k + 3]));
              }
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (k < nv) m0 = fmaxf(m0, __uint_as_float(v[k]));
            }
          }
        }
        ATT_STAMP();
        float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        xch_max[hcol * 128 + r] = mx;
        named_bar_sync(pair_bar, 64);
        mx = fmaxf(mx, xch_max[(hcol ^ 1) * 128 + r]);
        const float moff = mx * g.scale_log2e;
        // ---- pass 2: p = exp2((s - max) * scale), row sum, bf16 P into this warp's P columns
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        tmem_ld16(s_addr + (uint32_t)c0, buf[0]);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          if (i < n_ch) {
            tmem_ld_wait();
            if (i + 1 < n_ch) tmem_ld16(s_addr + (uint32_t)(c0 + 16 * (i + 1)), buf[(i + 1) & 1]);
            const uint32_t(&v)[16] = buf[i & 1];
            const int nv = chunk_nv(i);
            uint32_t pk[8];
            if (nv == 16) {
#pragma unroll
              for (int k = 0; k < 16; k += 4) {
                const float e0 = fast_exp2(fmaf(__uint_as_float(v[k]), g.scale_log2e, -moff));
                const float e1 = fast_exp2(fmaf(__uint_as_float(v[k + 1]), g.scale_log2e, -moff));
                const float e2 = fast_exp2(fmaf(__uint_as_float(v[k + 2]), g.scale_log2e, -moff));
                const float e3 = fast_exp2(fmaf(__uint_as_float(v[k + 3]), g.scale_log2e, -moff));
                s0 += e0; s1 += e1; s2 += e2; s3 += e3;
                pk[k >> 1] = pack_bf16x2(e0, e1);
                pk[(k >> 1) + 1] = pack_bf16x2(e2, e3);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 16; k += 2) {
                const float e0 = (k < nv) ? fast_exp2(fmaf(__uint_as_float(v[k]), g.scale_log2e, -moff)) : 0.0f;
                const float e1 = (k + 1 < nv) ? fast_exp2(fmaf(__uint_as_float(v[k + 1]), g.scale_log2e, -moff)) : 0.0f;
                s0 += e0; s1 += e1;
                pk[k >> 1] = pack_bf16x2(e0, e1);
              }
            }
            tmem_st8(s_addr + (uint32_t)(p_col0 + 8 * i), pk);
          }
        }
        sum = (s0 + s1) + (s2 + s3);
        xch_sum[hcol * 128 + r] = sum;
        tmem_st_wait();
      }
      // a quadrant without any valid query row leaves its P columns as they are; its O rows are never stored
      tc_fence_before();
      mbar_arrive(p_ready(slot));
      ATT_STAMP();
      mbar_wait(o_full(slot), par);
      tc_fence_after();
      ATT_STAMP();
      // O drain: this warp's 64 of the 128 columns -> registers (scaled by 1 / row sum, packed bf16); the TMEM slot is
      // released as soon as both loads have landed
      uint32_t ob[32];
      if (warp_valid) {
        named_bar_sync(pair_bar, 64);  // the partner's partial row sum is in shared memory
        const float total = sum + xch_sum[(hcol ^ 1) * 128 + r];
        const float inv = row_valid ? 1.0f / total : 0.0f;
        const uint32_t o_addr = s_addr + (uint32_t)(o_col + hcol * 64);
        uint32_t va[16], vb[16];
        tmem_ld16(o_addr, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 16 columns at a time, double-buffered
          tmem_ld_wait();
          if (c < 3) tmem_ld16(o_addr + (uint32_t)((c + 1) * 16), (c & 1) ? va : vb);
          const uint32_t(&v)[16] = (c & 1) ? vb : va;
#pragma unroll
          for (int k = 0; k < 16; k += 2)
            ob[c * 8 + (k >> 1)] = pack_bf16x2(__uint_as_float(v[k]) * inv, __uint_as_float(v[k + 1]) * inv);
        }
      }
      tc_fence_before();
      mbar_arrive(slot_free(slot));
      ATT_STAMP();
      if (warp_valid) {
        const int row_base = qt * 128 + quad * 32;  // first query row of this warp inside the sequence
        __nv_bfloat16* const o_base = out + ((size_t)seq * g.S + row_base) * g.d_model + head * ATT_DH + hcol * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // 32 columns at a time through the warp's [32 x 32] transpose box
          const uint32_t row_smem = my_box + lane * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t* o = &ob[c * 16 + q * 4];
            sts128(row_smem + ((q ^ ((lane >> 1) & 3)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
          }
          __syncwarp();
          const int piece = lane & 3;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = k * 8 + (lane >> 2);
            const uint4 val = lds128(my_box + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
            if (row_base + rr < g.S) stg128_attn(o_base + (size_t)rr * g.d_model + c * 32 + piece * 8, val);
          }
          __syncwarp();
        }
      }
      ATT_STAMP();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATT3_SM_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// Version 4 (default): ALL eight softmax warps work on ONE item at a time, two per TMEM lane quadrant (the column
// split and TMEM layout of v3), and the O drain of item i is deferred until the P of item i+1 has been handed to the
// tensor core.  Why (profiles/r02c_ncu_attention_v2_*.txt, warp-state samples of v2): its softmax warps wait 48 % of
// their time - 23 % for S, 25 % for O - because the two groups run in lock-step: both hand over P together, the two
// P V run back to back, both drain together, and the tensor pipe (26 % busy) idles during both exp2 passes.  One
// warp per scheduler also cannot overlap its own MUFU (8 cycles per instruction) with its FFMA / FADD / F2FP:
// 3.6 k cycles per exp2 pass against 1.7 k of SFU time.  Here
//   softmax(i):  wait S(i) -> row max (exchange) -> exp2, P -> hand over P(i)        [P V(i) starts]
//                then drain O(i-1) (complete long ago), free its slot               [Q K^T(i+1) starts], store O(i-1)
// so the warps never wait for an MMA they have just triggered, two warps share every scheduler, and a thread keeps
// 200 registers (10 warps per CTA; v3's 18 warps left 96 and spilled).
// ---------------------------------------------------------------------------
constexpr int ATT4_THREADS = 10 * 32;  // warps 0-7 softmax, 8 TMA producer, 9 MMA issuer
constexpr int ATT4_STAGING_BYTES = 8 * ATT_OBOX_BYTES;
constexpr int ATT4_XCH_BYTES = 2 * 2 * 2 * 128 * 4;  // [max | sum][item parity][column half][row]
constexpr int ATT4_SMEM_BYTES = 1024 + 2 * ATT_Q_BYTES + 2 * ATT_KV_BYTES + ATT4_STAGING_BYTES + ATT4_XCH_BYTES + 256;

__global__ void __launch_bounds__(ATT4_THREADS, 1)
tc_attention4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     __nv_bfloat16* __restrict__ out, const AttnGeom g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t kv_box = (uint32_t)g.s_pad * 128u;
  const uint32_t q_smem = base;
  const uint32_t k_smem = q_smem + 2 * ATT_Q_BYTES;
  const uint32_t v_smem = k_smem + ATT_KV_BYTES;
  const uint32_t o_smem = v_smem + ATT_KV_BYTES;
  const uint32_t xch_smem = o_smem + ATT4_STAGING_BYTES;
  const uint32_t bar_base = xch_smem + ATT4_XCH_BYTES;
  const uint32_t k_full = bar_base, k_free = bar_base + 8, v_full = bar_base + 16, v_free = bar_base + 24;
  auto q_full = [&](int b) { return bar_base + 32 + 8 * b; };
  auto q_free = [&](int b) { return bar_base + 48 + 8 * b; };
  auto s_full = [&](int s) { return bar_base + 64 + 8 * s; };
  auto p_ready = [&](int s) { return bar_base + 80 + 8 * s; };
  auto o_full = [&](int s) { return bar_base + 96 + 8 * s; };
  auto slot_free = [&](int s) { return bar_base + 112 + 8 * s; };
  const uint32_t tmem_slot = bar_base + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  float* const xch = reinterpret_cast<float*>(base_ptr + (xch_smem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = g.n_items / (int)gridDim.x, extra = g.n_items % (int)gridDim.x;
  const int i0 = (int)blockIdx.x * per + min((int)blockIdx.x, extra);
  const int i1 = i0 + per + ((int)blockIdx.x < extra ? 1 : 0);

  if (warp == 8 && elect_one()) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_kv);
    mbar_init(k_full, 1);
    mbar_init(k_free, 1);
    mbar_init(v_full, 1);
    mbar_init(v_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(q_full(b), 1);
      mbar_init(q_free(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(p_ready(b), 256);
      mbar_init(o_full(b), 1);
      mbar_init(slot_free(b), 256);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // the QKV projection has completed
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int a_cols = ((g.s_pad + 31) >> 5) << 4;                 // warp A: [0, a_cols), warp B: [a_cols, s_pad)
  const int pb_col = ATT_SLOT_COLS - ((g.s_pad - a_cols) >> 1);  // first TMEM column of B's packed P
  const int o_col = a_cols >> 1;                                  // first TMEM column of O

  if (warp == 8) {
    // ------------------------------------------------ TMA producer (as in v2)
    if (elect_one()) {
      int n_units = 0;
      auto load_q = [&](int it) {
        const int j = it - i0, qb = j & 1;
        const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
        const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
        mbar_wait(q_free(qb), ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(q_full(qb), ATT_Q_BYTES);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_3d(q_smem + qb * ATT_Q_BYTES + kc * 16384, &tmap_q, q_full(qb), head * ATT_DH + kc * 64, qt * 128, seq);
      };
      int q_loaded = i0;
      for (int it = i0; it < i1; ++it) {
        const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
        const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
        const bool new_unit = (it == i0) || qt == 0;
        if (new_unit) {
          mbar_wait(k_free, (n_units & 1) ^ 1);
          mbar_expect_tx(k_full, 2 * kv_box);
          for (int kc = 0; kc < 2; ++kc)
            tma_load_3d(k_smem + kc * kv_box, &tmap_kv, k_full, g.d_model + head * ATT_DH + kc * 64, 0, seq);
        }
        if (q_loaded <= it) load_q(q_loaded++);
        if (new_unit) {
          if (q_loaded == it + 1 && it + 1 < i1 && qt + 1 < g.n_qt) load_q(q_loaded++);
          mbar_wait(v_free, (n_units & 1) ^ 1);
          mbar_expect_tx(v_full, 2 * kv_box);
          for (int kc = 0; kc < 2; ++kc)
            tma_load_3d(v_smem + kc * kv_box, &tmap_kv, v_full, 2 * g.d_model + head * ATT_DH + kc * 64, 0, seq);
          ++n_units;
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc_bf16(128, g.s_pad, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, ATT_DH, 1);
      int k_units = 0, v_units = 0;
      long long* dbg = (g.dbg && blockIdx.x == 0) ? g.dbg + 1024 : nullptr;
      int di = 0;
      auto is_new_unit = [&](int it) { return it == i0 || (it % g.n_qt) == 0; };
      auto is_last_of_unit = [&](int it) { return it == i1 - 1 || ((it + 1) % g.n_qt) == 0; };
      auto issue_qk = [&](int it) {
        const int j = it - i0, slot = j & 1, qb = j & 1;
        if (is_new_unit(it)) {
          mbar_wait(k_full, k_units & 1);
          ++k_units;
        }
        mbar_wait(q_full(qb), (j >> 1) & 1);
        mbar_wait(slot_free(slot), ((j >> 1) & 1) ^ 1);  // O of item j-2 has been read out of this slot
        tc_fence_after();
        ATT_STAMP();
        const uint32_t d = tmem_base + (uint32_t)(slot * ATT_SLOT_COLS);
#pragma unroll
        for (int ks = 0; ks < ATT_DH / 16; ++ks) {
          const int kc = ks >> 2, k4 = ks & 3;
          const uint64_t adesc = make_smem_desc_sw128(q_smem + qb * ATT_Q_BYTES + kc * 16384 + k4 * 32, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(k_smem + kc * kv_box + k4 * 32, 0, 1024);
          mma_bf16_ss(d, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
        }
        mma_commit(s_full(slot));
        mma_commit(q_free(qb));
        if (is_last_of_unit(it)) mma_commit(k_free);
      };
      auto issue_pv = [&](int it) {
        const int j = it - i0, slot = j & 1;
        mbar_wait(p_ready(slot), (j >> 1) & 1);
        if (is_new_unit(it)) {
          mbar_wait(v_full, v_units & 1);
          ++v_units;
        }
        tc_fence_after();
        ATT_STAMP();
        const uint32_t slot_base = tmem_base + (uint32_t)(slot * ATT_SLOT_COLS);
        const int n_ks = g.s_pad / 16, ks_a = a_cols / 16;
        for (int ks = 0; ks < n_ks; ++ks) {
          const uint32_t p_addr = ks < ks_a ? slot_base + (uint32_t)(ks * 8) : slot_base + (uint32_t)(pb_col + (ks - ks_a) * 8);
          const uint64_t bdesc = make_smem_desc_sw128(v_smem + ks * 2048, kv_box, 1024);
          mma_bf16_ts(slot_base + (uint32_t)o_col, p_addr, bdesc, idesc_pv, ks != 0 ? 1u : 0u);
        }
        mma_commit(o_full(slot));
        if (is_last_of_unit(it)) mma_commit(v_free);
      };
      // Q K^T runs two items ahead of the softmax: S(i+1) is complete before the warps finish item i, and the
      // slot of item i is re-used for S(i+2) as soon as O(i) has been drained (right after P(i+1) was handed over)
      if (i0 < i1) issue_qk(i0);
      if (i0 + 1 < i1) issue_qk(i0 + 1);
      for (int it = i0; it < i1; ++it) {
        issue_pv(it);
        if (it + 2 < i1) issue_qk(it + 2);
      }
    }
  } else {
    // ------------------------------------------------ softmax / output warps: two per lane quadrant
    const int hcol = warp >> 2;  // 0: warp A (columns [0, a)), 1: warp B
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t my_box = o_smem + (uint32_t)warp * ATT_OBOX_BYTES;
    const int pair_bar = 1 + quad;  // named barrier of this quadrant's two warps
    float* const xch_max = xch;           // [item parity][hcol][row]
    float* const xch_sum = xch + 512;
    const int c0 = hcol ? a_cols : 0, c1 = hcol ? g.s_pad : a_cols;
    const int n_ch = (c1 - c0 + 31) >> 5;  // chunks of 32 columns; the last one may be 16 wide
    const int p_col0 = hcol ? pb_col : 0;
    long long* dbg = (g.dbg && blockIdx.x == 0 && warp == 0 && lane == 0) ? g.dbg + 2 * 1024 : nullptr;
    int di = 0;

    // drain + store of a finished item (its P V was issued one softmax earlier)
    auto drain_store = [&](int it, float sum_mine) {
      const int j = it - i0, slot = j & 1;
      const uint32_t par = (j >> 1) & 1;
      const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
      const int seq = unit / g.n_heads, head = unit - seq * g.n_heads;
      const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(slot * ATT_SLOT_COLS);
      const bool row_valid = qt * 128 + r < g.S;
      const bool warp_valid = qt * 128 + quad * 32 < g.S;
      mbar_wait(o_full(slot), par);
      tc_fence_after();
      uint32_t ob[32];
      if (warp_valid) {
        const float total = sum_mine + xch_sum[((j & 1) * 2 + (hcol ^ 1)) * 128 + r];
        const float inv = row_valid ? 1.0f / total : 0.0f;
        const uint32_t o_addr = s_addr + (uint32_t)(o_col + hcol * 64);
        uint32_t va[32], vb[32];
        tmem_ld32(o_addr, va);
        tmem_ld32(o_addr + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          ob[k >> 1] = pack_bf16x2(__uint_as_float(va[k]) * inv, __uint_as_float(va[k + 1]) * inv);
          ob[16 + (k >> 1)] = pack_bf16x2(__uint_as_float(vb[k]) * inv, __uint_as_float(vb[k + 1]) * inv);
        }
      }
      tc_fence_before();
      mbar_arrive(slot_free(slot));
      if (warp_valid) {
        const int row_base = qt * 128 + quad * 32;
        __nv_bfloat16* const o_base = out + ((size_t)seq * g.S + row_base) * g.d_model + head * ATT_DH + hcol * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // 32 columns at a time through the warp's [32 x 32] transpose box
          const uint32_t row_smem = my_box + lane * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t* o = &ob[c * 16 + q * 4];
            sts128(row_smem + ((q ^ ((lane >> 1) & 3)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
          }
          __syncwarp();
          const int piece = lane & 3;
          uint4 val[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = k * 8 + (lane >> 2);
            val[k] = lds128(my_box + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = k * 8 + (lane >> 2);
            if (row_base + rr < g.S) stg128_attn(o_base + (size_t)rr * g.d_model + c * 32 + piece * 8, val[k]);
          }
          __syncwarp();
        }
      }
    };

    float prev_sum = 0.0f;
    for (int it = i0; it < i1; ++it) {
      const int j = it - i0, slot = j & 1;
      const uint32_t par = (j >> 1) & 1;
      const int unit = it / g.n_qt, qt = it - unit * g.n_qt;
      const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(slot * ATT_SLOT_COLS);
      const bool warp_valid = qt * 128 + quad * 32 < g.S;
      ATT_STAMP();
      mbar_wait(s_full(slot), par);
      tc_fence_after();
      ATT_STAMP();
      float sum = 0.0f;
      if (warp_valid) {
        uint32_t buf[2][32];
        auto chunk_w = [&](int i) { return min(32, c1 - (c0 + 32 * i)); };
        auto chunk_nv = [&](int i) { return max(0, min(chunk_w(i), g.S - (c0 + 32 * i))); };
        auto load_chunk = [&](int i, uint32_t (&v)[32]) {
          if (chunk_w(i) == 32) {
            tmem_ld32(s_addr + (uint32_t)(c0 + 32 * i), v);
          } else {
            uint32_t t[16];
            tmem_ld16(s_addr + (uint32_t)(c0 + 32 * i), t);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = t[k];
          }
        };
        // ---- pass 1: row maximum over this warp's columns
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        if (n_ch > 0) load_chunk(0, buf[0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < n_ch) {
            tmem_ld_wait();
            if (i + 1 < n_ch) load_chunk(i + 1, buf[(i + 1) & 1]);
            const uint32_t(&v)[32] = buf[i & 1];
            const int nv = chunk_nv(i);
            if (nv == 32) {
#pragma unroll
              for (int k = 0; k < 32; k += 4) {
                m0 = fmaxf(m0, __uint_as_float(v[k]));
                m1 = fmaxf(m1, __uint_as_float(v[k + 1]));
                m2 = fmaxf(m2, __uint_as_float(v[k + 2]));
                m3 = fmaxf(m3, __uint_as_float(v[k + 3]));
              }
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (k < nv) m0 = fmaxf(m0, __uint_as_float(v[k]));
            }
          }
        }
        ATT_STAMP();
        float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        xch_max[((j & 1) * 2 + hcol) * 128 + r] = mx;
        named_bar_sync(pair_bar, 64);
        mx = fmaxf(mx, xch_max[((j & 1) * 2 + (hcol ^ 1)) * 128 + r]);
        const float moff = mx * g.scale_log2e;
        // ---- pass 2: p = exp2((s - max) * scale), row sum, bf16 P into this warp's P columns
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        if (n_ch > 0) load_chunk(0, buf[0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < n_ch) {
            tmem_ld_wait();
            if (i + 1 < n_ch) load_chunk(i + 1, buf[(i + 1) & 1]);
            const uint32_t(&v)[32] = buf[i & 1];
            const int nv = chunk_nv(i), w = chunk_w(i);
            const uint32_t p_addr = s_addr + (uint32_t)(p_col0 + 16 * i);
            uint32_t pk[16];
            if (nv == 32) {
#pragma unroll
              for (int k = 0; k < 32; k += 4) {
                const float e0 = fast_exp2(fmaf(__uint_as_float(v[k]), g.scale_log2e, -moff));
                const float e1 = fast_exp2(fmaf(__uint_as_float(v[k + 1]), g.scale_log2e, -moff));
                const float e2 = fast_exp2(fmaf(__uint_as_float(v[k + 2]), g.scale_log2e, -moff));
                const float e3 = fast_exp2(fmaf(__uint_as_float(v[k + 3]), g.scale_log2e, -moff));
                s0 += e0; s1 += e1; s2 += e2; s3 += e3;
                pk[k >> 1] = pack_bf16x2(e0, e1);
                pk[(k >> 1) + 1] = pack_bf16x2(e2, e3);
              }
              tmem_st16(p_addr, pk);
            } else {
#pragma unroll
              for (int k = 0; k < 32; k += 2) {
                const float e0 = (k < nv) ? fast_exp2(fmaf(__uint_as_float(v[k]), g.scale_log2e, -moff)) : 0.0f;
                const float e1 = (k + 1 < nv) ? fast_exp2(fmaf(__uint_as_float(v[k + 1]), g.scale_log2e, -moff)) : 0.0f;
                s0 += e0; s1 += e1;
                pk[k >> 1] = pack_bf16x2(e0, e1);
              }
              if (w == 32) {
                tmem_st16(p_addr, pk);
              } else {
                uint32_t pk8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) pk8[k] = pk[k];
                tmem_st8(p_addr, pk8);
              }
            }
          }
        }
        sum = (s0 + s1) + (s2 + s3);
        xch_sum[((j & 1) * 2 + hcol) * 128 + r] = sum;  // read by the partner when it drains this item (slot parity)
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(p_ready(slot));
      ATT_STAMP();
      // the previous item's O has been complete for a whole softmax: drain it now; its slot then takes S(it + 1)
      if (it > i0) {
        named_bar_sync(pair_bar, 64);  // partner's partial row sums (of item it-1, written one item ago) are visible
        drain_store(it - 1, prev_sum);
      }
      ATT_STAMP();
      prev_sum = sum;
    }
    if (i1 > i0) {
      named_bar_sync(pair_bar, 64);
      drain_store(i1 - 1, prev_sum);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

long long* attn_debug_ptr();  // tc_gemm.cu: the buffer set by mst_test_set_gemm_debug

int tc_attention(const TcAttnParams& p, cudaStream_t s) {
  MST_CHECK_ARG(p.qkv && p.out, "null pointer");
  MST_CHECK_ARG(p.n_seqs > 0 && p.S > 0, "empty problem");
  MST_CHECK_ARG(p.d_model == p.n_heads * ATT_DH, "attention kernel is built for head_dim 128");
  const int s_pad = (p.S + 15) / 16 * 16;
  if (s_pad > ATT_MAX_SPAD)
    return fail(MST_ERR_UNSUPPORTED, "tc_attention: sequences longer than 208 tokens (T > 207) are not supported");
  const uint64_t ld = 3 * (uint64_t)p.d_model;
  CUtensorMap tq, tkv, to;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tq, p.qkv, p.n_seqs, p.S, ld, ld, (uint64_t)p.S * ld, 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tkv, p.qkv, p.n_seqs, p.S, ld, ld, (uint64_t)p.S * ld, (uint32_t)s_pad, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&to, p.out, p.n_seqs, p.S, p.d_model, p.d_model, (uint64_t)p.S * p.d_model, 32, 32, 64)))
    return rc;
  static PerDeviceOnce attr_set;  // cudaFuncSetAttribute is per device
  if (attr_set.first()) {
    MST_CUDA_OK(cudaFuncSetAttribute(tc_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
  }
  AttnGeom g;
  g.S = p.S;
  g.s_pad = s_pad;
  g.n_qt = (p.S + 127) / 128;
  g.n_heads = p.n_heads;
  g.n_items = p.n_seqs * p.n_heads * g.n_qt;
  g.d_model = p.d_model;
  g.scale_log2e = 1.4426950408889634f / sqrtf((float)ATT_DH);
  g.dbg = attn_debug_ptr();
  const int grid = g.n_items < sm_count() ? g.n_items : sm_count();
  static const int attn_v = getenv("MST_ATTN_V") ? atoi(getenv("MST_ATTN_V")) : 2;
  if (attn_v == 4) {
    static PerDeviceOnce attr4_set;
    if (attr4_set.first()) {
      MST_CUDA_OK(cudaFuncSetAttribute(tc_attention4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES));
    }
    MST_CUDA_OK(launch_pdl(tc_attention4_kernel, dim3(grid), dim3(ATT4_THREADS), ATT4_SMEM_BYTES, s, tq, tkv, p.out, g));
    MST_LAUNCHED("tc_attention", s);
    return MST_OK;
  }
  if (attn_v == 3) {
    static PerDeviceOnce attr3_set;
    if (attr3_set.first()) {
      MST_CUDA_OK(cudaFuncSetAttribute(tc_attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT3_SMEM_BYTES));
    }
    MST_CUDA_OK(launch_pdl(tc_attention3_kernel, dim3(grid), dim3(ATT3_THREADS), ATT3_SMEM_BYTES, s, tq, tkv, p.out, g));
    MST_LAUNCHED("tc_attention", s);
    return MST_OK;
  }
  MST_CUDA_OK(launch_pdl(tc_attention_kernel, dim3(grid), dim3(ATT_THREADS), ATT_SMEM_BYTES, s, tq, tkv, to, g));
  MST_LAUNCHED("tc_attention", s);
  return MST_OK;
}

}  // namespace mst
