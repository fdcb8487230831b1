// Shared-memory register-tiled fp32 product used by the one-CTA-per-(sequence, head) attention kernels
// (train.cu: denoiser training attention; text.cu: causal attention of the CLIP text tower).
#pragma once

namespace mst {

// C(m, n) = sum_k A(m, k) B(k, n) for m < 16*TM, n < 16*TN (operands zero-padded in shared memory);
// A(m,k) = TA ? a[k*lda + m] : a[m*lda + k];  B(k,n) = TB ? b[n*ldb + k] : b[k*ldb + n];  thread (ty, tx) owns rows
// ty + 16 i and columns tx + 16 j and hands every result to `out(m, n, value)`.
template <bool TA, bool TB, int TM, int TN, typename Out>
__device__ __forceinline__ void smem_gemm(const float* a, int lda, const float* b, int ldb, int K, Out out) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float av[TM], bv[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) av[i] = TA ? a[k * lda + ty + 16 * i] : a[(ty + 16 * i) * lda + k];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = TB ? b[(tx + 16 * j) * ldb + k] : b[k * ldb + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) out(ty + 16 * i, tx + 16 * j, acc[i][j]);
}

// ---- packed variants (fma.rn.f32x2: two fp32 FMAs per instruction; 128-bit shared-memory loads) ---------------------
// The scalar template above issues (TM + TN) LDS.32 + TM*TN FFMA per k.  These two cut the instruction count of the
// training attention kernels by a third (ncu: 47.3 M -> 30.1 M warp instructions in the B=64 backward, shared-memory bank
// conflicts 1.08 M -> 0.12 M); the kernels stay latency-bound at 8 warps per SM, so the time moved only 99 -> 92 us.
// Leading dimensions must be multiples of 4 floats (16-byte aligned rows).
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// C(m, n) = sum_k a[m*lda + k] * b[n*ldb + k]   (both operands K-contiguous: Q K^T, dO V^T); rows ty + 16 i, columns
// tx + 16 j; K % 4 == 0.  The two lanes of a packed accumulator hold the sums over even and over odd k.
template <int TM, int TN, typename Out>
__device__ __forceinline__ void smem_gemm_nt2(const float* a, int lda, const float* b, int ldb, int K, Out out) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  unsigned long long acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0ull;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    ulonglong2 av[TM], bv[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const ulonglong2*>(a + (ty + 16 * i) * lda + k);
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = *reinterpret_cast<const ulonglong2*>(b + (tx + 16 * j) * ldb + k);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        acc[i][j] = f2_fma(av[i].x, bv[j].x, acc[i][j]);
        acc[i][j] = f2_fma(av[i].y, bv[j].y, acc[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float lo, hi;
      f2_unpack(acc[i][j], lo, hi);
      out(ty + 16 * i, tx + 16 * j, lo + hi);
    }
}

// C(m, n) = sum_k A(m, k) * b[k*ldb + n] for 128 columns (B N-contiguous: P V, dS K, P^T dO, dS^T Q);
// A(m, k) = TA ? a[k*lda + m] : a[m*lda + k]; rows ty + 16 i, columns 4 tx + {0..3} and 64 + 4 tx + {0..3}; K % 4 == 0.
// Same summation order over k as the scalar template (results are bit-identical to it).
template <bool TA, int TM, typename Out>
__device__ __forceinline__ void smem_gemm_n128(const float* a, int lda, const float* b, int ldb, int K, Out out) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  unsigned long long acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
  for (int k = 0; k < K; k += 4) {
    float av[TM][4];
    if (TA) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i][kk] = a[(k + kk) * lda + ty + 16 * i];
    } else {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(a + (ty + 16 * i) * lda + k);
        av[i][0] = t.x; av[i][1] = t.y; av[i][2] = t.z; av[i][3] = t.w;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(b + (k + kk) * ldb + 4 * tx);
      const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(b + (k + kk) * ldb + 64 + 4 * tx);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const unsigned long long x = f2_pack(av[i][kk], av[i][kk]);
        acc[i][0] = f2_fma(x, b0.x, acc[i][0]);
        acc[i][1] = f2_fma(x, b0.y, acc[i][1]);
        acc[i][2] = f2_fma(x, b1.x, acc[i][2]);
        acc[i][3] = f2_fma(x, b1.y, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float c0, c1, c2, c3;
      f2_unpack(acc[i][2 * g], c0, c1);
      f2_unpack(acc[i][2 * g + 1], c2, c3);
      const int n = 64 * g + 4 * tx;
      out(ty + 16 * i, n, c0);
      out(ty + 16 * i, n + 1, c1);
      out(ty + 16 * i, n + 2, c2);
      out(ty + 16 * i, n + 3, c3);
    }
}

// ---- tensor-core variants (mma.sync m16n8k8, tf32 operands, fp32 accumulation) -----------------------------------------
// The packed SIMT products above are bound by the shared-memory pipe (every thread re-reads whole operand rows: ~40 k
// wavefront cycles per CTA in the training attention).  Warp-level MMA fragments read each operand element once per 16 x 8
// output tile: ~4 k cycles.  Used by the 16-bit training mode only (the fp32 parity mode keeps the exact fp32 products); tf32
// keeps 10 mantissa bits, more than the bf16 operands of the linear layers around it.  256 threads = 8 warps.
// Fragment layout (g = lane / 4, t = lane % 4):  A: (g, t) (g+8, t) (g, t+4) (g+8, t+4);  B: (k = t, n = g) (k = t+4, n = g);
// C: (g, 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1); results are handed out as pairs of adjacent columns: out(m, n, v_n, v_n1).
// operands go in as raw fp32 bits: the tensor core reads the upper 19 (sign, exponent, 10 mantissa bits), i.e. truncates;
// cvt.rna.tf32 would round, but it expands to three instructions per element here and tripled the kernels' instruction count
__device__ __forceinline__ uint32_t to_tf32(float x) { return __float_as_uint(x); }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// C(m, n) = sum_k a[m*lda + k] * b[n*ldb + k] for m < 16 MT, n < 80 (operands zero padded); K % 8 == 0.
// Work unit = one 16-row tile x two 8-column tiles; the 5 MT units go round the 8 warps.
template <int MT, int NW = 8, typename Out>
__device__ __forceinline__ void mma_gemm_nt80(const float* a, int lda, const float* b, int ldb, int K, Out out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int u = warp; u < MT * 5; u += NW) {
    const int mt = u / 5, n0 = (u - mt * 5) * 16;
    float c[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[j][e] = 0.0f;
    const float* ar = a + (16 * mt + g) * lda + t;
    const float* br = b + (n0 + g) * ldb + t;
#pragma unroll 4
    for (int k = 0; k < K; k += 8) {
      const uint32_t af[4] = {to_tf32(ar[k]), to_tf32(ar[8 * lda + k]), to_tf32(ar[k + 4]), to_tf32(ar[8 * lda + k + 4])};
#pragma unroll
      for (int j = 0; j < 2; ++j)
        mma_tf32(c[j], af, to_tf32(br[j * 8 * ldb + k]), to_tf32(br[j * 8 * ldb + k + 4]));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = 16 * mt + g, n = n0 + 8 * j + 2 * t;
      out(m, n, c[j][0], c[j][1]);  // columns n, n + 1
      out(m + 8, n, c[j][2], c[j][3]);
    }
  }
}

// C(m, n) = sum_k A(m, k) * b[k*ldb + n] for 128 columns, m < 16 MT; A(m, k) = TA ? a[k*lda + m] : a[m*lda + k]; K % 8 == 0.
// Warp w of NW (8 or 16) owns the 128 / NW columns from (128 / NW) w on, of all row tiles.
template <bool TA, int MT, int NW = 8, typename Out>
__device__ __forceinline__ void mma_gemm_n128(const float* a, int lda, const float* b, int ldb, int K, Out out) {
  constexpr int NTW = 16 / NW;  // 8-column tiles per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float c[MT][NTW][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[i][j][e] = 0.0f;
  const float* bc = b + 8 * NTW * warp + g;
#pragma unroll 2
  for (int k = 0; k < K; k += 8) {
    uint32_t bf[NTW][2];
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      bf[j][0] = to_tf32(bc[(k + t) * ldb + 8 * j]);
      bf[j][1] = to_tf32(bc[(k + t + 4) * ldb + 8 * j]);
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      uint32_t af[4];
      const int m = 16 * i + g;
      if (TA) {
        af[0] = to_tf32(a[(k + t) * lda + m]);
        af[1] = to_tf32(a[(k + t) * lda + m + 8]);
        af[2] = to_tf32(a[(k + t + 4) * lda + m]);
        af[3] = to_tf32(a[(k + t + 4) * lda + m + 8]);
      } else {
        af[0] = to_tf32(a[m * lda + k + t]);
        af[1] = to_tf32(a[(m + 8) * lda + k + t]);
        af[2] = to_tf32(a[m * lda + k + t + 4]);
        af[3] = to_tf32(a[(m + 8) * lda + k + t + 4]);
      }
#pragma unroll
      for (int j = 0; j < NTW; ++j) mma_tf32(c[i][j], af, bf[j][0], bf[j][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int m = 16 * i + g, n = 8 * NTW * warp + 8 * j + 2 * t;
      out(m, n, c[i][j][0], c[i][j][1]);  // columns n, n + 1
      out(m + 8, n, c[i][j][2], c[i][j][3]);
    }
}

}  // namespace mst
