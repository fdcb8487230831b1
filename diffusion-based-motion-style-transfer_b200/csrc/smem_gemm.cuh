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

}  // namespace mst
