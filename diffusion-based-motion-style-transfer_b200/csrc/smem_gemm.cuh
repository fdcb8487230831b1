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

}  // namespace mst
