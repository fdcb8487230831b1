// Counter-based dropout masks of the training path (train.cu; the GELU epilogue of the training GEMM in tc_gemm.cu).
#pragma once
#include <stdint.h>

namespace mst {

// ---------------------------------------------------------------------------------------------------------
// Dropout of the training forward (nn.TransformerEncoderLayer(dropout=0.1) and PositionalEncoding(dropout=0.1) are
// active once the reference calls model.train(), train/finetune_style_diffusion.py:256).  Counter-based: the keep /
// drop decision of element i at site s is Philox4x32-10(key = *seed_dev, counter = (i / 4, site, 0, 0)) lane i % 4
// < p, so the backward recomputes the mask instead of storing it, and a CUDA-graph replay picks up a new seed from
// device memory.  Sites: 0 = token sequence after the positional encoding; 8 * (layer + 1) + {1: attention
// probabilities, 2: out-proj output, 3: GELU output, 4: linear2 output}.
// ---------------------------------------------------------------------------------------------------------
struct Drop {
  float p = 0.0f;                      // 0 = off
  const unsigned long long* seed = nullptr;  // device: one key per sequence of the call (the masks of a sequence do not
                                             // depend on which other sequences share the launch, so forwards recorded one
                                             // by one can be back-propagated as one batch)
  int n_seqs = 1;
  __host__ __device__ bool on() const { return p > 0.0f; }
};

__device__ __forceinline__ void philox4(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t x = c0, y = c1, z = 0u, w = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, x), lo0 = M0 * x, hi1 = __umulhi(M1, z), lo1 = M1 * z;
    const uint32_t nx = hi1 ^ y ^ k0, nz = hi0 ^ w ^ k1;
    x = nx; y = lo1; z = nz; w = lo0;
    k0 += W0; k1 += W1;
  }
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

// scale factors (0 or 1/(1-p)) of elements [4v, 4v+3] of `site` of sequence `seq` (v counts inside the sequence)
__device__ __forceinline__ float4 drop_scale4(const Drop& d, uint32_t site, int seq, long long v) {
  const unsigned long long seed = d.seed[seq];
  uint32_t r[4];
  philox4((uint32_t)v, site ^ ((uint32_t)(v >> 32) << 16), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float keep = 1.0f / (1.0f - d.p);
  const uint32_t thr = (uint32_t)(d.p * 4294967296.0);  // drop when r < thr
  return make_float4(r[0] < thr ? 0.0f : keep, r[1] < thr ? 0.0f : keep, r[2] < thr ? 0.0f : keep, r[3] < thr ? 0.0f : keep);
}

}  // namespace mst
