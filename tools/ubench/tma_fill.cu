// How many bytes per clock can all SMs together pull into shared memory through the bulk-copy (TMA) path?
// One CTA per SM keeps STAGES bulk copies of STAGE_BYTES in flight (cp.async.bulk global -> shared, mbarrier completion)
// and re-issues a slot as soon as it has landed - the operand ring of the tcgen05 GEMMs without the MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_fill tools/ubench/tma_fill.cu && tools/ubench/tma_fill
// Source patterns: every CTA its own region (streams from HBM when the footprint exceeds L2), all CTAs the same small region
// (L2-resident), pairs / quads of CTAs the same region (what the clusters of the GEMMs do with their A / W tiles).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int STAGE_BYTES = 32 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_fill(const uint8_t* __restrict__ src, long long region_bytes, int share, int iters,
                                                 int stages, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint8_t* base = src + (long long)(blockIdx.x / share) * region_bytes;
  const long long n_chunks = region_bytes / STAGE_BYTES;
  long long chunk = (blockIdx.x % share) * 7;  // sharers walk the same region a few chunks apart, like CTAs of a cluster
  uint32_t phase = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters + stages; ++i) {
    const int s = i % stages;
    if (i >= stages) {  // wait for the copy that occupies the slot
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(phase) : "memory");
      if (s == stages - 1) phase ^= 1;
    }
    if (i < iters) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(STAGE_BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + (size_t)s * STAGE_BYTES)), "l"(base + (chunk % n_chunks) * STAGE_BYTES), "r"(STAGE_BYTES),
                     "r"(smem_u32(&bars[s])) : "memory");
      ++chunk;
    }
  }
  out[blockIdx.x] = clock64() - t0;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const size_t total = (size_t)sms * (64 << 20);  // 64 MB per CTA: 9.5 GB, far beyond the 126 MB L2
  uint8_t* d_src;
  long long* d_out;
  if (cudaMalloc(&d_src, total) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMemset(d_src, 1, total);
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  cudaFuncSetAttribute(k_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * STAGE_BYTES + 1024);
  const int iters = 4000;
  struct Case { const char* name; long long region; int share; } cases[] = {
      {"every CTA its own 64 MB (HBM stream)", 64ll << 20, 1},
      {"every CTA its own 512 KB (75 MB in all: L2-resident)", 512ll << 10, 1},
      {"pairs share 1 MB", 1ll << 20, 2},
      {"quads share 1 MB", 1ll << 20, 4},
      {"all CTAs the same 2 MB", 2ll << 20, sms},
  };
  for (int stages : {3, 4, 6}) {
    for (const Case& c : cases) {
      for (int rep = 0; rep < 2; ++rep) {
        k_fill<<<sms, 128, stages * STAGE_BYTES + 1024>>>(d_src, c.region, c.share, iters, stages, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      }
      long long h[1024], mx = 0;
      cudaMemcpy(h, d_out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bpc = (double)sms * iters * STAGE_BYTES / (double)mx;
      printf("stages=%d  %-56s %8lld clk  %7.0f B/clk chip-wide  %5.1f B/clk/SM\n", stages, c.name, mx, bpc, bpc / sms);
    }
  }
  printf("(%d SMs; 32 KB per copy; the tcgen05 pair GEMMs need 64 B/clk/SM = %d B/clk chip-wide at the full MMA rate)\n", sms, 64 * sms);
  return 0;
}
