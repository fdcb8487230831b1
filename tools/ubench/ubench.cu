// Micro-benchmarks that calibrate the epilogue / softmax cost models of the tcgen05 kernels (B200, sm_100a):
// issue rates of the instructions those loops are made of, measured with clock64 on ONE CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench tools/ubench/ubench.cu && tools/ubench/ubench
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 1/2: shared-memory loads, all lanes the same address (parameter broadcast) or lane-distinct rows ----
template <int MODE>  // 0: lds128 broadcast, 1: lds32 broadcast, 2: lds128 conflict-free distinct, 3: lds64 broadcast
__global__ void k_lds(long long* out, int nwarps_active) {
  __shared__ __align__(16) float buf[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps_active) return;
  uint32_t base = smem_u32(buf);
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {
      uint32_t a = base + ((i * 16) & 8191) * 1;
      float x, y, z, w;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
      acc0 += x; acc1 += y; acc2 += z; acc3 += w;
    } else if (MODE == 1) {
      uint32_t a = base + ((i * 4) & 8191);
      float x;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
      acc0 += x;
    } else if (MODE == 2) {
      uint32_t a = base + (((i & 7) * 512 + lane * 16) & 32767);
      float x, y, z, w;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
      acc0 += x; acc1 += y; acc2 += z; acc3 += w;
    } else {
      uint32_t a = base + ((i * 8) & 8191);
      float x, y;
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
      acc0 += x; acc1 += y;
    }
  }
  long long t1 = clock64();
  if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
  if (acc0 + acc1 + acc2 + acc3 == 12345.678f) out[100] = 1;
}

// ---- 3/4: arithmetic issue rates ----
template <int MODE>  // 0: fma.f32 scalar x2 (two per iteration), 1: fma.rn.f32x2, 2: cvt.rn.bf16x2.f32, 3: ex2.approx, 4: fmnmx
__global__ void k_alu(long long* out, int nwarps_active, float seed) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps_active) return;
  float a[8], b = seed + lane * 1e-3f, c = 0.5f;
  uint64_t p[8];
  uint32_t u[8];
  for (int j = 0; j < 8; ++j) { a[j] = seed * j + lane; p[j] = ((uint64_t)__float_as_uint(a[j]) << 32) | __float_as_uint(b); u[j] = j; }
  const uint64_t pb = ((uint64_t)__float_as_uint(b) << 32) | __float_as_uint(b);
  const uint64_t pc = ((uint64_t)__float_as_uint(c) << 32) | __float_as_uint(c);
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) a[j] = fmaf(a[j], b, c);
      else if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(pb), "l"(pc));
      else if (MODE == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[j]) : "f"(a[j]), "f"(__uint_as_float(u[j])));
      else if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      else a[j] = fmaxf(a[j], b + (float)i);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += a[j] + (float)(p[j] & 0xff) + (float)u[j];
  if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
  if (s == 12345.678f) out[100] = 1;
}

// ---- 5: TMEM -> register loads ----
__global__ void k_tmem(long long* out, int nwarps_active, int cols_per_ld) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = slot;
  if (warp < nwarps_active) {
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      uint32_t r[32];
      const uint32_t addr = tbase + lane_addr + (uint32_t)(((i * 32) + (warp >> 2) * 64) & 511 & ~31);
      if (cols_per_ld == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(addr)
            : "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[31];
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    if (acc == 0x12345678u) out[100] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

// ---- 6: st.global row-chunk stores: 16 B per lane, lanes 8 per 128-byte line (the LN v5 output pattern) vs lane-per-row
template <int MODE>  // 0: coalesced 4 rows x 128 B per instruction, 1: one 16-byte piece of 32 different rows
__global__ void k_stg(long long* out, uint4* dst, int nwarps_active) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps_active) return;
  uint4 v = make_uint4(lane, warp, 3, 4);
  char* base = reinterpret_cast<char*>(dst) + ((size_t)blockIdx.x * 16 + warp) * (size_t)(1 << 20);
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    char* p;
    if (MODE == 0) p = base + (size_t)(i * 4 + (lane >> 3)) * 1024 + (lane & 7) * 16;
    else p = base + (size_t)((i >> 3) * 32 + lane) * 1024 + (i & 7) * 16;
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  long long t1 = clock64();
  if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
}


// ---- 7: TMEM -> register loads WHILE the tensor core runs back-to-back MMAs into the same TMEM (attention: the softmax
// warps read S / O of one item while Q K^T / P V of the neighbouring item execute)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__global__ void __launch_bounds__(512) k_tmem_mma(long long* out, int nwarps_ld, int n_mma, int mma_n) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(sm) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // bf16 ~0.0078
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = slot;
  if (warp == 15) {  // MMA issuer: n_mma instructions M=128, N=mma_n, K=16 into columns [256, 256 + mma_n)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t ad = desc_sw128(sbase + (i & 3) * 32, 0, 1024), bd = desc_sw128(sbase + 32768 + (i & 3) * 32, 0, 1024);
        const uint32_t acc = 1;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                         tbase + 256),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar))
                     : "memory");
      long long t1 = clock64();
      if (blockIdx.x == 0) out[40] = t1 - t0;
    }
  } else if (warp < nwarps_ld) {
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      uint32_t r[32];
      const uint32_t addr = tbase + lane_addr + (uint32_t)(((i * 32) + (warp >> 2) * 64) & 255 & ~31);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(addr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[31];
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    if (acc == 0x12345678u) out[100] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

static void report(const char* name, long long* d_out, int nw, double ops_per_iter) {
  long long h[32];
  cudaDeviceSynchronize();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e)); return; }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per_warp_instr = (double)mx / (ITERS * ops_per_iter);
  printf("%-44s warps=%2d  %8lld clk  -> %.2f clk per warp-instruction per warp, %.2f clk per instruction SM-wide\n", name, nw,
         mx, per_warp_instr, per_warp_instr / nw);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  uint4* d_dst;
  cudaMalloc(&d_dst, (size_t)148 * 16 << 20);
  const int grid = 148;
  for (int nw : {4, 8, 16}) {
    k_lds<0><<<grid, 512>>>(d_out, nw); report("lds128 broadcast", d_out, nw, 1);
    k_lds<3><<<grid, 512>>>(d_out, nw); report("lds64 broadcast", d_out, nw, 1);
    k_lds<1><<<grid, 512>>>(d_out, nw); report("lds32 broadcast", d_out, nw, 1);
    k_lds<2><<<grid, 512>>>(d_out, nw); report("lds128 lane-distinct (conflict-free)", d_out, nw, 1);
    k_alu<0><<<grid, 512>>>(d_out, nw, 1.0001f); report("ffma scalar", d_out, nw, 8);
    k_alu<1><<<grid, 512>>>(d_out, nw, 1.0001f); report("fma.rn.f32x2", d_out, nw, 8);
    k_alu<2><<<grid, 512>>>(d_out, nw, 1.0001f); report("cvt.rn.bf16x2.f32", d_out, nw, 8);
    k_alu<3><<<grid, 512>>>(d_out, nw, 1.0001f); report("ex2.approx", d_out, nw, 8);
    k_alu<4><<<grid, 512>>>(d_out, nw, 1.0001f); report("fmnmx", d_out, nw, 8);
    k_tmem<<<grid, 512>>>(d_out, nw, 32); report("tcgen05.ld 32x32b.x32 (4 KB per warp-instr)", d_out, nw, 1);
    k_stg<0><<<grid, 512>>>(d_out, d_dst, nw); report("st.global.v4 4 rows x 128 B", d_out, nw, 1);
    k_stg<1><<<grid, 512>>>(d_out, d_dst, nw); report("st.global.v4 32 rows x 16 B", d_out, nw, 1);
  }
  cudaFuncSetAttribute(k_tmem_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int nw : {4, 8}) {
    for (int n_mma : {0, 2000}) {
      for (int mma_n : {208, 128}) {
        if (n_mma == 0 && mma_n == 128) continue;
        cudaMemset(d_out, 0, 1024 * sizeof(long long));
        k_tmem_mma<<<grid, 512, 100 * 1024>>>(d_out, nw, n_mma, mma_n);
        char name[128];
        snprintf(name, sizeof(name), "tcgen05.ld x32 with %d MMAs (N=%d) in flight", n_mma, mma_n);
        report(name, d_out, nw, 1);
        long long h[64];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        if (n_mma) printf("      the %d MMAs took %lld clk = %.1f clk each\n", n_mma, h[40], (double)h[40] / n_mma);
      }
    }
  }
  return 0;
}
