// Shared helpers for the mst (motion-style-transfer sampler) sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mst.h"

namespace mst {

// thread-local last error, exposed through mst_last_error()
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define MST_CHECK_ARG(cond, msg)                                              \
  do {                                                                        \
    if (!(cond)) return ::mst::fail(MST_ERR_INVALID, std::string(__func__) + ": " + (msg)); \
  } while (0)

#define MST_CUDA_OK(expr)                                                     \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess)                                                    \
      return ::mst::fail(MST_ERR_CUDA, std::string(__func__) + ": " #expr ": " + cudaGetErrorString(_e)); \
  } while (0)

// after every kernel launch: surface launch errors, count the launch, and (when a
// profile is open on this thread, mst_profile_begin) record a CUDA event after it.
void note_launch(const char* name, cudaStream_t s);

#define MST_LAUNCHED(name, stream)                                            \
  do {                                                                        \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess)                                                    \
      return ::mst::fail(MST_ERR_CUDA, std::string(__func__) + ": launch of " + (name) + ": " + cudaGetErrorString(_e)); \
    ::mst::note_launch((name), (stream));                                     \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every kernel of the denoise step is launched with programmaticStreamSerialization: its CTAs may be scheduled
// (and run their prologue: barrier init, TMEM allocation, tensor-map prefetch) as soon as SMs drain from the
// previous kernel; pdl_wait() then blocks until that kernel has completed and flushed its memory.  Without it
// ~7 us of launch latency sat between consecutive kernels of the captured step (profiles/r01c_*).
bool pdl_enabled();  // MST_PDL=0 switches it off

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();  // cached multiProcessorCount of the current device

// first() is true once per CUDA device: per-function attributes (cudaFuncSetAttribute) belong to a device, so a process
// that touches a second GPU must set them again there
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// ---------------------------------------------------------------------------
// engine (defined in api.cu); kernels get what they need through these structs
// ---------------------------------------------------------------------------
struct LayerF32 {
  const float *qkv_w, *qkv_b, *o_w, *o_b, *w1, *b1, *w2, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct LayerBF16 {
  const __nv_bfloat16 *qkv_w, *o_w, *w1, *w2;  // [N, K] row-major bf16
};
struct LayerF16 {
  const __half *qkv_w, *w1;  // fp16 copies of the two weights that multiply the fp16 residual stream (sampler only)
};
struct LayerBF16T {
  const __nv_bfloat16 *qkv_w, *o_w, *w1, *w2;  // the same weights transposed, [K, N] row-major (dX = dY W)
};

struct Engine {
  mst_model_desc desc;
  bool weights_loaded = false;
  // fp32 views (always valid once loaded; biases / LN params stay fp32 in both modes)
  const float *in_w, *in_b, *pe, *t_w1, *t_b1, *t_w2, *t_b2, *txt_w, *txt_b, *out_w, *out_b;
  LayerF32 lf[MST_MAX_LAYERS];
  // bf16 packed copies (MST_PREC_BF16)
  const __nv_bfloat16* in_w_bf = nullptr;   // [d, Fpad]
  const __nv_bfloat16* out_w_bf = nullptr;  // [Fpad, d]
  const __nv_bfloat16* in_wt_bf = nullptr;   // [Fpad, d]  in_w^T  (training backward of the in-projection on tensor cores)
  const __nv_bfloat16* out_wt_bf = nullptr;  // [d, Fpad]  out_w^T (... of the out-projection)
  const float* zero_pad = nullptr;           // [Fpad] zeros (bias-free out-projection-style epilogue)
  const float* out_b_pad = nullptr;         // [Fpad]
  LayerBF16 lb[MST_MAX_LAYERS];
  LayerBF16T lbt[MST_MAX_LAYERS];
  // fp16 residual stream of the sampler (MST_STREAM_F16, default on): LayerNorm outputs / in-projection output are
  // written in IEEE fp16, and the GEMMs that read them (QKV, FFN linear1, final projection) use fp16 weights
  bool stream_f16 = false;
  LayerF16 lh[MST_MAX_LAYERS];
  const __half* out_w_h = nullptr;  // [Fpad, d]
  int f_pad = 0;  // F rounded up to 64
};

}  // namespace mst
