// C-ABI of the mst library: engine lifetime, weight packing, and the launch
// sequence of one denoiser forward (MDM.forward / StyleDiffusion.forward,
// reference model/mdm_forstyledataset.py:315-364, :602-625) in either
// precision.  All functions are asynchronous on the caller's stream and are
// CUDA-graph capturable.
#include "common.cuh"
#include "simt.cuh"
#include "tc.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

namespace mst {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// ---- launch accounting / per-launch timing -----------------------------------
static std::atomic<uint64_t> g_launches{0};
struct Profile {
  bool open = false;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> ev;  // ev[0] = begin, ev[i+1] = after launch i
  std::vector<const char*> names;
};
// One profile per process (not per thread): torch runs backward passes on its autograd thread, and their launches
// belong in the table too.  Profiling is a single-threaded diagnostic; the mutex only keeps the vectors consistent.
static Profile g_prof;
static std::mutex g_prof_mu;

// During stream capture an ordinary record is swallowed into the graph's internal dependencies; an EXTERNAL
// record becomes an event-record node that updates the real event on every replay.
static cudaError_t record_event(cudaEvent_t e, cudaStream_t s) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
    return cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
  return cudaEventRecord(e, s);
}

void note_launch(const char* name, cudaStream_t s) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof.open && s == g_prof.stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) {
      record_event(e, s);
      g_prof.ev.push_back(e);
      g_prof.names.push_back(name);
    }
  }
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MST_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

int sm_count() {
  static int cached[64] = {};  // per device: a process may drive several GPUs (the reference selects one with --device N)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& c = cached[dev & 63];
  if (c == 0) {
    int n = 0;
    c = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return c;
}

// ---- workspace carving -----------------------------------------------------
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 1024);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return r;
  }
};

struct WsF32 {
  float *x, *y, *qkv, *ao, *h, *tmp;
};
struct WsBF16 {
  __nv_bfloat16 *x, *y, *qkv, *ao, *h, *xa;
};

static size_t carve_f32(const mst_model_desc& d, int n_seqs, int T, void* ws, WsF32* o) {
  const size_t M = (size_t)n_seqs * (T + 1);
  Carver c(ws);
  WsF32 w;
  w.x = c.take<float>(M * d.d_model);
  w.y = c.take<float>(M * d.d_model);
  w.qkv = c.take<float>(M * 3 * d.d_model);
  w.ao = c.take<float>(M * d.d_model);
  w.h = c.take<float>(M * d.d_ff);
  w.tmp = c.take<float>(M * d.d_model);
  if (o) *o = w;
  return align_up(c.off, 1024);
}

static size_t carve_bf16(const mst_model_desc& d, int f_pad, int n_seqs, int T, void* ws, WsBF16* o) {
  // +256 rows of slack: the attention kernel's TMA boxes may start inside the
  // last sequence and run past its end (OOB rows are zero-filled by TMA, the
  // slack only keeps plain pointer arithmetic in bounds).
  const size_t M = (size_t)n_seqs * (T + 1) + 256;
  Carver c(ws);
  WsBF16 w;
  w.x = c.take<__nv_bfloat16>(M * d.d_model);
  w.y = c.take<__nv_bfloat16>(M * d.d_model);
  w.qkv = c.take<__nv_bfloat16>(M * 3 * d.d_model);
  w.ao = c.take<__nv_bfloat16>(M * d.d_model);
  w.h = c.take<__nv_bfloat16>(M * d.d_ff);
  w.xa = c.take<__nv_bfloat16>((size_t)n_seqs * T * f_pad + 256 * f_pad);
  if (o) *o = w;
  return align_up(c.off, 1024);
}

static size_t packed_bytes(const mst_model_desc& d, int f_pad) {
  if (d.precision == MST_PREC_FP32) return 1024;
  Carver c(nullptr);
  c.take<__nv_bfloat16>((size_t)d.d_model * f_pad);   // in_w
  c.take<__nv_bfloat16>((size_t)f_pad * d.d_model);   // out_w
  c.take<float>(f_pad);                               // out_b padded
  for (int l = 0; l < d.n_layers; ++l) {
    c.take<__nv_bfloat16>((size_t)3 * d.d_model * d.d_model);
    c.take<__nv_bfloat16>((size_t)d.d_model * d.d_model);
    c.take<__nv_bfloat16>((size_t)d.d_ff * d.d_model);
    c.take<__nv_bfloat16>((size_t)d.d_model * d.d_ff);
  }
  for (int l = 0; l < d.n_layers; ++l) {  // transposed copies for the training backward (dX = dY W)
    c.take<__nv_bfloat16>((size_t)3 * d.d_model * d.d_model);
    c.take<__nv_bfloat16>((size_t)d.d_model * d.d_model);
    c.take<__nv_bfloat16>((size_t)d.d_ff * d.d_model);
    c.take<__nv_bfloat16>((size_t)d.d_model * d.d_ff);
  }
  c.take<__half>((size_t)f_pad * d.d_model);  // fp16 residual stream: final projection, QKV and linear1 weights
  for (int l = 0; l < d.n_layers; ++l) {
    c.take<__half>((size_t)3 * d.d_model * d.d_model);
    c.take<__half>((size_t)d.d_ff * d.d_model);
  }
  c.take<__nv_bfloat16>((size_t)f_pad * d.d_model);  // in_w^T, out_w^T: backward of the in-/out-projection (training)
  c.take<__nv_bfloat16>((size_t)d.d_model * f_pad);
  c.take<float>(f_pad);                               // zeros
  return align_up(c.off, 1024);
}

__global__ void pad_copy_f32_kernel(const float* src, float* dst, int n, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = i < n ? src[i] : 0.0f;
}

}  // namespace mst

using namespace mst;

extern "C" const char* mst_version(void) { return "mst-b200 0.1 (sm_100a)"; }
extern "C" const char* mst_last_error(void) { return g_last_error.c_str(); }

extern "C" uint64_t mst_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int mst_profile_begin(void* stream) {
  MST_CHECK_ARG(!g_prof.open, "a profile is already open on this thread");
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.stream = (cudaStream_t)stream;
  g_prof.ev.clear();
  g_prof.names.clear();
  cudaEvent_t e;
  MST_CUDA_OK(cudaEventCreate(&e));
  MST_CUDA_OK(record_event(e, g_prof.stream));  // during stream capture: an external event-record node
  g_prof.ev.push_back(e);
  g_prof.open = true;
  return MST_OK;
}

static int profile_collect(float* ms, char* names, int32_t cap, size_t names_cap, int32_t* n_out) {
  const int n = (int)g_prof.names.size();
  cudaError_t err = g_prof.ev.empty() ? cudaSuccess : cudaEventSynchronize(g_prof.ev.back());
  size_t off = 0;
  if (names && names_cap) names[0] = 0;
  for (int i = 0; i < n && err == cudaSuccess; ++i) {
    float t = 0.0f;
    err = cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]);
    if (ms && i < cap) ms[i] = t;
    if (names && i < cap) {
      size_t len = strlen(g_prof.names[i]);
      if (off + len + 2 <= names_cap) {
        memcpy(names + off, g_prof.names[i], len);
        names[off + len] = '\n';
        names[off + len + 1] = 0;
        off += len + 1;
      }
    }
  }
  *n_out = n;
  if (err != cudaSuccess) return fail(MST_ERR_CUDA, std::string("mst_profile: ") + cudaGetErrorString(err));
  return MST_OK;
}

extern "C" int mst_profile_end(float* ms, char* names, int32_t cap, size_t names_cap, int32_t* n_out) {
  MST_CHECK_ARG(g_prof.open, "no profile is open on this thread");
  g_prof.open = false;
  if (!ms && !names) {  // deferred: the launches were captured into a CUDA graph; read after replaying it
    if (n_out) *n_out = (int)g_prof.names.size();
    return MST_OK;
  }
  MST_CHECK_ARG(n_out != nullptr, "null n_out");
  return profile_collect(ms, names, cap, names_cap, n_out);
}

extern "C" int mst_profile_collect(float* ms, char* names, int32_t cap, size_t names_cap, int32_t* n_out) {
  MST_CHECK_ARG(!g_prof.open, "close the profile with mst_profile_end first");
  MST_CHECK_ARG(n_out != nullptr, "null n_out");
  return profile_collect(ms, names, cap, names_cap, n_out);
}

extern "C" int mst_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MST_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm) *sm = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return MST_OK;
}

extern "C" int mst_abi_sizes(size_t* model_desc, size_t* weights, size_t* forward_args, size_t* update_args) {
  if (model_desc) *model_desc = sizeof(mst_model_desc);
  if (weights) *weights = sizeof(mst_weights);
  if (forward_args) *forward_args = sizeof(mst_forward_args);
  if (update_args) *update_args = sizeof(mst_update_args);
  return MST_OK;
}

extern "C" int mst_engine_create(const mst_model_desc* desc, mst_engine_t* out) {
  MST_CHECK_ARG(desc && out, "null argument");
  const mst_model_desc& d = *desc;
  MST_CHECK_ARG(d.n_feats > 0 && d.d_model > 0 && d.n_heads > 0 && d.d_ff > 0, "non-positive dimension");
  MST_CHECK_ARG(d.n_layers > 0 && d.n_layers <= MST_MAX_LAYERS, "n_layers out of range");
  MST_CHECK_ARG(d.d_model % d.n_heads == 0, "d_model must be divisible by n_heads");
  MST_CHECK_ARG(d.d_model % 32 == 0 && d.d_model <= 1024, "d_model must be a multiple of 32 and <= 1024");
  MST_CHECK_ARG(d.precision == MST_PREC_FP32 || d.precision == MST_PREC_BF16, "unknown precision");
  if (d.precision == MST_PREC_BF16) {
    if (d.d_model != 512 || d.d_model / d.n_heads != 128 || d.d_ff % 256 != 0)
      return fail(MST_ERR_UNSUPPORTED,
                  "mst_engine_create: the bf16 tcgen05 path is built for d_model=512, head_dim=128, d_ff%256==0");
  }
  Engine* e = new Engine();
  e->desc = d;
  e->f_pad = (d.n_feats + 63) / 64 * 64;
  *out = reinterpret_cast<mst_engine_t>(e);
  return MST_OK;
}

extern "C" int mst_engine_destroy(mst_engine_t h) {
  delete reinterpret_cast<Engine*>(h);
  return MST_OK;
}

extern "C" int mst_engine_packed_weight_bytes(mst_engine_t h, size_t* bytes) {
  MST_CHECK_ARG(h && bytes, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  *bytes = packed_bytes(e->desc, e->f_pad);
  return MST_OK;
}

extern "C" int mst_engine_load_weights(mst_engine_t h, const mst_weights* w, void* packed_dev, size_t bytes,
                                       void* stream) {
  MST_CHECK_ARG(h && w, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  const mst_model_desc& d = e->desc;
  MST_CHECK_ARG(w->in_w && w->in_b && w->pe && w->t_w1 && w->t_b1 && w->t_w2 && w->t_b2 && w->out_w && w->out_b,
                "missing weight pointer");
  for (int l = 0; l < d.n_layers; ++l) {
    const mst_layer_weights& L = w->layers[l];
    MST_CHECK_ARG(L.qkv_w && L.qkv_b && L.o_w && L.o_b && L.w1 && L.b1 && L.w2 && L.b2 && L.ln1_g && L.ln1_b &&
                      L.ln2_g && L.ln2_b,
                  "missing layer weight pointer");
  }
  e->in_w = w->in_w; e->in_b = w->in_b; e->pe = w->pe;
  e->t_w1 = w->t_w1; e->t_b1 = w->t_b1; e->t_w2 = w->t_w2; e->t_b2 = w->t_b2;
  e->txt_w = w->txt_w; e->txt_b = w->txt_b; e->out_w = w->out_w; e->out_b = w->out_b;
  for (int l = 0; l < d.n_layers; ++l) {
    const mst_layer_weights& L = w->layers[l];
    e->lf[l] = LayerF32{L.qkv_w, L.qkv_b, L.o_w, L.o_b, L.w1, L.b1, L.w2, L.b2, L.ln1_g, L.ln1_b, L.ln2_g, L.ln2_b};
  }
  if (d.precision == MST_PREC_BF16) {
    MST_CHECK_ARG(packed_dev != nullptr, "packed_dev is required in bf16 mode");
    MST_CHECK_ARG(bytes >= packed_bytes(d, e->f_pad), "packed buffer too small");
    cudaStream_t s = (cudaStream_t)stream;
    Carver c(packed_dev);
    int rc;
    // every re-pack below is one job of ONE launch (this runs after every optimizer step of the finetune loop)
    CvtJobs js;
    const int dm = d.d_model, ff = d.d_ff;
    auto* in_w = c.take<__nv_bfloat16>((size_t)dm * e->f_pad);
    auto* out_w = c.take<__nv_bfloat16>((size_t)e->f_pad * dm);
    auto* out_b = c.take<float>(e->f_pad);
    pad_copy_f32_kernel<<<ceil_div(e->f_pad, 128), 128, 0, s>>>(w->out_b, out_b, d.n_feats, e->f_pad);
    MST_LAUNCHED("pad_copy", s);
    e->in_w_bf = in_w; e->out_w_bf = out_w; e->out_b_pad = out_b;
    for (int l = 0; l < d.n_layers; ++l) {
      auto* qkv = c.take<__nv_bfloat16>((size_t)3 * dm * dm);
      auto* o = c.take<__nv_bfloat16>((size_t)dm * dm);
      auto* w1 = c.take<__nv_bfloat16>((size_t)ff * dm);
      auto* w2 = c.take<__nv_bfloat16>((size_t)dm * ff);
      e->lb[l] = LayerBF16{qkv, o, w1, w2};
    }
    for (int l = 0; l < d.n_layers; ++l) {
      auto* qkv = c.take<__nv_bfloat16>((size_t)3 * dm * dm);  // [d, 3d]
      auto* o = c.take<__nv_bfloat16>((size_t)dm * dm);         // [d, d]
      auto* w1 = c.take<__nv_bfloat16>((size_t)ff * dm);        // [d, ff]
      auto* w2 = c.take<__nv_bfloat16>((size_t)dm * ff);        // [ff, d]
      e->lbt[l] = LayerBF16T{qkv, o, w1, w2};
    }
    auto* out_h = c.take<__half>((size_t)e->f_pad * dm);
    e->out_w_h = out_h;
    for (int l = 0; l < d.n_layers; ++l) {
      auto* qkv = c.take<__half>((size_t)3 * dm * dm);
      auto* w1 = c.take<__half>((size_t)ff * dm);
      e->lh[l] = LayerF16{qkv, w1};
    }
    auto bf = [](const __nv_bfloat16* p) { return const_cast<__nv_bfloat16*>(p); };
    auto hf = [](const __half* p) { return const_cast<__half*>(p); };
    cvt_jobs_add(js, w->in_w, dm, d.n_feats, d.n_feats, in_w, nullptr, dm, e->f_pad, nullptr, 0, nullptr);
    cvt_jobs_add(js, w->out_w, d.n_feats, dm, dm, out_w, out_h, e->f_pad, dm, nullptr, 0, nullptr);
    for (int l = 0; l < d.n_layers; ++l) {
      const mst_layer_weights& L = w->layers[l];
      if (js.n + 4 > CVT_MAX_JOBS) {
        if ((rc = cvt_multi(js, s, "pack_weights"))) return rc;
        js = CvtJobs();
      }
      cvt_jobs_add(js, L.qkv_w, 3 * dm, dm, dm, bf(e->lb[l].qkv_w), hf(e->lh[l].qkv_w), 3 * dm, dm, bf(e->lbt[l].qkv_w), 3 * dm, nullptr);
      cvt_jobs_add(js, L.o_w, dm, dm, dm, bf(e->lb[l].o_w), nullptr, dm, dm, bf(e->lbt[l].o_w), dm, nullptr);
      cvt_jobs_add(js, L.w1, ff, dm, dm, bf(e->lb[l].w1), hf(e->lh[l].w1), ff, dm, bf(e->lbt[l].w1), ff, nullptr);
      cvt_jobs_add(js, L.w2, dm, ff, ff, bf(e->lb[l].w2), nullptr, dm, ff, bf(e->lbt[l].w2), dm, nullptr);
    }
    {
      // transposed projections for the training backward; rows [n_feats, f_pad) of in_w^T and the zero vector are cleared
      auto* in_wt = c.take<__nv_bfloat16>((size_t)e->f_pad * dm);
      auto* out_wt = c.take<__nv_bfloat16>((size_t)dm * e->f_pad);
      auto* zeros = c.take<float>(e->f_pad);
      MST_CUDA_OK(cudaMemsetAsync(in_wt, 0, (size_t)e->f_pad * dm * sizeof(__nv_bfloat16), s));
      MST_CUDA_OK(cudaMemsetAsync(zeros, 0, (size_t)e->f_pad * sizeof(float), s));
      if (js.n + 2 > CVT_MAX_JOBS) {
        if ((rc = cvt_multi(js, s, "pack_weights"))) return rc;
        js = CvtJobs();
      }
      cvt_jobs_add(js, w->in_w, dm, d.n_feats, d.n_feats, nullptr, nullptr, 0, 0, in_wt, dm, nullptr);     // [F, d]
      cvt_jobs_add(js, w->out_w, d.n_feats, dm, dm, nullptr, nullptr, 0, 0, out_wt, e->f_pad, nullptr);    // [d, Fpad]
      e->in_wt_bf = in_wt; e->out_wt_bf = out_wt; e->zero_pad = zeros;
    }
    if ((rc = cvt_multi(js, s, "pack_weights"))) return rc;
    {
      const char* env = getenv("MST_STREAM_F16");
      e->stream_f16 = !(env && env[0] == '0');
    }
  }
  e->weights_loaded = true;
  return MST_OK;
}

extern "C" int mst_engine_workspace_bytes(mst_engine_t h, int n_seqs, int n_frames, size_t* bytes) {
  MST_CHECK_ARG(h && bytes, "null argument");
  MST_CHECK_ARG(n_seqs > 0 && n_frames > 0, "non-positive size");
  Engine* e = reinterpret_cast<Engine*>(h);
  *bytes = e->desc.precision == MST_PREC_FP32 ? carve_f32(e->desc, n_seqs, n_frames, nullptr, nullptr)
                                               : carve_bf16(e->desc, e->f_pad, n_seqs, n_frames, nullptr, nullptr);
  return MST_OK;
}

extern "C" int mst_time_embed(mst_engine_t h, const int64_t* t_dev, int n, float* out_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream) {
  MST_CHECK_ARG(h && t_dev && out_dev && workspace_dev, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  MST_CHECK_ARG(e->weights_loaded, "weights not loaded");
  MST_CHECK_ARG(n > 0, "n must be positive");
  const int d = e->desc.d_model;
  MST_CHECK_ARG(workspace_bytes >= (size_t)n * d * sizeof(float), "workspace too small (need n*d floats)");
  cudaStream_t s = (cudaStream_t)stream;
  float* hid = static_cast<float*>(workspace_dev);
  GemmF32Params p;
  p.a = e->pe; p.lda = d; p.a_mode = A_GATHER_ROWS; p.gather = t_dev;
  p.w = e->t_w1; p.ldw = d; p.bias = e->t_b1; p.c = hid; p.ldc = d;
  p.M = n; p.N = d; p.K = d; p.epi = EPI_SILU; p.row_invariant = 1;
  int rc = gemm_f32(p, s);
  if (rc) return rc;
  GemmF32Params q;
  q.a = hid; q.lda = d; q.w = e->t_w2; q.ldw = d; q.bias = e->t_b2; q.c = out_dev; q.ldc = d;
  q.M = n; q.N = d; q.K = d; q.epi = EPI_PLAIN; q.row_invariant = 1;
  return gemm_f32(q, s);
}

extern "C" int mst_text_embed(mst_engine_t h, const float* feat_dev, int n, float* out_dev, void* stream) {
  MST_CHECK_ARG(h && feat_dev && out_dev, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  MST_CHECK_ARG(e->weights_loaded, "weights not loaded");
  MST_CHECK_ARG(e->txt_w && e->txt_b, "model has no embed_text weights");
  MST_CHECK_ARG(n > 0, "n must be positive");
  const int d = e->desc.d_model;
  GemmF32Params p;
  p.a = feat_dev; p.lda = e->desc.clip_dim; p.w = e->txt_w; p.ldw = e->desc.clip_dim; p.bias = e->txt_b;
  p.c = out_dev; p.ldc = d; p.M = n; p.N = d; p.K = e->desc.clip_dim; p.epi = EPI_PLAIN; p.row_invariant = 1;
  return gemm_f32(p, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
static int forward_f32(Engine* e, const mst_forward_args& a, cudaStream_t s) {
  const mst_model_desc& d = e->desc;
  const int B = a.batch, T = a.n_frames, S = T + 1, n_pass = a.cfg ? 2 : 1, NS = B * n_pass;
  const int M = NS * S, dm = d.d_model;
  WsF32 w;
  size_t need = carve_f32(d, NS, T, a.workspace, &w);
  MST_CHECK_ARG(a.workspace_bytes >= need, "workspace too small");
  int rc;
  Token0Params t0;
  t0.temb = a.temb; t0.temb_row_dev = a.temb_row_dev; t0.temb_row_offset = a.temb_row_offset;
  t0.text_emb = a.text_emb; t0.txt_b = e->txt_b; t0.pe = e->pe; t0.x_f32 = w.x;
  t0.B = B; t0.T = T; t0.d = dm; t0.cfg = a.cfg; t0.uncond = a.uncond;
  if ((rc = token0(t0, NS, s))) return rc;
  {
    GemmF32Params p;
    p.a = a.x; p.a_mode = A_MOTION; p.w = e->in_w; p.ldw = d.n_feats; p.bias = e->in_b;
    p.c = w.x; p.ldc = dm; p.M = B * T; p.N = dm; p.K = d.n_feats; p.epi = EPI_INPROJ;
    p.pe = e->pe; p.B = B; p.T = T; p.n_pass = n_pass;
    if ((rc = gemm_f32(p, s))) return rc;
  }
  for (int l = 0; l < d.n_layers; ++l) {
    const LayerF32& L = e->lf[l];
    GemmF32Params p;
    p.a = w.x; p.lda = dm; p.w = L.qkv_w; p.ldw = dm; p.bias = L.qkv_b; p.c = w.qkv; p.ldc = 3 * dm;
    p.M = M; p.N = 3 * dm; p.K = dm; p.epi = EPI_PLAIN;
    if ((rc = gemm_f32(p, s))) return rc;
    if ((rc = attention_f32(w.qkv, w.ao, NS, S, dm, d.n_heads, s))) return rc;
    GemmF32Params o;
    o.a = w.ao; o.lda = dm; o.w = L.o_w; o.ldw = dm; o.bias = L.o_b; o.c = w.tmp; o.ldc = dm;
    o.M = M; o.N = dm; o.K = dm; o.epi = EPI_RESIDUAL; o.residual = w.x;
    if ((rc = gemm_f32(o, s))) return rc;
    if ((rc = layernorm_f32(w.tmp, L.ln1_g, L.ln1_b, w.y, M, dm, s))) return rc;
    GemmF32Params f1;
    f1.a = w.y; f1.lda = dm; f1.w = L.w1; f1.ldw = dm; f1.bias = L.b1; f1.c = w.h; f1.ldc = d.d_ff;
    f1.M = M; f1.N = d.d_ff; f1.K = dm; f1.epi = EPI_GELU;
    if ((rc = gemm_f32(f1, s))) return rc;
    GemmF32Params f2;
    f2.a = w.h; f2.lda = d.d_ff; f2.w = L.w2; f2.ldw = d.d_ff; f2.bias = L.b2; f2.c = w.tmp; f2.ldc = dm;
    f2.M = M; f2.N = dm; f2.K = d.d_ff; f2.epi = EPI_RESIDUAL; f2.residual = w.y;
    if ((rc = gemm_f32(f2, s))) return rc;
    if ((rc = layernorm_f32(w.tmp, L.ln2_g, L.ln2_b, w.x, M, dm, s))) return rc;
  }
  for (int pass = 0; pass < n_pass; ++pass) {
    GemmF32Params p;
    p.a = w.x + (size_t)pass * B * S * dm; p.lda = dm; p.w = e->out_w; p.ldw = dm; p.bias = e->out_b;
    p.c = pass == 0 ? a.out_cond : a.out_uncond; p.M = B * S; p.N = d.n_feats; p.K = dm;
    p.epi = EPI_OUTPROJ; p.T = T; p.B = B;
    if ((rc = gemm_f32(p, s))) return rc;
  }
  return MST_OK;
}

static int forward_bf16(Engine* e, const mst_forward_args& a, cudaStream_t s) {
  const mst_model_desc& d = e->desc;
  const int B = a.batch, T = a.n_frames, S = T + 1, n_pass = a.cfg ? 2 : 1, NS = B * n_pass;
  const int M = NS * S, dm = d.d_model;
  const int f16 = e->stream_f16 ? 1 : 0;  // fp16 residual stream (x, y)
  WsBF16 w;
  size_t need = carve_bf16(d, e->f_pad, NS, T, a.workspace, &w);
  MST_CHECK_ARG(a.workspace_bytes >= need, "workspace too small");
  int rc;
  Token0Params t0;
  t0.temb = a.temb; t0.temb_row_dev = a.temb_row_dev; t0.temb_row_offset = a.temb_row_offset;
  t0.text_emb = a.text_emb; t0.txt_b = e->txt_b; t0.pe = e->pe;
  if (f16) t0.x_f16 = reinterpret_cast<__half*>(w.x); else t0.x_bf16 = w.x;
  t0.B = B; t0.T = T; t0.d = dm; t0.cfg = a.cfg; t0.uncond = a.uncond;
  if ((rc = motion_to_tokens_bf16(a.x, w.xa, B, d.n_feats, T, e->f_pad, &t0, NS, s))) return rc;
  {
    TcGemmParams p;
    p.a = w.xa; p.w = e->in_w_bf; p.bias = e->in_b; p.out = w.x; p.ldo = dm;
    p.M = B * T; p.N = dm; p.K = e->f_pad; p.epi = TC_EPI_INPROJ; p.pe = e->pe; p.B = B; p.T = T; p.n_pass = n_pass;
    p.io_f16 = f16;
    if ((rc = tc_gemm(p, s))) return rc;
  }
  for (int l = 0; l < d.n_layers; ++l) {
    const LayerF32& L = e->lf[l];
    const LayerBF16& Lb = e->lb[l];
    TcGemmParams p;
    p.a = w.x; p.w = f16 ? reinterpret_cast<const __nv_bfloat16*>(e->lh[l].qkv_w) : Lb.qkv_w; p.bias = L.qkv_b;
    p.out = w.qkv; p.ldo = 3 * dm; p.ab_f16 = f16;
    p.M = M; p.N = 3 * dm; p.K = dm; p.epi = TC_EPI_BIAS_BF16;
    if ((rc = tc_gemm(p, s))) return rc;
    TcAttnParams at;
    at.qkv = w.qkv; at.out = w.ao; at.n_seqs = NS; at.S = S; at.d_model = dm; at.n_heads = d.n_heads;
    if ((rc = tc_attention(at, s))) return rc;
    TcGemmParams o;
    o.a = w.ao; o.w = Lb.o_w; o.bias = L.o_b; o.out = w.y; o.ldo = dm; o.M = M; o.N = dm; o.K = dm;
    o.epi = TC_EPI_BIAS_RES_LN; o.residual = w.x; o.ln_g = L.ln1_g; o.ln_b = L.ln1_b; o.io_f16 = f16;
    if ((rc = tc_gemm(o, s))) return rc;
    TcGemmParams f1;
    f1.a = w.y; f1.w = f16 ? reinterpret_cast<const __nv_bfloat16*>(e->lh[l].w1) : Lb.w1; f1.bias = L.b1; f1.out = w.h;
    f1.ldo = d.d_ff; f1.M = M; f1.N = d.d_ff; f1.K = dm; f1.ab_f16 = f16;
    f1.epi = TC_EPI_BIAS_GELU_BF16;
    if ((rc = tc_gemm(f1, s))) return rc;
    TcGemmParams f2;
    f2.a = w.h; f2.w = Lb.w2; f2.bias = L.b2; f2.out = w.x; f2.ldo = dm; f2.M = M; f2.N = dm; f2.K = d.d_ff;
    f2.epi = TC_EPI_BIAS_RES_LN; f2.residual = w.y; f2.ln_g = L.ln2_g; f2.ln_b = L.ln2_b; f2.io_f16 = f16;
    if ((rc = tc_gemm(f2, s))) return rc;
  }
  {
    TcGemmParams p;
    p.a = w.x; p.w = f16 ? reinterpret_cast<const __nv_bfloat16*>(e->out_w_h) : e->out_w_bf; p.ab_f16 = f16;
    p.bias = e->out_b_pad; p.out = a.out_cond; p.out2 = a.out_uncond;
    p.M = M; p.N = e->f_pad; p.K = dm; p.epi = TC_EPI_OUTPROJ_F32; p.B = B; p.T = T; p.n_valid = d.n_feats;
    if ((rc = tc_gemm(p, s))) return rc;
  }
  return MST_OK;
}

extern "C" int mst_denoiser_forward(mst_engine_t h, const mst_forward_args* ap, void* stream) {
  MST_CHECK_ARG(h && ap, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  const mst_forward_args& a = *ap;
  MST_CHECK_ARG(e->weights_loaded, "weights not loaded");
  MST_CHECK_ARG(a.batch > 0 && a.n_frames > 0, "empty batch");
  MST_CHECK_ARG(a.n_frames + 1 <= e->desc.pe_len, "sequence longer than the positional table");
  MST_CHECK_ARG(a.x && a.temb && a.out_cond && a.workspace, "null tensor pointer");
  MST_CHECK_ARG(!a.cfg || a.out_uncond, "cfg needs out_uncond");
  MST_CHECK_ARG(!a.cfg || (a.text_emb && e->txt_b), "cfg needs text_emb and a text-conditioned model");
  cudaStream_t s = (cudaStream_t)stream;
  return e->desc.precision == MST_PREC_FP32 ? forward_f32(e, a, s) : forward_bf16(e, a, s);
}

extern "C" int mst_test_gemm_bf16(const void* a_bf16, const void* w_bf16, const float* bias, float* c, int32_t m,
                                  int32_t n, int32_t k, void* stream) {
  MST_CHECK_ARG(a_bf16 && w_bf16 && c, "null pointer");
  TcGemmParams p;
  p.a = static_cast<const __nv_bfloat16*>(a_bf16);
  p.w = static_cast<const __nv_bfloat16*>(w_bf16);
  p.bias = bias; p.out = c; p.ldo = n; p.M = m; p.N = n; p.K = k; p.epi = TC_EPI_BIAS_F32;
  return tc_gemm(p, (cudaStream_t)stream);
}

extern "C" int mst_test_gemm_epi_bf16(int32_t epi, const void* a_bf16, const void* w_bf16, const float* bias,
                                      const void* residual_bf16, const float* ln_g, const float* ln_b, void* out_bf16,
                                      int32_t m, int32_t n, int32_t k, void* stream) {
  MST_CHECK_ARG(a_bf16 && w_bf16 && bias && out_bf16, "null pointer");
  MST_CHECK_ARG(epi >= 0 && epi <= 2, "epi must be 0 (bias), 1 (gelu) or 2 (residual + layernorm)");
  TcGemmParams p;
  p.a = static_cast<const __nv_bfloat16*>(a_bf16);
  p.w = static_cast<const __nv_bfloat16*>(w_bf16);
  p.bias = bias; p.out = out_bf16; p.ldo = n; p.M = m; p.N = n; p.K = k;
  p.epi = epi == 0 ? TC_EPI_BIAS_BF16 : (epi == 1 ? TC_EPI_BIAS_GELU_BF16 : TC_EPI_BIAS_RES_LN);
  p.residual = static_cast<const __nv_bfloat16*>(residual_bf16);
  p.ln_g = ln_g; p.ln_b = ln_b;
  return tc_gemm(p, (cudaStream_t)stream);
}

extern "C" int mst_test_set_gemm_debug(void* dev_buf_int64) {
  set_gemm_debug(static_cast<long long*>(dev_buf_int64));
  return MST_OK;
}

extern "C" int mst_test_attention_bf16(mst_engine_t h, const void* qkv_bf16, void* out_bf16, int32_t n_seqs,
                                       int32_t seq_len, void* workspace, size_t workspace_bytes, void* stream) {
  MST_CHECK_ARG(h && qkv_bf16 && out_bf16, "null pointer");
  Engine* e = reinterpret_cast<Engine*>(h);
  (void)workspace; (void)workspace_bytes;
  TcAttnParams at;
  at.qkv = static_cast<const __nv_bfloat16*>(qkv_bf16);
  at.out = static_cast<__nv_bfloat16*>(out_bf16);
  at.n_seqs = n_seqs; at.S = seq_len; at.d_model = e->desc.d_model; at.n_heads = e->desc.n_heads;
  return tc_attention(at, (cudaStream_t)stream);
}
