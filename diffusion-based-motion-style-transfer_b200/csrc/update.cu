// Fused per-step diffusion update (memory-bound, one pass over the state).
//
// Replaces, per element of the [B,F,1,T] motion state and in ONE kernel:
//   cfg lerp          reference model/cfg_sampler.py:43
//   inpainting blend  reference diffusion/gaussian_diffusion.py:341-349
//   x0 clamp          reference diffusion/gaussian_diffusion.py:390-396
//   posterior mean    reference diffusion/gaussian_diffusion.py:295-298
//   masked noise add  reference diffusion/inpainting_gaussian_diffusion.py:51-63
//   DDIM variant      reference diffusion/inpainting_gaussian_diffusion.py:155-174
// The reference issues ~25 elementwise torch kernels plus 5 host->device table
// copies per step for this; here every table is resident and the state is read
// and written exactly once (algorithmic bytes per element are listed in
// DESIGN.md).  Arithmetic is done with explicit round-to-nearest intrinsics in
// the reference's operation order so that no FMA contraction changes a bit.
#include "common.cuh"

namespace mst {

// ------------------------------- Philox4x32-10 -----------------------------
struct U4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    U4 n;
    n.x = hi1 ^ ctr.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ ctr.w ^ k1;
    n.w = lo0;
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}

// uniform in (0,1) with 24 bits, then Box-Muller.  Mirrored by oracle/philox.py.
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  // SFU forms (lg2 / sin / cos approximations, ~1e-6 absolute on the result): the accurate logf / sincospif sequences made
  // the update kernel issue-bound at B >= 2048 (73 % issue-active, 0.77 of the HBM copy peak; profiles/r01e_*)
  float u1 = u01(a), u2 = u01(b);
  float rad = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  n0 = rad * c;
  n1 = rad * s;
}

// 4 normals for elements [4*vec, 4*vec+3] of sample `sample` at step key `t`.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t sample, uint32_t vec, int t) {
  U4 ctr{vec, (uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)t};
  U4 r = philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  float4 o;
  box_muller(r.x, r.y, o.x, o.y);
  box_muller(r.z, r.w, o.z, o.w);
  return o;
}

// ------------------------------- the update --------------------------------
struct StepCoef {
  float c1, c2, sig, recip, recipm1;
};

template <int SAMPLER>
__device__ __forceinline__ void update_elem(float oc, float ou, bool has_u, float scale, float xt, float m,
                                            bool has_inp, float xin, bool clip, float eps, bool has_noise,
                                            bool mask_noise, const StepCoef& k, float& x_prev, float& x0o) {
  float x0 = oc;
  if (has_u) x0 = __fadd_rn(ou, __fmul_rn(scale, __fsub_rn(oc, ou)));
  if (has_inp) x0 = __fadd_rn(__fmul_rn(x0, __fsub_rn(1.0f, m)), __fmul_rn(xin, m));
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  float mean;
  if (SAMPLER == MST_SAMPLER_DDPM) {
    mean = __fadd_rn(__fmul_rn(k.c1, x0), __fmul_rn(k.c2, xt));
  } else {
    float e = __fdiv_rn(__fsub_rn(__fmul_rn(k.recip, xt), x0), k.recipm1);
    mean = __fadd_rn(__fmul_rn(x0, k.c1), __fmul_rn(k.c2, e));
  }
  if (has_noise) {
    if (mask_noise) eps = __fmul_rn(eps, __fsub_rn(1.0f, m));
    mean = __fadd_rn(mean, __fmul_rn(k.sig, eps));
  }
  x_prev = mean;
  x0o = x0;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

template <int SAMPLER, int NOISE, bool VEC4>
__global__ void __launch_bounds__(256) update_kernel(mst_update_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t per = (int64_t)a.n_feats * a.n_frames;  // elements per sample
  const int T = a.n_frames;
  const bool has_u = a.out_uncond != nullptr;
  const bool has_inp = a.mask_kind != MST_MASK_NONE && a.x_inpaint != nullptr;
  const bool has_mask = a.mask_kind != MST_MASK_NONE;
  const bool clip = a.clip_denoised != 0;
  const bool mask_noise = a.mask_noise != 0 && has_mask;
  int t_shared = a.t_imm;
  if (a.t_scalar_dev) t_shared = *a.t_scalar_dev;

  constexpr int W = VEC4 ? 4 : 1;
  const int64_t units_per = per / W;  // VEC4 requires per % 4 == 0 and T % 4 == 0
  const int64_t total = units_per * a.batch;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total;
       u += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(u / units_per);
    const int64_t r = (u - (int64_t)b * units_per) * W;  // element offset inside the sample
    const int64_t i = (int64_t)b * per + r;
    const int t = a.t_vec ? (int)a.t_vec[b] : t_shared;
    StepCoef k;
    k.c1 = a.coef1[t];
    k.c2 = a.coef2[t];
    const bool nz = (t != 0);
    const bool has_noise = NOISE != MST_NOISE_NONE;
    k.sig = has_noise ? __fmul_rn(nz ? 1.0f : 0.0f, a.sigma[t]) : 0.0f;
    if (SAMPLER == MST_SAMPLER_DDIM) {
      k.recip = a.recip[t];
      k.recipm1 = a.recipm1[t];
    }
    const float scale = (has_u && a.cfg_scale) ? a.cfg_scale[b] : 1.0f;
    const int f = (int)(r / T);
    if (VEC4) {
      float4 oc = ld4_stream(a.out_cond + i);
      float4 ou = has_u ? ld4_stream(a.out_uncond + i) : make_float4(0, 0, 0, 0);
      float4 xt = ld4_stream(a.x_t + i);
      float4 xin = has_inp ? ld4(a.x_inpaint + i) : make_float4(0, 0, 0, 0);
      float4 m = make_float4(0, 0, 0, 0);
      if (a.mask_kind == MST_MASK_FULL) m = ld4(a.mask + i);
      else if (a.mask_kind == MST_MASK_FT) m = ld4(a.mask + r);
      else if (a.mask_kind == MST_MASK_F) { float v = __ldg(a.mask + f); m = make_float4(v, v, v, v); }
      float4 eps = make_float4(0, 0, 0, 0);
      if (NOISE == MST_NOISE_TENSOR) {
        eps = ld4_stream(a.noise + (a.const_noise ? r : i));
      } else if (NOISE == MST_NOISE_PHILOX) {
        uint64_t sample = a.const_noise ? a.philox_sample_offset : a.philox_sample_offset + (uint64_t)b;
        eps = philox_normal4(a.philox_seed, sample, (uint32_t)(r >> 2), t);
      }
      float4 xp, x0;
      update_elem<SAMPLER>(oc.x, ou.x, has_u, scale, xt.x, m.x, has_inp, xin.x, clip, eps.x, has_noise, mask_noise, k, xp.x, x0.x);
      update_elem<SAMPLER>(oc.y, ou.y, has_u, scale, xt.y, m.y, has_inp, xin.y, clip, eps.y, has_noise, mask_noise, k, xp.y, x0.y);
      update_elem<SAMPLER>(oc.z, ou.z, has_u, scale, xt.z, m.z, has_inp, xin.z, clip, eps.z, has_noise, mask_noise, k, xp.z, x0.z);
      update_elem<SAMPLER>(oc.w, ou.w, has_u, scale, xt.w, m.w, has_inp, xin.w, clip, eps.w, has_noise, mask_noise, k, xp.w, x0.w);
      *reinterpret_cast<float4*>(a.x_prev + i) = xp;
      if (a.pred_xstart) __stcs(reinterpret_cast<float4*>(a.pred_xstart + i), x0);
    } else {
      float oc = a.out_cond[i];
      float ou = has_u ? a.out_uncond[i] : 0.0f;
      float xt = a.x_t[i];
      float xin = has_inp ? a.x_inpaint[i] : 0.0f;
      float m = 0.0f;
      if (a.mask_kind == MST_MASK_FULL) m = a.mask[i];
      else if (a.mask_kind == MST_MASK_FT) m = a.mask[r];
      else if (a.mask_kind == MST_MASK_F) m = a.mask[f];
      float eps = 0.0f;
      if (NOISE == MST_NOISE_TENSOR) {
        eps = a.noise[a.const_noise ? r : i];
      } else if (NOISE == MST_NOISE_PHILOX) {
        uint64_t sample = a.const_noise ? a.philox_sample_offset : a.philox_sample_offset + (uint64_t)b;
        float4 e4 = philox_normal4(a.philox_seed, sample, (uint32_t)(r >> 2), t);
        int lane = (int)(r & 3);
        eps = lane == 0 ? e4.x : lane == 1 ? e4.y : lane == 2 ? e4.z : e4.w;
      }
      float xp, x0;
      update_elem<SAMPLER>(oc, ou, has_u, scale, xt, m, has_inp, xin, clip, eps, has_noise, mask_noise, k, xp, x0);
      a.x_prev[i] = xp;
      if (a.pred_xstart) a.pred_xstart[i] = x0;
    }
  }

  if (a.advance_t && a.t_scalar_dev) {
    // last block to finish moves the shared timestep down by one so that a
    // captured CUDA graph of one step can be replayed for the next step.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      int done = atomicAdd(a.block_counter, 1);
      if (done == (int)gridDim.x - 1) {
        *a.t_scalar_dev = t_shared - 1;
        *a.block_counter = 0;
        __threadfence();
      }
    }
  }
}

template <int SAMPLER, int NOISE>
static int launch_update(const mst_update_args& a, cudaStream_t s) {
  const int64_t per = (int64_t)a.n_feats * a.n_frames;
  const bool vec = (per % 4 == 0) && (a.n_frames % 4 == 0);
  const int64_t units = (vec ? per / 4 : per) * a.batch;
  const int threads = 256;
  int64_t blocks = (units + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec)
    MST_CUDA_OK(launch_pdl(update_kernel<SAMPLER, NOISE, true>, dim3((unsigned)blocks), dim3(threads), 0, s, a));
  else
    MST_CUDA_OK(launch_pdl(update_kernel<SAMPLER, NOISE, false>, dim3((unsigned)blocks), dim3(threads), 0, s, a));
  MST_LAUNCHED("update", s);
  return MST_OK;
}

// ------------------------------- q_sample ----------------------------------
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       int mask_kind, const float* __restrict__ mask,
                                                       const int64_t* __restrict__ t_vec, int t_imm,
                                                       const float* __restrict__ sa, const float* __restrict__ sb,
                                                       float* __restrict__ xt, int B, int F, int T) {
  const int64_t per = (int64_t)F * T, total = per * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / per);
    int64_t r = i - (int64_t)b * per;
    int t = t_vec ? (int)t_vec[b] : t_imm;
    float m = 0.0f;
    if (mask_kind == MST_MASK_FULL) m = mask[i];
    else if (mask_kind == MST_MASK_FT) m = mask[r];
    else if (mask_kind == MST_MASK_F) m = mask[r / T];
    float e = noise[i];
    if (mask_kind != MST_MASK_NONE) e = __fmul_rn(e, __fsub_rn(1.0f, m));
    xt[i] = __fadd_rn(__fmul_rn(sa[t], x0[i]), __fmul_rn(sb[t], e));
  }
}

__global__ void __launch_bounds__(256) cfg_combine_kernel(const float* __restrict__ oc, const float* __restrict__ ou,
                                                          const float* __restrict__ scale, float* __restrict__ out,
                                                          int B, int64_t per) {
  const int64_t total = per * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / per);
    float u = ou[i];
    out[i] = __fadd_rn(u, __fmul_rn(scale[b], __fsub_rn(oc[i], u)));
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, int B, int64_t per, uint64_t seed,
                                                            uint64_t sample_offset, int t) {
  const int64_t nvec = (per + 3) / 4, total = nvec * B;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(u / nvec);
    int64_t v = u - (int64_t)b * nvec;
    float4 e = philox_normal4(seed, sample_offset + (uint64_t)b, (uint32_t)v, t);
    float vals[4] = {e.x, e.y, e.z, e.w};
    for (int j = 0; j < 4; ++j) {
      int64_t r = v * 4 + j;
      if (r < per) out[(int64_t)b * per + r] = vals[j];
    }
  }
}

static int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mst

using namespace mst;

extern "C" int mst_update_step(const mst_update_args* ap, void* stream) {
  MST_CHECK_ARG(ap != nullptr, "null args");
  const mst_update_args& a = *ap;
  MST_CHECK_ARG(a.batch > 0 && a.n_feats > 0 && a.n_frames > 0, "empty shape");
  MST_CHECK_ARG(a.out_cond && a.x_t && a.x_prev, "out_cond, x_t and x_prev are required");
  MST_CHECK_ARG(a.coef1 && a.coef2, "coef tables are required");
  MST_CHECK_ARG(a.sampler == MST_SAMPLER_DDPM || a.sampler == MST_SAMPLER_DDIM, "unknown sampler");
  MST_CHECK_ARG(a.sampler != MST_SAMPLER_DDIM || (a.recip && a.recipm1), "DDIM needs recip/recipm1 tables");
  MST_CHECK_ARG(a.noise_kind == MST_NOISE_NONE || a.sigma, "sigma table required when noise is added");
  MST_CHECK_ARG(a.noise_kind != MST_NOISE_TENSOR || a.noise, "noise tensor missing");
  MST_CHECK_ARG(a.mask_kind == MST_MASK_NONE || a.mask, "mask pointer missing");
  MST_CHECK_ARG(a.mask_kind >= MST_MASK_NONE && a.mask_kind <= MST_MASK_F, "unknown mask kind");
  MST_CHECK_ARG(!a.out_uncond || a.cfg_scale, "cfg_scale required with out_uncond");
  MST_CHECK_ARG(!a.advance_t || (a.t_scalar_dev && a.block_counter), "advance_t needs t_scalar_dev and block_counter");
  MST_CHECK_ARG(!(a.t_vec && a.t_scalar_dev), "give t_vec or t_scalar_dev, not both");
  cudaStream_t s = (cudaStream_t)stream;
#define MST_DISPATCH(S)                                                          \
  switch (a.noise_kind) {                                                        \
    case MST_NOISE_NONE: return launch_update<S, MST_NOISE_NONE>(a, s);          \
    case MST_NOISE_TENSOR: return launch_update<S, MST_NOISE_TENSOR>(a, s);      \
    case MST_NOISE_PHILOX: return launch_update<S, MST_NOISE_PHILOX>(a, s);      \
    default: return fail(MST_ERR_INVALID, "mst_update_step: unknown noise kind"); \
  }
  if (a.sampler == MST_SAMPLER_DDPM) {
    MST_DISPATCH(MST_SAMPLER_DDPM)
  } else {
    MST_DISPATCH(MST_SAMPLER_DDIM)
  }
#undef MST_DISPATCH
}

extern "C" int mst_q_sample(const float* x_start, const float* noise, int32_t mask_kind, const float* mask,
                            const int64_t* t_vec, int32_t t_imm, const float* sqrt_ab, const float* sqrt_1m_ab,
                            float* x_t, int32_t batch, int32_t n_feats, int32_t n_frames, void* stream) {
  MST_CHECK_ARG(x_start && noise && x_t && sqrt_ab && sqrt_1m_ab, "null pointer");
  MST_CHECK_ARG(batch > 0 && n_feats > 0 && n_frames > 0, "empty shape");
  MST_CHECK_ARG(mask_kind == MST_MASK_NONE || mask, "mask pointer missing");
  int64_t total = (int64_t)batch * n_feats * n_frames;
  q_sample_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x_start, noise, mask_kind, mask, t_vec, t_imm,
                                                                     sqrt_ab, sqrt_1m_ab, x_t, batch, n_feats, n_frames);
  MST_LAUNCHED("q_sample", (cudaStream_t)stream);
  return MST_OK;
}

extern "C" int mst_cfg_combine(const float* out_cond, const float* out_uncond, const float* scale, float* out,
                               int32_t batch, int64_t per_sample, void* stream) {
  MST_CHECK_ARG(out_cond && out_uncond && scale && out, "null pointer");
  MST_CHECK_ARG(batch > 0 && per_sample > 0, "empty shape");
  cfg_combine_kernel<<<grid_for((int64_t)batch * per_sample), 256, 0, (cudaStream_t)stream>>>(out_cond, out_uncond, scale,
                                                                                            out, batch, per_sample);
  MST_LAUNCHED("cfg_combine", (cudaStream_t)stream);
  return MST_OK;
}

extern "C" int mst_philox_normal(float* out, int32_t batch, int64_t per_sample, uint64_t seed, uint64_t sample_offset,
                                 int32_t t, void* stream) {
  MST_CHECK_ARG(out != nullptr && batch > 0 && per_sample > 0, "bad arguments");
  philox_normal_kernel<<<grid_for((int64_t)batch * ((per_sample + 3) / 4)), 256, 0, (cudaStream_t)stream>>>(
      out, batch, per_sample, seed, sample_offset, t);
  MST_LAUNCHED("philox_normal", (cudaStream_t)stream);
  return MST_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Post-sampling decode: inv_transform (x * std + mean, data_loaders/humanml/data/dataset.py:478-479)
// + recover_from_ric (data_loaders/humanml/scripts/motion_process.py:389-411, :444-461) fused, reading the sampler's
// own [B, F, 1, T] layout (the reference permutes to [B, 1, T, F] on the CPU first, sample/demo_style_transfer.py:265).
//   r_rot_ang[t] = sum_{s<t} rot_vel[s];  q = (cos, 0, sin, 0);  r_pos = cumsum_t qrot(q, (vx[t-1], 0, vz[t-1])), y = root_y
//   joint j >= 1: qrot(q, ric_j) + (r_pos.x, 0, r_pos.z);  joint 0 = r_pos.           out: [B, T, J, 3]
// One CTA per sample: the three prefix sums are a 3 x T sequential scan by one thread in the reference's summation
// order (T <= a few hundred), the joints are elementwise.
// ---------------------------------------------------------------------------------------------------------
namespace mst {

__device__ __forceinline__ void qrot_y(float c, float sn, float vx, float vy, float vz, float& ox, float& oy, float& oz) {
  // qrot(q, v) with q = (c, 0, sn, 0) written as the reference's two cross products (quaternion.py:88-99), u = (0, sn, 0)
  const float uvx = sn * vz - 0.0f * vy, uvy = 0.0f * vx - 0.0f * vz, uvz = 0.0f * vy - sn * vx;
  const float uuvx = sn * uvz - 0.0f * uvy, uuvy = 0.0f * uvx - 0.0f * uvz, uuvz = 0.0f * uvy - sn * uvx;
  ox = vx + 2.0f * (c * uvx + uuvx);
  oy = vy + 2.0f * (c * uvy + uuvy);
  oz = vz + 2.0f * (c * uvz + uuvz);
}

__global__ void __launch_bounds__(256) recover_from_ric_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                               const float* __restrict__ stdv, float* __restrict__ out, int F,
                                                               int T, int J) {
  extern __shared__ float sm[];
  float* ang = sm;          // [T] r_rot_ang, then cos
  float* sn = sm + T;       // [T] sin
  float* px = sm + 2 * T;   // [T] root x
  float* pz = sm + 3 * T;   // [T] root z
  const int b = blockIdx.x;
  const float* xb = x + (size_t)b * F * T;
  auto feat = [&](int f, int t) { return mean ? xb[(size_t)f * T + t] * stdv[f] + mean[f] : xb[(size_t)f * T + t]; };
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    ang[t] = t == 0 ? 0.0f : feat(0, t - 1);
    px[t] = t == 0 ? 0.0f : feat(1, t - 1);
    pz[t] = t == 0 ? 0.0f : feat(2, t - 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // torch.cumsum order
    float a = 0.0f;
    for (int t = 0; t < T; ++t) { a += ang[t]; ang[t] = a; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float a = ang[t];
    const float c = cosf(a), s_ = sinf(a);
    float ox, oy, oz;
    qrot_y(c, s_, px[t], 0.0f, pz[t], ox, oy, oz);
    ang[t] = c;
    sn[t] = s_;
    px[t] = ox;
    pz[t] = oz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ax = 0.0f, az = 0.0f;
    for (int t = 0; t < T; ++t) { ax += px[t]; az += pz[t]; px[t] = ax; pz[t] = az; }
  }
  __syncthreads();
  float* ob = out + (size_t)b * T * J * 3;
  for (int i = threadIdx.x; i < T * J; i += blockDim.x) {
    const int j = i / T, t = i - j * T;  // t fastest: coalesced reads of x[b][f][t]
    float* o = ob + ((size_t)t * J + j) * 3;
    if (j == 0) {
      o[0] = px[t]; o[1] = feat(3, t); o[2] = pz[t];
    } else {
      const int f = 4 + 3 * (j - 1);
      float ox, oy, oz;
      qrot_y(ang[t], sn[t], feat(f, t), feat(f + 1, t), feat(f + 2, t), ox, oy, oz);
      o[0] = ox + px[t]; o[1] = oy; o[2] = oz + pz[t];
    }
  }
}

}  // namespace mst

extern "C" int mst_recover_from_ric(const float* x, const float* mean, const float* stdv, float* joints, int32_t batch,
                                    int32_t n_feats, int32_t n_frames, int32_t joints_num, void* stream) {
  MST_CHECK_ARG(x && joints, "null tensor pointer");
  MST_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "give both mean and std, or neither");
  MST_CHECK_ARG(batch > 0 && n_frames > 0 && joints_num > 0, "non-positive size");
  MST_CHECK_ARG(n_feats >= 4 + 3 * (joints_num - 1), "n_feats too small for joints_num (needs 4 + 3 (J - 1) features)");
  const size_t smem = (size_t)4 * n_frames * sizeof(float);
  MST_CHECK_ARG(smem <= 48 * 1024, "sequence too long for the decode kernel (n_frames <= 3072)");
  cudaStream_t s = (cudaStream_t)stream;
  mst::recover_from_ric_kernel<<<batch, 256, smem, s>>>(x, mean, stdv, joints, n_feats, n_frames, joints_num);
  MST_LAUNCHED("recover_from_ric", s);
  return MST_OK;
}
