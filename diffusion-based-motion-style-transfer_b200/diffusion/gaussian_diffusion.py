"""Host side of the Gaussian diffusion process, B200-native.

Mirrors the public surface of the reference's ``diffusion/gaussian_diffusion.py``
(class and method names, argument order, return conventions) for the sampling
hot path, so callers such as ``sample/demo_style_transfer.py:244-258`` and
``train/finetune_style_diffusion.py:195-211`` keep working unchanged.  The
work underneath is different:

* schedule tables are built once on the host in float64 exactly as the
  reference does (``gaussian_diffusion.py:183-219``) and uploaded ONCE per
  device as fp32 (the reference re-uploads a float64 table five times per
  step, ``:1615``);
* every per-step elementwise op (CFG lerp, inpainting blend, clamp, posterior
  mean, masked noise) is one fused CUDA kernel (``mst_update_step``);
* a native denoiser (``model.mdm_forstyledataset.MDM`` / ``StyleDiffusion``,
  optionally inside ``ClassifierFreeSampleModel``) is run through the
  hand-written sm_100a forward with cond+uncond batched, text/time embeddings
  hoisted out of the loop, and the step replayed as a CUDA graph.

Nothing here runs on the CPU: tensors must be CUDA tensors.
"""
from __future__ import annotations

import enum
import math
import contextlib
import os
from copy import deepcopy

import numpy as np
import torch
import torch as th

from .. import _lib as L
from .. import engine as K


# ----------------------------------------------------------------------------
# schedules (reference gaussian_diffusion.py:22-66)
# ----------------------------------------------------------------------------
def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.):
    """'linear' (Ho et al., rescaled to the step count) or 'cosine' (Nichol & Dhariwal)."""
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(
            num_diffusion_timesteps,
            lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2,
        )
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """beta_i = min(1 - abar((i+1)/N) / abar(i/N), max_beta)."""
    n = num_diffusion_timesteps
    out = [min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)]
    return np.array(out)


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """fp32 gather of a host table, broadcast to ``broadcast_shape`` (reference :1605-1618).
    Kept for API compatibility; the fused kernels never call it."""
    res = th.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


def _require_cuda(t, name):
    if not isinstance(t, th.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the mst sampler has no CPU path")


def _f32c(t):
    """fp32 + contiguous view of a CUDA tensor (copy only when needed)."""
    if t.dtype != th.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _MaskedL2Fn(th.autograd.Function):
    """masked_l2 rows with an explicit CUDA backward w.r.t. ``b`` (the stacked x0 predictions)."""

    @staticmethod
    def forward(ctx, a, b, mask):
        ctx.save_for_backward(a, b, mask)
        return K.masked_l2_forward(a, b, mask)

    @staticmethod
    def backward(ctx, g):
        a, b, mask = ctx.saved_tensors
        return None, K.masked_l2_backward(a, b, mask, _f32c(g)), None


class _UpdateStepFn(th.autograd.Function):
    """Fused per-step update with a gradient path x0-prediction -> model output (what
    p_sample_with_grad / ddim_sample_with_grad keep in the graph, inpainting_gaussian_diffusion.py:66-123, :176-239).
    ``sample`` is produced outside the graph there (under the loop's no_grad), so only pred_xstart is differentiable."""

    @staticmethod
    def forward(ctx, out, run, mask, clip_denoised, shape):
        sample, x0 = run(out)
        ctx.mask, ctx.clip, ctx.shape = mask, bool(clip_denoised), tuple(shape)
        ctx.x0 = x0 if clip_denoised else None
        ctx.mark_non_differentiable(sample)
        return sample, x0

    @staticmethod
    def backward(ctx, d_sample, d_x0):
        d_out = K.update_step_backward(_f32c(d_x0), None, None, None, ctx.mask, ctx.x0, ctx.clip, ctx.shape)
        return d_out, None, None, None, None


class GaussianDiffusion:
    """Sampling utilities of the diffusion process (reference class of the same name, :111)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False,
                 lambda_rcxyz=0., lambda_vel=0., lambda_pose=1., lambda_orient=1., lambda_loc=1., data_rep='rot6d',
                 lambda_root_vel=0., lambda_vel_rcxyz=0., lambda_fc=0., lambda_sty_cons=0., lambda_sty_trans=0.,
                 lambda_cont_pers=0., lambda_cont_vel=0., lambda_diff_sty=0., lambda_l1=10.):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        self.data_rep = data_rep
        if data_rep != 'rot_vel' and lambda_pose != 1.:
            raise ValueError('lambda_pose is relevant only when training on velocities!')
        self.lambda_pose, self.lambda_orient, self.lambda_loc = lambda_pose, lambda_orient, lambda_loc
        self.lambda_rcxyz, self.lambda_vel, self.lambda_root_vel = lambda_rcxyz, lambda_vel, lambda_root_vel
        self.lambda_vel_rcxyz, self.lambda_fc, self.lambda_l1 = lambda_vel_rcxyz, lambda_fc, lambda_l1
        self.lambda_sty_cons, self.lambda_sty_trans = lambda_sty_cons, lambda_sty_trans
        self.lambda_cont_pers, self.lambda_cont_vel, self.lambda_diff_sty = lambda_cont_pers, lambda_cont_vel, lambda_diff_sty
        if (lambda_rcxyz > 0. or lambda_vel > 0. or lambda_root_vel > 0. or lambda_vel_rcxyz > 0. or lambda_fc > 0.):
            assert self.loss_type == LossType.MSE, 'Geometric losses are supported by MSE loss type only!'

        # float64 host tables, same formulas and evaluation order as the reference (:183-219)
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        assert self.alphas_cumprod_prev.shape == (self.num_timesteps,)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

        self.l2_loss = lambda a, b: (a - b) ** 2

        # --- B200 runtime knobs (not part of the reference API) ---
        # rng: 'torch'  -> th.randn_like per step from the global generator, like the reference
        #      'philox' -> noise generated inside the update kernel, keyed by (seed, global sample, t)
        self.rng = os.environ.get("MST_RNG", "torch")
        self.philox_seed = int(os.environ.get("MST_SEED", "0"))
        self.philox_sample_offset = 0          # global index of this shard's first sample (multi-GPU)
        self.noise_fn = None                   # optional callable(step_number, shape, device) -> eps tensor
        self.use_cuda_graph = os.environ.get("MST_GRAPH", "1") != "0"
        # How a fused trajectory whose caller only wants the final sample is submitted: "full" = ALL its denoise steps
        # captured as ONE CUDA graph and launched once (the timestep lives in device memory and the update kernel
        # decrements it, so the N steps are N copies of the same node sequence); "step" = one captured step replayed N
        # times from the host; an integer K = graphs of K steps.  Graphs above MST_TRAJ_GRAPH_MAX_NODES kernel nodes fall
        # back to chunks.
        self.trajectory_graph = os.environ.get("MST_TRAJ_GRAPH", "full")
        self.steps_per_graph = int(os.environ.get("MST_STEPS_PER_GRAPH", "1"))
        self._dev = {}
        self._graphs = {}
        self.last_plan_launches = 0        # kernels per denoise step of the most recent fused trajectory

    # ------------------------------------------------------------------ tables
    def _variance_tables(self):
        if self.model_var_type == ModelVarType.FIXED_SMALL:
            return self.posterior_variance, self.posterior_log_variance_clipped
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            v = np.append(self.posterior_variance[1], self.betas[1:])
            return v, np.log(v)
        raise NotImplementedError(
            f"{self.model_var_type}: learned variances are not on the reference's hot path "
            "(utils/model_util.py:176 fixes learn_sigma=False)")

    def device_tables(self, device, eta=0.0):
        """fp32 device copies of the per-timestep coefficients (uploaded once per device/eta)."""
        key = (str(device), float(eta))
        tabs = self._dev.get(key)
        if tabs is None:
            var, logvar = self._variance_tables()
            ab, abp = self.alphas_cumprod, self.alphas_cumprod_prev
            ddim_sigma = eta * np.sqrt((1 - abp) / (1 - ab)) * np.sqrt(1 - ab / abp)
            host = {
                "c1": self.posterior_mean_coef1, "c2": self.posterior_mean_coef2,
                "sigma": np.exp(0.5 * logvar), "var": var, "logvar": logvar,
                "sqrt_ab": self.sqrt_alphas_cumprod, "sqrt_1m_ab": self.sqrt_one_minus_alphas_cumprod,
                "recip": self.sqrt_recip_alphas_cumprod, "recipm1": self.sqrt_recipm1_alphas_cumprod,
                "ddim_c1": np.sqrt(abp), "ddim_c2": np.sqrt(1 - abp - ddim_sigma ** 2), "ddim_sigma": ddim_sigma,
                "ab": ab, "abp": abp,
            }
            tabs = {k: th.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(th.float32).to(device)
                    for k, v in host.items()}
            self._dev[key] = tabs
        return tabs

    # ------------------------------------------------------------------ small API
    def masked_l2(self, a, b, mask):
        """sum((a-b)^2 * mask) / (sum(mask) * J*Jdim) per sample (reference :223-235)."""
        _require_cuda(b, "masked_l2 input")
        assert a.shape[1:] == b.shape[1:] and mask.shape[-1] == b.shape[-1]
        if a.requires_grad and th.is_grad_enabled():
            raise NotImplementedError("masked_l2 is differentiable w.r.t. its second argument only (the reference "
                                      "passes the target first, gaussian_diffusion.py:1381)")
        # expand() views of the reference (target / mask repeated over the stacked steps) are read modulo their rows
        a_rows = a[:1] if (a.shape[0] > 1 and a.stride(0) == 0) else a
        m_rows = mask[:1] if (mask.shape[0] > 1 and mask.stride(0) == 0) else mask
        assert b.shape[0] % a_rows.shape[0] == 0 and b.shape[0] % m_rows.shape[0] == 0
        return _MaskedL2Fn.apply(_f32c(a_rows), _f32c(b), _f32c(m_rows.reshape(m_rows.shape[0], 1, 1, -1)))

    def q_mean_variance(self, x_start, t):
        mean = _extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = _extract_into_tensor(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = _extract_into_tensor(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def _inpainting_mask_for_noise(self, model_kwargs):
        """Mask applied to the noise; the base process applies none (InpaintingGaussianDiffusion overrides)."""
        return None

    def q_sample(self, x_start, t, noise=None, model_kwargs=None):
        """x_t ~ q(x_t | x_0): sqrt(abar_t) x_0 + sqrt(1-abar_t) eps (reference :267-285)."""
        _require_cuda(x_start, "x_start")
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        tabs = self.device_tables(x_start.device)
        mask = self._inpainting_mask_for_noise(model_kwargs)
        if mask is not None:
            mask = _f32c(mask)
        out = K.q_sample(_f32c(x_start), _f32c(noise), mask, t.to(th.int64).contiguous(), 0, tabs["sqrt_ab"],
                         tabs["sqrt_1m_ab"])
        if mask is not None:
            # the reference mutates its `noise` argument in place (inpainting_gaussian_diffusion.py:18);
            # callers can observe that through `noise=` / `img`, so it is reproduced
            noise.mul_(1. - mask)
        return out

    def q_posterior_mean_variance(self, x_start, x_t, t):
        assert x_start.shape == x_t.shape
        tabs = self.device_tables(x_t.device)
        mean = th.empty_like(x_t, dtype=th.float32)
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=_f32c(x_start), x_t=_f32c(x_t), x_prev=mean,
                      coef1=tabs["c1"], coef2=tabs["c2"], t_vec=t.to(th.int64).contiguous(), noise_kind=L.NOISE_NONE)
        pv = _extract_into_tensor(self.posterior_variance, t, x_t.shape)
        plv = _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, pv, plv

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _predict_xstart_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return ((_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - pred_xstart)
                / _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    # ------------------------------------------------------------------ model plumbing
    def _wrap_model(self, model):
        return model

    def _model_time_index(self, t_index):
        """timestep handed to the denoiser for process index t (identity here; respaced subclasses remap)."""
        return t_index

    def _map_model_t(self, t):
        """tensor version of _model_time_index (what _WrappedModel.__call__ computes, respace.py:129-134)."""
        return self._scale_timesteps(t)

    @staticmethod
    def _unwrap(model):
        """(inner native denoiser or None, cfg wrapper or None)."""
        from ..model.cfg_sampler import ClassifierFreeSampleModel
        from ..model.mdm_forstyledataset import NativeDenoiser
        m = getattr(model, "model", None) if type(model).__name__ == "_WrappedModel" else None
        m = m if m is not None else model
        if isinstance(m, ClassifierFreeSampleModel):
            inner = m.model
            return (inner if isinstance(inner, NativeDenoiser) else None), m
        return (m if isinstance(m, NativeDenoiser) else None), None

    def _check_inpainting(self, model_kwargs, shape):
        y = model_kwargs['y']  # KeyError when absent, like the reference (:341)
        if 'inpainting_mask' in y.keys() and 'inpainted_motion' in y.keys():
            assert self.model_mean_type == ModelMeanType.START_X, 'This feature supports only X_start pred for mow!'
            mask, inp = y['inpainting_mask'], y['inpainted_motion']
            assert tuple(shape) == tuple(mask.shape) == tuple(inp.shape)
            return _f32c(mask), _f32c(inp)
        return None, None

    def _model_outputs(self, model, x, t, model_kwargs):
        """Run the denoiser.  Returns (out_cond, out_uncond or None, cfg_scale or None)."""
        native, cfgw = self._unwrap(model)
        if native is not None and native.mst_ready(x):
            y = model_kwargs['y']
            t_model = self._map_model_t(t)
            if cfgw is not None:
                oc, ou = native.forward_cfg(x, t_model, y)
                return oc, ou, _f32c(y['scale'].to(x.device)).view(-1)
            return _f32c(native(x, t_model, **model_kwargs)), None, None
        out = self._wrap_model(model)(x, self._scale_timesteps(t), **model_kwargs)
        return _f32c(out), None, None

    # ------------------------------------------------------------------ one step
    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """p(x_{t-1} | x_t) and the x_0 prediction (reference :311-424): dict with
        'mean', 'variance', 'log_variance', 'pred_xstart'."""
        if model_kwargs is None:
            model_kwargs = {}
        _require_cuda(x, "x")
        B, C = x.shape[:2]
        assert t.shape == (B,)
        if self.model_mean_type != ModelMeanType.START_X:
            raise NotImplementedError(
                f"{self.model_mean_type}: only START_X is on the reference's hot path (utils/model_util.py:173)")
        x = _f32c(x)
        oc, ou, scale = self._model_outputs(model, x, t, model_kwargs)
        mask, inp = self._check_inpainting(model_kwargs, x.shape)
        tabs = self.device_tables(x.device)
        mean, x0 = th.empty_like(x), th.empty_like(x)
        if denoised_fn is not None:
            # arbitrary user hook: materialise x0 first (blend only), apply, then finish the posterior
            K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=mean,
                          pred_xstart=x0, mask=mask, x_inpaint=inp, coef1=tabs["c1"], coef2=tabs["c2"],
                          t_vec=t.to(th.int64).contiguous(), clip_denoised=False)
            oc, ou, scale, mask_k, inp_k = _f32c(denoised_fn(x0)), None, None, None, None
        else:
            mask_k, inp_k = mask, inp
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=mean,
                      pred_xstart=x0, mask=mask_k, x_inpaint=inp_k, coef1=tabs["c1"], coef2=tabs["c2"],
                      t_vec=t.to(th.int64).contiguous(), clip_denoised=clip_denoised)
        shape4 = (-1,) + (1,) * (x.dim() - 1)
        tl = t.to(th.int64)
        out = {
            "mean": mean,
            "variance": tabs["var"][tl].view(shape4).expand(x.shape),
            "log_variance": tabs["logvar"][tl].view(shape4).expand(x.shape),
            "pred_xstart": x0,
        }
        assert out["mean"].shape == out["log_variance"].shape == out["pred_xstart"].shape == x.shape
        return out

    def condition_mean(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        gradient = cond_fn(x, self._scale_timesteps(t), **model_kwargs)
        return p_mean_var["mean"].float() + p_mean_var["variance"] * gradient.float()

    def _draw_noise(self, x, step_no, const_noise=False):
        if self.noise_fn is not None:
            eps = self.noise_fn(step_no, tuple(x.shape), x.device)
            _require_cuda(eps, "noise_fn result")
            eps = _f32c(eps)
        else:
            eps = th.randn_like(x)
        if const_noise:
            eps = eps[[0]].repeat(x.shape[0], 1, 1, 1)
        return eps

    def _sample_step(self, sampler, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, const_noise=False,
                     eta=0.0, step_no=0):
        """One reverse step with everything after the denoiser fused (p_sample :532 / ddim_sample :796)."""
        if model_kwargs is None:
            model_kwargs = {}
        _require_cuda(x, "x")
        if self.model_mean_type != ModelMeanType.START_X:
            raise NotImplementedError(f"{self.model_mean_type}: only START_X is supported")
        if cond_fn is not None or denoised_fn is not None:
            return self._sample_step_unfused(sampler, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                                             const_noise, eta, step_no)
        x = _f32c(x)
        assert t.shape == (x.shape[0],)
        oc, ou, scale = self._model_outputs(model, x, t, model_kwargs)
        mask, inp = self._check_inpainting(model_kwargs, x.shape)
        nmask = self._inpainting_mask_for_noise(model_kwargs)
        tabs = self.device_tables(x.device, eta)
        sample, x0 = th.empty_like(x), th.empty_like(x)
        if self.rng == "philox" and self.noise_fn is None:
            noise_kind, eps = L.NOISE_PHILOX, None
        else:
            noise_kind, eps = L.NOISE_TENSOR, self._draw_noise(x, step_no, const_noise)
            const_noise = False  # already materialised
        kmask = mask if mask is not None else (_f32c(nmask) if nmask is not None else None)
        common = dict(out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=sample, pred_xstart=x0, mask=kmask,
                      x_inpaint=inp, mask_noise=nmask is not None, clip_denoised=clip_denoised,
                      t_vec=t.to(th.int64).contiguous(), noise_kind=noise_kind, noise=eps, const_noise=const_noise,
                      philox_seed=self.philox_seed, philox_sample_offset=self.philox_sample_offset)
        if sampler == L.SAMPLER_DDPM:
            K.update_step(sampler=sampler, coef1=tabs["c1"], coef2=tabs["c2"], sigma=tabs["sigma"], **common)
        else:
            K.update_step(sampler=sampler, coef1=tabs["ddim_c1"], coef2=tabs["ddim_c2"], sigma=tabs["ddim_sigma"],
                          recip=tabs["recip"], recipm1=tabs["recipm1"], **common)
        return {"sample": sample, "pred_xstart": x0}

    def _sample_step_unfused(self, sampler, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, const_noise,
                             eta, step_no):
        """cond_fn / denoised_fn hooks (unused by every reference caller): p_mean_variance + torch glue."""
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        eps = self._draw_noise(x, step_no, const_noise)
        nmask = self._inpainting_mask_for_noise(model_kwargs)
        if nmask is not None:
            eps = eps * (1. - nmask)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        if sampler == L.SAMPLER_DDPM:
            if cond_fn is not None:
                out["mean"] = self.condition_mean(cond_fn, out, x, t, model_kwargs=model_kwargs)
            sample = out["mean"] + nonzero * th.exp(0.5 * out["log_variance"]) * eps
            return {"sample": sample, "pred_xstart": out["pred_xstart"]}
        x0 = out["pred_xstart"]
        if cond_fn is not None:
            alpha_bar = _extract_into_tensor(self.alphas_cumprod, t, x.shape)
            e = self._predict_eps_from_xstart(x, t, x0)
            e = e - (1 - alpha_bar).sqrt() * cond_fn(x, self._scale_timesteps(t), **model_kwargs)
            x0 = self._predict_xstart_from_eps(x, t, e)
        e = self._predict_eps_from_xstart(x, t, x0)
        alpha_bar = _extract_into_tensor(self.alphas_cumprod, t, x.shape)
        alpha_bar_prev = _extract_into_tensor(self.alphas_cumprod_prev, t, x.shape)
        sigma = eta * th.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) * th.sqrt(1 - alpha_bar / alpha_bar_prev)
        mean_pred = x0 * th.sqrt(alpha_bar_prev) + th.sqrt(1 - alpha_bar_prev - sigma ** 2) * e
        return {"sample": mean_pred + nonzero * sigma * eps, "pred_xstart": out["pred_xstart"]}

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 const_noise=False, pred_xstart_in_graph=False):
        """x_{t-1} ~ p(x_{t-1} | x_t) (reference :532-585); returns {'sample', 'pred_xstart'}."""
        return self._sample_step(L.SAMPLER_DDPM, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                                 const_noise=const_noise)

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0,
                    pred_xstart_in_graph=False):
        """One DDIM step (reference :796-847)."""
        return self._sample_step(L.SAMPLER_DDIM, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, eta=eta)

    def _sample_step_with_grad(self, sampler, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                               const_noise=False, eta=0.0, step_no=0, pred_xstart_in_graph=False):
        """p_sample_with_grad / ddim_sample_with_grad (inpainting_gaussian_diffusion.py:66-123, :176-239): the
        denoiser runs with an activation tape and pred_xstart stays attached to it; x is detached at entry."""
        if model_kwargs is None:
            model_kwargs = {}
        if cond_fn is not None or denoised_fn is not None:
            raise NotImplementedError("cond_fn / denoised_fn hooks are not supported on the differentiable sampling "
                                      "path (no reference caller passes them)")
        _require_cuda(x, "x")
        native, cfgw = self._unwrap(model)
        if native is None or cfgw is not None:
            raise NotImplementedError("differentiable sampling needs a native denoiser without the CFG wrapper "
                                      "(the reference finetunes the unwrapped StyleDiffusion)")
        x = _f32c(x.detach())
        assert t.shape == (x.shape[0],)
        mask, inp = self._check_inpainting(model_kwargs, x.shape)
        nmask = self._inpainting_mask_for_noise(model_kwargs)
        tabs = self.device_tables(x.device, eta)
        if self.rng == "philox" and self.noise_fn is None:
            noise_kind, eps = L.NOISE_PHILOX, None
        else:
            noise_kind, eps = L.NOISE_TENSOR, self._draw_noise(x, step_no, const_noise)
            const_noise = False
        kmask = mask if mask is not None else (_f32c(nmask) if nmask is not None else None)
        tl = t.to(th.int64).contiguous()

        def run(out):
            sample, x0 = th.empty_like(x), th.empty_like(x)
            common = dict(out_cond=_f32c(out), x_t=x, x_prev=sample, pred_xstart=x0, mask=kmask, x_inpaint=inp,
                          mask_noise=nmask is not None, clip_denoised=clip_denoised, t_vec=tl, noise_kind=noise_kind,
                          noise=eps, const_noise=const_noise, philox_seed=self.philox_seed,
                          philox_sample_offset=self.philox_sample_offset)
            if sampler == L.SAMPLER_DDPM:
                K.update_step(sampler=sampler, coef1=tabs["c1"], coef2=tabs["c2"], sigma=tabs["sigma"], **common)
            else:
                K.update_step(sampler=sampler, coef1=tabs["ddim_c1"], coef2=tabs["ddim_c2"], sigma=tabs["ddim_sigma"],
                              recip=tabs["recip"], recipm1=tabs["recipm1"], **common)
            return sample, x0

        with th.enable_grad():
            out = native(x, self._map_model_t(t), **model_kwargs)
            # the blend only zeroes the gradient where the inpainting mask is set (mask is None without inpainting)
            sample, x0 = _UpdateStepFn.apply(out, run, mask, clip_denoised, tuple(x.shape))
        if not pred_xstart_in_graph:
            x0 = x0.detach()
        return {"sample": sample.detach(), "pred_xstart": x0}

    def p_sample_with_grad(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                           pred_xstart_in_graph=False, const_noise=False):
        return self._sample_step_with_grad(L.SAMPLER_DDPM, model, x, t, clip_denoised, denoised_fn, cond_fn,
                                           model_kwargs, const_noise=const_noise,
                                           pred_xstart_in_graph=pred_xstart_in_graph)

    def ddim_sample_with_grad(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                              eta=0.0, pred_xstart_in_graph=False):
        return self._sample_step_with_grad(L.SAMPLER_DDIM, model, x, t, clip_denoised, denoised_fn, cond_fn,
                                           model_kwargs, eta=eta, pred_xstart_in_graph=pred_xstart_in_graph)

    def few_shot_style_finetune_losses(self, model, x_start, t, x_content_start, x_style_start, skip_steps=700,
                                       model_kwargs=None, noise=None, model_t2m_kwargs=None, semantic_guidance=0,
                                       use_ddim=0, Ls=10):
        """Few-shot style finetune loss (reference :1317-1399): masked L2 between the style example and every x0
        prediction of a short differentiable sampling run started from the content motion, plus - with
        ``semantic_guidance`` - ``Ls`` x (1 - cos) between the MotionEncoder's mu of the denoised t2m batch and
        the CLIP feature of its (style-word-modified) captions.  Returns {'rot_mse', ['text_cosine'], 'loss'}."""
        native, _ = self._unwrap(model)
        if native is None:
            raise NotImplementedError("few_shot_style_finetune_losses needs a native denoiser (StyleDiffusion)")
        motion_enc = native.controlmdm.motion_enc if hasattr(native, "controlmdm") else native.motion_enc
        mask = model_kwargs['y']['mask']
        if noise is None:
            noise = th.randn_like(x_content_start)
        terms = {}
        mu = text_features = None
        noise_t2m = th.rand_like(x_start)  # sic: uniform noise in the reference (:1334); drawn in the same RNG order
        early_cos = None
        hook = getattr(self, "early_t2m_backward", None) if semantic_guidance else None
        side = getattr(self, "t2m_stream", None) if hook is not None else None
        if semantic_guidance:
            if side is not None:
                # the trainer's hook: this branch (the only one that differs between data-parallel ranks) runs on its own
                # stream, next to the style steps below.  Weight re-packs happen on the main stream first, both streams
                # read them; tensors of the main stream's allocator that the side stream reads are marked for it.
                native.mst_engine(x_start.device, precision=native.mst_train_prec())
                motion_enc.mst_engine(x_start.device, precision=motion_enc.mst_train_prec())
                side.wait_stream(th.cuda.current_stream())
                for tensor in (noise_t2m, t, x_start):
                    tensor.record_stream(side)
            with (th.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                # (the reference also runs this forward with semantic_guidance == 0 and discards the result, :1335-1337)
                x_t = self.q_sample(x_start, t, noise=noise_t2m, model_kwargs=model_t2m_kwargs)
                model_output = native(x_t, self._map_model_t(t), **model_t2m_kwargs)
                mu, text_features = motion_enc(model_output, **model_t2m_kwargs)
                if hook is not None:
                    # back-propagate the term NOW: the trainer accumulates (and, data parallel, all-reduces) its gradient
                    # apart from the style term's, which is still to be computed
                    features_norm = text_features / text_features.norm(dim=-1, keepdim=True)
                    mu_norm = mu / mu.norm(dim=-1, keepdim=True)
                    early_cos = (1 - th.nn.functional.cosine_similarity(features_norm, mu_norm, dim=1, eps=1e-6)).mean()
                    hook(early_cos * Ls)
                    early_cos = early_cos.detach()
                    mu = text_features = model_output = x_t = None
        if not use_ddim:
            sample_fn = self.p_sample_loop
        else:
            sample_fn = self.ddim_sample_loop
            skip_steps = int(skip_steps / 1000 * 20)
        sample = sample_fn(model, x_content_start.shape, clip_denoised=False, model_kwargs=model_kwargs,
                           skip_timesteps=skip_steps, init_image=x_content_start, progress=False, dump_steps=None,
                           noise=None, const_noise=False, cond_fn_with_grad=True, pred_xstart_in_graph=True,
                           dump_all_xstart=True)
        num_step = len(sample)
        sample = th.cat(sample, dim=0)  # noised_step * b_size * J * 1 * seq
        if self.loss_type not in (LossType.MSE, LossType.RESCALED_MSE):
            raise NotImplementedError(self.loss_type)
        assert self.model_mean_type == ModelMeanType.START_X  # only support predict x_0
        target = x_style_start
        assert target.shape == x_content_start.shape
        target = target.expand(num_step, -1, -1, -1)
        mask = mask.expand(num_step, -1, -1, -1)
        terms["rot_mse"] = self.masked_l2(target, sample, mask)
        if semantic_guidance and early_cos is not None:
            # already back-propagated (detached): the values are kept for logging.  They live on the side stream, so the
            # sum with the style term is left to the trainer (after it has joined the streams): 'loss' is the part that
            # still needs a backward pass, 'loss_t2m' the part that had its own
            terms["text_cosine"] = early_cos
            terms["loss_t2m"] = early_cos * Ls
            terms["loss"] = terms["rot_mse"].mean()
            if side is None:
                terms["loss"] = terms["loss"] + terms.pop("loss_t2m")
        elif semantic_guidance:
            features_norm = text_features / text_features.norm(dim=-1, keepdim=True)
            mu_norm = mu / mu.norm(dim=-1, keepdim=True)
            cos = th.nn.functional.cosine_similarity(features_norm, mu_norm, dim=1, eps=1e-6)
            terms["text_cosine"] = (1 - cos).mean()
            terms["loss"] = terms["rot_mse"].mean() + terms["text_cosine"] * Ls
        else:
            terms["loss"] = terms["rot_mse"].mean()
        return terms

    # ------------------------------------------------------------------ loops
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False,
                      pred_xstart_in_graph=False, dump_all_xstart=False, stop_timesteps=None):
        """Full DDPM trajectory (reference :644-715).  Returns the final sample, or a list when
        ``dump_steps`` / ``dump_all_xstart`` is given."""
        final = None
        dump = [] if (dump_steps is not None or dump_all_xstart) else None
        for i, sample in enumerate(self.p_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                model_kwargs=model_kwargs, device=device, progress=progress, skip_timesteps=skip_timesteps,
                init_image=init_image, randomize_class=randomize_class, cond_fn_with_grad=cond_fn_with_grad,
                const_noise=const_noise, pred_xstart_in_graph=pred_xstart_in_graph, stop_timesteps=stop_timesteps,
                _want_xstart=dump_all_xstart, _own_buffers=dump is None)):
            if dump_steps is not None and i in dump_steps:
                dump.append(deepcopy(sample["sample"]))
            if dump_all_xstart:
                dump.append(sample["pred_xstart"])
            final = sample
        if dump is not None:
            return dump
        return final["sample"].clone() if final.get("_live") else final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                                  randomize_class=False, cond_fn_with_grad=False, const_noise=False,
                                  pred_xstart_in_graph=False, stop_timesteps=None, _want_xstart=True,
                                  _own_buffers=False):
        """Generator over the per-step dicts of p_sample (reference :717-794)."""
        yield from self._loop(L.SAMPLER_DDPM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                              device, progress, skip_timesteps, init_image, randomize_class,
                              cond_fn_with_grad or pred_xstart_in_graph, const_noise, stop_timesteps, 0.0,
                              _want_xstart, _own_buffers, pred_xstart_in_graph)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False,
                         pred_xstart_in_graph=False, dump_all_xstart=False, stop_timesteps=None):
        """Full DDIM trajectory (reference :948-1005)."""
        dump = [] if (dump_steps is not None or dump_all_xstart) else None
        if const_noise == True:  # noqa: E712  (reference :978)
            raise NotImplementedError()
        final = None
        for sample in self.ddim_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                model_kwargs=model_kwargs, device=device, progress=progress, eta=eta, skip_timesteps=skip_timesteps,
                init_image=init_image, randomize_class=randomize_class, cond_fn_with_grad=cond_fn_with_grad,
                pred_xstart_in_graph=pred_xstart_in_graph, stop_timesteps=stop_timesteps,
                _want_xstart=dump_all_xstart, _own_buffers=dump is None):
            if dump_all_xstart:
                dump.append(sample["pred_xstart"])
            final = sample
        if dump_all_xstart:
            return dump
        return final["sample"].clone() if final.get("_live") else final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0,
                                     init_image=None, randomize_class=False, cond_fn_with_grad=False,
                                     pred_xstart_in_graph=False, stop_timesteps=None, _want_xstart=True,
                                     _own_buffers=False):
        """Generator over the per-step dicts of ddim_sample (reference :1007-1082)."""
        yield from self._loop(L.SAMPLER_DDIM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                              device, progress, skip_timesteps, init_image, randomize_class,
                              cond_fn_with_grad or pred_xstart_in_graph, False, stop_timesteps, eta, _want_xstart,
                              _own_buffers, pred_xstart_in_graph)

    # the shared driver -----------------------------------------------------------
    def _loop(self, sampler, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress,
              skip_timesteps, init_image, randomize_class, with_grad, const_noise, stop_timesteps, eta, want_xstart,
              own_buffers, pred_xstart_in_graph=False):
        if device is None:
            try:
                device = next(model.parameters()).device
            except Exception:
                device = next(model.model.parameters()).device
        device = th.device(device)
        if device.type != "cuda":
            raise RuntimeError("the mst sampler runs on CUDA devices only (model is on %s)" % device)
        assert isinstance(shape, (tuple, list))
        if noise is not None:
            img = noise
        elif self.noise_fn is not None:
            img = self.noise_fn(-1, tuple(shape), device)  # test hook: draw -1 is x_T (the reference's th.randn(*shape))
        else:
            img = th.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = th.zeros_like(img)
        indices = list(range(self.num_timesteps - skip_timesteps))[::-1]
        if stop_timesteps is not None:
            indices = list(range(stop_timesteps, self.num_timesteps - skip_timesteps))[::-1]
        if init_image is not None:
            my_t = th.ones([shape[0]], device=device, dtype=th.long) * indices[0]
            img = self.q_sample(init_image, my_t, img, model_kwargs=model_kwargs)
        if not indices:
            return

        native, cfgw = self._unwrap(model)
        fused = (native is not None and cond_fn is None and denoised_fn is None and not randomize_class
                 and not with_grad and model_kwargs is not None and 'y' in model_kwargs and native.mst_ready(img))
        if fused:
            yield from self._fused_trajectory(sampler, native, cfgw, img, indices, clip_denoised, model_kwargs,
                                              const_noise, eta, progress, want_xstart, own_buffers)
            return

        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for step_no, i in enumerate(indices):
            t = th.full((shape[0],), i, device=device, dtype=th.long)
            if randomize_class and 'y' in model_kwargs:
                model_kwargs['y'] = th.randint(low=0, high=model.num_classes, size=model_kwargs['y'].shape,
                                               device=model_kwargs['y'].device)
            with th.no_grad():
                if with_grad:  # reference :781 / :1069: the *_with_grad step re-enables autograd around the model
                    out = self._sample_step_with_grad(sampler, model, img, t, clip_denoised, denoised_fn, cond_fn,
                                                      model_kwargs, const_noise=const_noise, eta=eta, step_no=step_no,
                                                      pred_xstart_in_graph=pred_xstart_in_graph)
                else:
                    out = self._sample_step(sampler, model, img, t, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                                            const_noise=const_noise, eta=eta, step_no=step_no)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ fused native trajectory
    def _fused_trajectory(self, sampler, native, cfgw, img, indices, clip_denoised, model_kwargs, const_noise, eta,
                          progress, want_xstart, own_buffers):
        """Whole trajectory on the hand-written path: per step = one denoiser forward (cond+uncond batched)
        + one fused update kernel, replayed as a CUDA graph with the timestep living in device memory.

        The static buffers and the captured graph form a *plan* that is cached on this diffusion object per
        (sampler, shape, conditioning layout, precision ...): a second trajectory of the same kind only copies
        its inputs into the plan's buffers and replays."""
        device = img.device
        y = model_kwargs['y']
        B, T = img.shape[0], img.shape[-1]
        eng = native.mst_engine(device)
        use_cfg = cfgw is not None
        with th.no_grad():
            mask, inp = self._check_inpainting(model_kwargs, img.shape)
            nmask = self._inpainting_mask_for_noise(model_kwargs)
            kmask = mask if mask is not None else (_f32c(nmask) if nmask is not None else None)
            kmask = native.compact_mask(kmask)
            uncond_all = bool(y.get('uncond', False)) and not use_cfg
            # hoisted out of the loop: CLIP/text embedding (the reference re-encodes every step,
            # mdm_forstyledataset.py:326)
            text_emb = None if uncond_all else native.text_embedding(y, device)
            use_philox = self.rng == "philox" and self.noise_fn is None
            use_graph = self.use_cuda_graph and self.noise_fn is None
            n_steps = len(indices)
            assert all(indices[k] - 1 == indices[k + 1] for k in range(n_steps - 1))
            key = (sampler, str(device), tuple(img.shape), id(eng), use_cfg, uncond_all, text_emb is not None,
                   None if kmask is None else tuple(kmask.shape), inp is not None, nmask is not None,
                   bool(clip_denoised), use_philox, bool(const_noise), bool(want_xstart or not own_buffers),
                   float(eta), use_graph)
            plan = self._graphs.get(key)
            if plan is None:
                plan = _TrajectoryPlan(self, sampler, eng, device, tuple(img.shape), use_cfg, uncond_all, text_emb,
                                       kmask, inp, nmask is not None, clip_denoised, use_philox, const_noise,
                                       want_xstart or not own_buffers, eta, use_graph)
                if len(self._graphs) >= 16:  # bound the cache: plans own state-sized buffers
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = plan
            scale = _f32c(y['scale'].to(device)).view(-1) if use_cfg else None
            t_model = th.tensor([self._model_time_index(i) for i in range(self.num_timesteps)], device=device,
                                dtype=th.long)
            plan.load(img, text_emb, scale, kmask, inp, eng.time_embed(t_model), indices[0],
                      self.philox_seed, self.philox_sample_offset)
            self.last_plan_launches = plan.launches_per_step

            if (own_buffers and not progress and self.noise_fn is None and plan.graph is not None
                    and self.trajectory_graph != "step"):
                # nobody looks at the intermediate states: submit the whole trajectory at once
                plan.run_trajectory(n_steps, self.trajectory_graph)
                yield {"sample": plan.x, "pred_xstart": plan.x0, "_live": True}
                return
            it = range(n_steps)
            if progress:
                from tqdm.auto import tqdm
                it = tqdm(it)
            for k in it:
                if self.noise_fn is not None:
                    e = _f32c(self.noise_fn(k, tuple(img.shape), device))
                    if const_noise:
                        e = e[[0]].repeat(B, 1, 1, 1)
                    plan.eps.copy_(e)
                plan.step()
                if own_buffers:
                    # the caller (p_sample_loop / ddim_sample_loop) only keeps the last dict and clones it
                    yield {"sample": plan.x, "pred_xstart": plan.x0, "_live": True}
                else:
                    yield {"sample": plan.x.clone(), "pred_xstart": plan.x0.clone()}


class _TrajectoryPlan:
    """Static device buffers + the captured one-step CUDA graph of a fused trajectory."""

    def __init__(self, diffusion, sampler, eng, device, shape, use_cfg, uncond_all, text_emb, kmask, inp, mask_noise,
                 clip_denoised, use_philox, const_noise, want_xstart, eta, use_graph):
        self.eng, self.device, self.shape = eng, device, shape
        B, T = shape[0], shape[-1]
        f32 = dict(dtype=th.float32, device=device)
        self.x = th.empty(shape, **f32)
        self.out_c = th.empty(shape, **f32)
        self.out_u = th.empty(shape, **f32) if use_cfg else None
        self.x0 = th.empty(shape, **f32) if want_xstart else None
        self.eps = None if use_philox else th.empty(shape, **f32)
        self.text_emb = th.empty_like(text_emb) if text_emb is not None else None
        self.scale = th.empty(B, **f32) if use_cfg else None
        self.mask = th.empty_like(kmask) if kmask is not None else None
        self.inp = th.empty(shape, **f32) if inp is not None else None
        self.temb = th.empty(diffusion.num_timesteps, eng.d_model, **f32)
        self.t_dev = th.zeros(1, dtype=th.int32, device=device)
        self.counter = th.zeros(1, dtype=th.int32, device=device)
        self.ws = th.empty(eng.workspace_bytes(B * (2 if use_cfg else 1), T), dtype=th.uint8, device=device)
        self.seed = th.zeros(2, dtype=th.int64, device=device)  # reserved: philox seed/offset are launch arguments
        tabs = diffusion.device_tables(device, eta)
        if sampler == L.SAMPLER_DDPM:
            coefs = dict(coef1=tabs["c1"], coef2=tabs["c2"], sigma=tabs["sigma"])
        else:
            coefs = dict(coef1=tabs["ddim_c1"], coef2=tabs["ddim_c2"], sigma=tabs["ddim_sigma"],
                         recip=tabs["recip"], recipm1=tabs["recipm1"])
        self._args = dict(sampler=sampler, use_cfg=use_cfg, uncond_all=uncond_all, mask_noise=mask_noise,
                          clip_denoised=clip_denoised, use_philox=use_philox, const_noise=const_noise, coefs=coefs)
        self.use_graph = use_graph
        self.graph = None
        self.graph_key = None
        self.traj_graphs = {}  # number of steps -> graph of that many consecutive steps
        self.last_trajectory_launches = 0
        self.launches_per_step = 0
        self.draw_noise = self.eps is not None and diffusion.noise_fn is None

    def _one_step(self, seed, offset):
        a = self._args
        self.eng.forward(self.x, self.temb, self.text_emb, cfg=a["use_cfg"], uncond=a["uncond_all"], out_cond=self.out_c,
                         out_uncond=self.out_u, temb_row_dev=self.t_dev, workspace=self.ws)
        if self.draw_noise:
            self.eps.normal_()
            if a["const_noise"]:
                self.eps.copy_(self.eps[[0]].expand_as(self.eps).clone())
        K.update_step(sampler=a["sampler"], out_cond=self.out_c, out_uncond=self.out_u, cfg_scale=self.scale,
                      x_t=self.x, x_prev=self.x, pred_xstart=self.x0, mask=self.mask, x_inpaint=self.inp,
                      mask_noise=a["mask_noise"], clip_denoised=a["clip_denoised"], t_scalar_dev=self.t_dev,
                      advance_t=True, block_counter=self.counter,
                      noise_kind=L.NOISE_PHILOX if a["use_philox"] else L.NOISE_TENSOR, noise=self.eps,
                      const_noise=a["const_noise"] and a["use_philox"], philox_seed=seed, philox_sample_offset=offset,
                      **a["coefs"])

    def load(self, img, text_emb, scale, kmask, inp, temb, t_start, seed, offset):
        """Copy this trajectory's inputs into the static buffers; (re)capture the step graph when needed."""
        self.x.copy_(_f32c(img))
        if self.text_emb is not None:
            self.text_emb.copy_(text_emb)
        if self.scale is not None:
            self.scale.copy_(scale)
        if self.mask is not None:
            self.mask.copy_(kmask)
        if self.inp is not None:
            self.inp.copy_(inp)
        self.temb.copy_(temb)
        self.seed_off = (int(seed), int(offset))
        # the captured launches bake in the Philox seed AND the engine's weight pointers (fp32 views of the parameters,
        # packed copies): a reload of the weights - parameters moved, replaced or re-packed after an optimizer step -
        # bumps the engine's generation and the step is captured again
        want_key = (self.seed_off, self.eng.weights_generation)
        if self.use_graph and (self.graph is None or self.graph_key != want_key):
            self._capture()
        self.t_dev.fill_(int(t_start))
        self.counter.zero_()

    def _capture(self):
        device = self.device
        snap = self.x.clone()
        self.t_dev.fill_(1)
        self.counter.zero_()
        rng_state = th.cuda.get_rng_state(device)
        s = th.cuda.Stream(device=device)
        s.wait_stream(th.cuda.current_stream(device))
        with th.cuda.stream(s):  # warm-up outside capture: lazy module loads, allocator
            self._one_step(*self.seed_off)
        th.cuda.current_stream(device).wait_stream(s)
        th.cuda.synchronize(device)
        th.cuda.set_rng_state(rng_state, device)
        n0 = K.launch_count()
        graph = th.cuda.CUDAGraph()
        with th.cuda.graph(graph):
            self._one_step(*self.seed_off)
        self.launches_per_step = K.launch_count() - n0
        self.graph, self.graph_key = graph, (self.seed_off, self.eng.weights_generation)
        self.traj_graphs = {}
        self.x.copy_(snap)

    def run_trajectory(self, n_steps, mode="full"):
        """Submit n_steps denoise steps: one graph launch when the whole trajectory fits one graph."""
        max_nodes = int(os.environ.get("MST_TRAJ_GRAPH_MAX_NODES", "65536"))
        per = max(1, self.launches_per_step)
        chunk = n_steps if mode == "full" else max(1, int(mode))
        chunk = max(1, min(chunk, n_steps, max_nodes // per))
        done = 0
        while done < n_steps:
            k = min(chunk, n_steps - done)
            g = self.traj_graphs.get(k)
            if g is None:
                g = self.traj_graphs[k] = self._capture_steps(k)
            g.replay()
            done += k
        K.count_graph_replay(n_steps)
        self.last_trajectory_launches = -(-n_steps // chunk)

    def _capture_steps(self, k):
        """k consecutive denoise steps as one CUDA graph (the buffers, the device-side timestep and the Philox key are
        the plan's; nothing in a step depends on the host)."""
        snap, t_snap, c_snap = self.x.clone(), self.t_dev.clone(), self.counter.clone()
        graph = th.cuda.CUDAGraph()
        with th.cuda.graph(graph):
            for _ in range(k):
                self._one_step(*self.seed_off)
        self.x.copy_(snap)
        self.t_dev.copy_(t_snap)
        self.counter.copy_(c_snap)
        return graph

    def step(self):
        if self.graph is not None:
            self.graph.replay()
            K.count_graph_replay()
        else:
            n0 = K.launch_count()
            self._one_step(*self.seed_off)
            self.launches_per_step = K.launch_count() - n0
