"""torch-CPU fp32 restatement of the sampling loops (oracle; test infrastructure only).

Follows the reference's operation order so that fp32 rounding matches:
* _extract_into_tensor       diffusion/gaussian_diffusion.py:1605-1618 (float64 table -> gather -> .float())
* q_sample (inpainting)      diffusion/inpainting_gaussian_diffusion.py:6-23
* p_mean_variance            diffusion/gaussian_diffusion.py:311-424
* p_sample (inpainting)      diffusion/inpainting_gaussian_diffusion.py:25-64
* ddim_sample (inpainting)   diffusion/inpainting_gaussian_diffusion.py:125-174
* p_sample_loop_progressive  diffusion/gaussian_diffusion.py:717-794 (ddim: :1007-1082)
"""
import numpy as np
import torch


def extract(arr, t, ndim=4):
    r = torch.from_numpy(np.asarray(arr, dtype=np.float64))[t].float()
    return r.view(-1, *([1] * (ndim - 1)))


def q_sample(sch, x0, t, noise, mask=None):
    if mask is not None:
        noise = noise * (1.0 - mask)
    return extract(sch.sqrt_abar, t) * x0 + extract(sch.sqrt_1m_abar, t) * noise


def pred_xstart(model_out, mask=None, x_inp=None, clip=False):
    x0 = model_out
    if mask is not None and x_inp is not None:
        x0 = (x0 * (1 - mask)) + (x_inp * mask)
    if clip:
        x0 = x0.clamp(-1, 1)
    return x0


def p_sample(sch, model_out, x, t, noise, mask=None, x_inp=None, clip=False, mask_noise=True):
    x0 = pred_xstart(model_out, mask, x_inp, clip)
    mean = extract(sch.coef1, t) * x0 + extract(sch.coef2, t) * x
    logvar = extract(sch.logvar, t)
    if mask is not None and mask_noise:
        noise = noise * (1.0 - mask)
    nonzero = (t != 0).float().view(-1, 1, 1, 1)
    return mean + nonzero * torch.exp(0.5 * logvar) * noise, x0


def ddim_sample(sch, model_out, x, t, noise, mask=None, x_inp=None, clip=False, eta=0.0, mask_noise=True):
    x0 = pred_xstart(model_out, mask, x_inp, clip)
    eps = (extract(sch.sqrt_recip_abar, t) * x - x0) / extract(sch.sqrt_recipm1_abar, t)
    ab, abp = extract(sch.abar, t), extract(sch.abar_prev, t)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    if mask is not None and mask_noise:
        noise = noise * (1.0 - mask)
    mean = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
    nonzero = (t != 0).float().view(-1, 1, 1, 1)
    return mean + nonzero * sigma * noise, x0


def sample_loop(sch, model_fn, shape, tape, *, ddim=False, mask=None, x_inp=None, clip=False, skip_timesteps=0,
                init_image=None, stop_timesteps=None, eta=0.0, max_steps=None, mask_noise=True):
    """model_fn(x, t_original) -> model output; ``tape.draw(shape)`` supplies every N(0,1) draw in the order the
    reference consumes them (initial image first, then one per step).  Returns (final sample, [pred_xstart per step])."""
    B = shape[0]
    img = tape.draw(shape)
    if skip_timesteps and init_image is None:
        init_image = torch.zeros_like(img)
    lo = 0 if stop_timesteps is None else stop_timesteps
    indices = list(range(lo, sch.N - skip_timesteps))[::-1]
    if init_image is not None:
        t0 = torch.full((B,), indices[0], dtype=torch.long)
        img = q_sample(sch, init_image, t0, img, mask if mask_noise else None)
    xstarts = []
    tmap = torch.tensor(sch.timestep_map, dtype=torch.long)
    for n, i in enumerate(indices):
        if max_steps is not None and n >= max_steps:
            break
        t = torch.full((B,), i, dtype=torch.long)
        out = model_fn(img, tmap[t])
        noise = tape.draw(shape)
        step = ddim_sample if ddim else p_sample
        kw = dict(eta=eta) if ddim else {}
        img, x0 = step(sch, out, img, t, noise, mask, x_inp, clip, mask_noise=mask_noise, **kw)
        xstarts.append(x0)
    return img, xstarts
