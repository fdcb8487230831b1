"""Classifier-free guidance wrapper (reference ``model/cfg_sampler.py``).

Same constructor contract and attributes as the reference's
``ClassifierFreeSampleModel``.  The reference runs the model twice and
deep-copies ``y`` on every call (``cfg_sampler.py:39-43``); here the
conditional and unconditional passes are ONE batched launch sequence of the
native denoiser (2B sequences), and inside a sampling loop the final lerp
``out_u + scale * (out_c - out_u)`` is folded into the fused update kernel.
"""
import torch
import torch.nn as nn

from .. import engine as K


class ClassifierFreeSampleModel(nn.Module):

    def __init__(self, model):
        super().__init__()
        if not model.cond_mask_prob > 0:  # the reference asserts the same (cfg_sampler.py:12-13)
            raise AssertionError('Cannot run a guided diffusion on a model that has not been trained with no conditions')
        self.model = model
        # what callers read off the wrapper instead of the wrapped denoiser (reference cfg_sampler.py:17-25)
        self.rot2xyz = getattr(model, "rot2xyz", None)
        for name in ("translation", "njoints", "nfeats", "data_rep", "cond_mode"):
            setattr(self, name, getattr(model, name))

    def forward(self, x, timesteps, y=None):
        cond_mode = self.model.cond_mode
        assert cond_mode in ['text', 'action']
        from .mdm_forstyledataset import NativeDenoiser
        if isinstance(self.model, NativeDenoiser) and self.model.mst_ready(x):
            out_c, out_u = self.model.forward_cfg(x, timesteps, y)
        else:
            # foreign denoiser: two calls, but still no deepcopy of the state-sized tensors in y
            y_uncond = dict(y)
            y_uncond['uncond'] = True
            out_c = self.model(x, timesteps, y).float().contiguous()
            out_u = self.model(x, timesteps, y_uncond).float().contiguous()
        scale = y['scale'].to(x.device).float().contiguous().view(-1)
        return K.cfg_combine(out_c, out_u, scale)
