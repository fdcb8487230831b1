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
        self.model = model  # the actual denoiser
        assert self.model.cond_mask_prob > 0, \
            'Cannot run a guided diffusion on a model that has not been trained with no conditions'
        # pointers to the inner model (reference cfg_sampler.py:17-25)
        try:
            self.rot2xyz = self.model.rot2xyz
        except Exception:
            self.rot2xyz = None
        self.translation = self.model.translation
        self.njoints = self.model.njoints
        self.nfeats = self.model.nfeats
        self.data_rep = self.model.data_rep
        self.cond_mode = self.model.cond_mode

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
