"""Inpainting override of the diffusion process (reference
``diffusion/inpainting_gaussian_diffusion.py``).

The reference subclasses ``SpacedDiffusion`` and re-implements ``q_sample``,
``p_sample`` and ``ddim_sample`` only to multiply the freshly drawn noise by
``1 - y['inpainting_mask']`` (``:18``, ``:54``, ``:162``).  In this build the
noise masking is a flag of the fused update kernel, so the subclass just says
which mask applies; the blend of the model output with ``y['inpainted_motion']``
itself lives in the base class exactly as in the reference
(``gaussian_diffusion.py:341-349``).
"""
from .respace import SpacedDiffusion


class InpaintingGaussianDiffusion(SpacedDiffusion):
    def _inpainting_mask_for_noise(self, model_kwargs):
        # the reference indexes model_kwargs['y']['inpainting_mask'] unconditionally
        # (inpainting_gaussian_diffusion.py:18, :54): a missing key is a KeyError / TypeError there too
        return model_kwargs['y']['inpainting_mask']
