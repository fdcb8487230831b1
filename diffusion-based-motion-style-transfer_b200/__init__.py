"""B200-native drop-in for the sampling hot path of hlcdyy/diffusion-based-motion-style-transfer.

Import name: ``mst_b200`` (the directory name carries the reference's hyphenated name and is not a
valid Python identifier; ``mst_b200/__init__.py`` at the repo root aliases it).

Sub-packages mirror the reference's module paths so that its scripts can bind to them unchanged:

    mst_b200.diffusion.gaussian_diffusion / .respace / .inpainting_gaussian_diffusion
    mst_b200.diffusion.resample / .fp16_util,  mst_b200.train.training_loop   (finetune step)
    mst_b200.model.mdm_forstyledataset / .cfg_sampler
    mst_b200.utils.model_util
    mst_b200.data_loaders.{stylexia_posrot,bandai_posrot,humanml}_utils / .tensors

``install(reference_root)`` registers those modules under the reference's own names
(``diffusion.respace`` ...) so ``sample/demo_style_transfer.py`` picks them up - see INTEGRATION.md.
"""
import importlib
import os
import sys

__version__ = "0.1.0"

_OVERLAY = {
    "diffusion.gaussian_diffusion": "diffusion.gaussian_diffusion",
    "diffusion.respace": "diffusion.respace",
    "diffusion.inpainting_gaussian_diffusion": "diffusion.inpainting_gaussian_diffusion",
    "diffusion.resample": "diffusion.resample",
    "diffusion.fp16_util": "diffusion.fp16_util",
    "train.training_loop": "train.training_loop",
    "model.cfg_sampler": "model.cfg_sampler",
    "model.mdm_forstyledataset": "model.mdm_forstyledataset",
    "utils.model_util": "utils.model_util",
    "data_loaders.stylexia_posrot_utils": "data_loaders.stylexia_posrot_utils",
    "data_loaders.bandai_posrot_utils": "data_loaders.bandai_posrot_utils",
    "data_loaders.humanml_utils": "data_loaders.humanml_utils",
}


def install(reference_root=None):
    """Make the reference's import statements resolve to this package for the hot-path modules.

    After ``install()``, ``from diffusion.respace import SpacedDiffusion`` (and the other modules in
    ``_OVERLAY``) return the B200-native implementations; every other reference module
    (``utils.parser_util``, ``data_loaders.get_data`` ...) keeps coming from ``reference_root``.
    """
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    for ref_name, ours in _OVERLAY.items():
        mod = importlib.import_module(f"{__name__}.{ours}")
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition(".")
        try:
            pkg = importlib.import_module(parent)
            setattr(pkg, leaf, mod)
        except Exception:
            pass  # parent package not importable (no reference on the path): the sys.modules entry suffices
    return sorted(_OVERLAY)


def lib_path():
    from . import _lib
    return _lib.LIB_PATH
