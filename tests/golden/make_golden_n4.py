#!/usr/bin/env python
"""Golden for scope row N4: the REAL reference ``CompMDMGeneratedDataset`` (data_loaders/humanml/motion_loaders/
comp_v6_model_dataset.py:146-233) run on CPU over a small seeded data loader with a CFG model, guidance scale 2.5,
a 4-step respaced sampler and the injected noise tape; also the ``dump_steps`` behaviour of ``p_sample_loop``
(gaussian_diffusion.py:644-716) and the literal demo call sequence with the reference's own ``collate``.

    python tests/golden/make_golden_n4.py        # build container only (needs /root/reference)
"""
import importlib
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import n4_inputs as NI  # noqa: E402
from oracle.weights import NoiseTape, mdm_state_dict, text_features  # noqa: E402


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return mock.MagicMock(name=f"{self.__name__}.{name}")


def stub_plotting_modules():
    """comp_v6_model_dataset.py imports the evaluators' trainers, which import matplotlib & co at module level"""
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.mplot3d.art3d", "mpl_toolkits.mplot3d.axes3d", "spacy", "blobfile", "PIL", "PIL.Image"]:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = _Stub(name)
                m.__path__ = []
                sys.modules[name] = m


def main():
    torch.set_grad_enabled(False)
    MG.install_shims()
    stub_plotting_modules()
    from utils import model_util as ref_mu
    from diffusion.inpainting_gaussian_diffusion import InpaintingGaussianDiffusion
    from diffusion.respace import SpacedDiffusion
    from model.cfg_sampler import ClassifierFreeSampleModel
    from model.mdm_forstyledataset import MDM
    from data_loaders.humanml.motion_loaders import comp_v6_model_dataset as ref_n4
    from data_loaders import stylexia_posrot_utils as ref_masks
    from data_loaders.tensors import collate as ref_collate
    import torch as th

    args = MG.Args()
    state = mdm_state_dict(n_feats=181, seed=0)
    ref_model = MDM(**ref_mu.get_transfer_args(args))
    ref_model.load_state_dict(state, strict=False)
    ref_model.train(False)
    out = {}

    def with_tape(tape, fn):
        orig_randn, orig_like = th.randn, th.randn_like
        th.randn = lambda *s, **k: tape.draw(s[0] if len(s) == 1 and isinstance(s[0], (tuple, list, torch.Size)) else s)
        th.randn_like = lambda a, **k: tape.draw(a.shape)
        try:
            return fn()
        finally:
            th.randn, th.randn_like = orig_randn, orig_like

    # ---- 1. CompMDMGeneratedDataset
    d4 = ref_mu.create_gaussian_diffusion(args, SpacedDiffusion, timestep_respacing=NI.SPEC)
    np.random.seed(123)
    data = NI.loader()  # built BEFORE torch.randn is patched: the fake dataset draws its motions with torch.randn too
    ds = with_tape(NoiseTape(41), lambda: ref_n4.CompMDMGeneratedDataset(
        ClassifierFreeSampleModel(ref_model), d4, data, NI.MM_SAMPLES, NI.MM_REPEATS, NI.T, None, scale=NI.SCALE))
    out["n4_motion"] = np.stack([d["motion"] for d in ds.generated_motion])               # [6, T, F]
    out["n4_length"] = np.array([int(d["length"]) for d in ds.generated_motion])
    out["n4_cap_len"] = np.array([d["cap_len"] for d in ds.generated_motion])
    out["n4_mm_idx"] = np.array([i for i, d in enumerate(ds.generated_motion)
                                 if any(d["caption"] == m["caption"] for m in ds.mm_generated_motion)])
    out["n4_mm_motions"] = np.stack([np.stack([m["motion"] for m in d["mm_motions"]]) for d in ds.mm_generated_motion])
    print("N4:", out["n4_motion"].shape, "mm", out["n4_mm_motions"].shape, "mm items", out["n4_mm_idx"].tolist())

    # ---- 2. dump_steps (p_sample_loop returns deep copies of the sample at the listed loop indices)
    B, T = 2, 24
    shape = (B, 181, 1, T)
    g = torch.Generator().manual_seed(21)
    x_inp = torch.randn(shape, generator=g)
    mask = torch.from_numpy(ref_masks.get_inpainting_mask("root_horizontal", shape)).float()
    texts = ["a person walks like an old man", "a person jumps happily"]
    yk = {"y": {"text": texts, "mask": torch.ones(B, 1, 1, T), "lengths": torch.tensor([T] * B), "scale": torch.tensor([2.5, 1.5]),
                "inpainted_motion": x_inp, "inpainting_mask": mask}}
    d8 = ref_mu.create_gaussian_diffusion(args, InpaintingGaussianDiffusion, timestep_respacing="8")
    dump = with_tape(NoiseTape(5), lambda: d8.p_sample_loop(ClassifierFreeSampleModel(ref_model), shape, clip_denoised=False,
                                                            model_kwargs=yk, dump_steps=[0, 3, 7]))
    assert isinstance(dump, list) and len(dump) == 3
    out["dump_steps_037"] = torch.stack(dump).numpy()

    # ---- 3. the demo's call sequence with the reference's collate (sample/demo_style_transfer.py:196-258)
    T = 76
    g = torch.Generator().manual_seed(31)
    content, style = torch.randn(181, 1, T, generator=g), torch.randn(181, 1, T, generator=g)
    collate_args = [{"inp": content, "tokens": None, "lengths": T, "text": "a person walks"}]
    _, model_kwargs = ref_collate(collate_args)
    shape = (1, 181, 1, T)
    model_kwargs["y"]["inpainted_motion"] = style[None]
    model_kwargs["y"]["inpainting_mask"] = torch.tensor(ref_masks.get_inpainting_mask("root_horizontal", shape)).float()
    model_kwargs["y"]["scale"] = torch.ones(1) * 2.5
    d20 = ref_mu.create_gaussian_diffusion(args, InpaintingGaussianDiffusion, timestep_respacing="ddim20")
    res = with_tape(NoiseTape(9), lambda: d20.ddim_sample_loop(ref_model, shape, clip_denoised=False, model_kwargs=model_kwargs,
                                                               skip_timesteps=14, init_image=content[None], progress=False,
                                                               dump_all_xstart=True))
    out["demo_collate_xstart_last"] = res[-1].numpy()
    out["demo_collate_mask"] = model_kwargs["y"]["mask"].numpy()
    out["demo_collate_lengths"] = model_kwargs["y"]["lengths"].numpy()

    path = os.path.join(HERE, "n4_dump_demo.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
