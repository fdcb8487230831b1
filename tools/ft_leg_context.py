#!/usr/bin/env python
"""Why is the finetune leg slower inside bench.py than alone?  Runs it (a) first, (b) after one sampler trajectory of the
headline shape in the same process, (c) after a garbage collection + cache flush.   python tools/ft_leg_context.py"""
import gc
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.chdir(REPO)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
a = types.SimpleNamespace()
print("alone:", bench.finetune_leg(a, dev, 0, 1, None)["value"], flush=True)

from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask  # noqa: E402
from mst_b200.model.cfg_sampler import ClassifierFreeSampleModel  # noqa: E402
from mst_b200.model.mdm_forstyledataset import MDM  # noqa: E402
from mst_b200.utils import model_util as mu  # noqa: E402

args = bench.Args()
torch.manual_seed(0)
model = MDM(load_clip=False, **mu.get_transfer_args(args)).to(dev).eval()
cfg = ClassifierFreeSampleModel(model)
diffusion = mu.create_gaussian_diffusion(args, mu.InpaintingGaussianDiffusion, timestep_respacing="")
diffusion.rng, diffusion.philox_seed = "philox", 3
B, T, F = 64, 196, 181
y = {"y": {"text": ["a"] * B, "text_feat": torch.randn(B, 512, device=dev), "scale": torch.full((B,), 2.5, device=dev),
           "inpainted_motion": torch.randn(B, F, 1, T, device=dev),
           "inpainting_mask": torch.from_numpy(get_inpainting_mask("root_horizontal", (B, F, 1, T))).float().to(dev),
           "mask": torch.ones(B, 1, 1, T, device=dev), "lengths": torch.full((B,), T, device=dev)}}
with torch.no_grad():
    for _ in range(2):
        diffusion.p_sample_loop(cfg, (B, F, 1, T), clip_denoised=False, model_kwargs=y, progress=False)
torch.cuda.synchronize()
print("after two 1000-step trajectories:", bench.finetune_leg(a, dev, 0, 1, None)["value"], flush=True)
with torch.no_grad():
    diffusion.trajectory_graph = "step"
    diffusion.p_sample_loop(cfg, (B, F, 1, T), clip_denoised=False, model_kwargs=y, progress=False)
torch.cuda.synchronize()
print("after a per-step-replay trajectory:", bench.finetune_leg(a, dev, 0, 1, None)["value"], flush=True)
del diffusion, cfg, model, y
gc.collect()
torch.cuda.empty_cache()
print("after dropping the sampler + gc:", bench.finetune_leg(a, dev, 0, 1, None)["value"], flush=True)
