"""Deterministic inputs / weights shared by tests/golden/make_golden_finetune.py (which runs the real reference in the
build container) and tests/test_gpu_finetune.py / tests/test_oracle_finetune.py (which regenerate them anywhere)."""
import numpy as np
import torch

from oracle.weights import encoder_layer_weights

ENC_SEED, MENC_SEED = 7, 5

def style_encoder_state(seed=ENC_SEED, d=512, ff=1024, n_layers=8):
    w = {}
    for i in range(n_layers):
        w.update(encoder_layer_weights(f"seqTransEncoder.layers.{i}.", d, ff, seed))
    return w


def motion_encoder_state(seed=MENC_SEED, d=512, ff=1024, n_layers=8):
    w = style_encoder_state(seed, d, ff, n_layers)
    g = torch.Generator().manual_seed(seed * 31 + 1)
    w["muQuery"] = torch.randn(1, d, generator=g)
    w["sigmaQuery"] = torch.randn(1, d, generator=g)
    return w


def digest(t):
    t = t.detach().double().flatten().cpu()
    return np.array([t.norm().item(), t.sum().item()] + t[:14].tolist(), dtype=np.float64)


def cases():
    """(name, kwargs) of the finetune-loss configurations that are pinned."""
    return [
        ("ddim_sg0", dict(respacing="ddim20", use_ddim=1, skip_steps=700, semantic_guidance=0)),
        ("ddim_sg1", dict(respacing="ddim20", use_ddim=1, skip_steps=700, semantic_guidance=1)),
        ("ddpm_sg0", dict(respacing="10", use_ddim=0, skip_steps=7, semantic_guidance=0)),
    ]


def make_inputs(F=181, T=20, B=3):
    g = torch.Generator().manual_seed(41)
    x_start = torch.randn(B, F, 1, T, generator=g)
    content = torch.randn(1, F, 1, T, generator=g)
    style = torch.randn(1, F, 1, T, generator=g)
    noise_t2m = torch.rand(B, F, 1, T, generator=g)
    lengths = [T, 14, 9][:B]
    frame_mask_t2m = torch.arange(T)[None, :] < torch.tensor(lengths)[:, None]
    texts_t2m = ["a person walks angrily", "a person jumps old", "a person runs proud"][:B]
    texts_style = ["a person walks neutral"]
    return dict(x_start=x_start, content=content, style=style, noise_t2m=noise_t2m, lengths=lengths,
                frame_mask_t2m=frame_mask_t2m, texts_t2m=texts_t2m, texts_style=texts_style,
                t=torch.tensor([3, 1, 5][:B]))


