"""Deterministic random-init weights shared by the reference run, the oracle and the CUDA engine
(oracle; test infrastructure only).  Each tensor gets its own torch.Generator seeded from
(seed, crc32(name)), so the set can be regenerated anywhere without shipping 70 MB of fixtures."""
import math
import zlib

import numpy as np
import torch

_torch_randn = torch.randn  # bound early: the golden script monkeypatches torch.randn while the reference runs


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def _uniform(shape, bound, seed, name):
    return (torch.rand(shape, generator=_gen(seed, name), dtype=torch.float32) * 2 - 1) * bound


def positional_table(d_model, max_len=5000):
    """reference model/mdm_forstyledataset.py:392-397"""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()  # [max_len, 1, d]


def encoder_layer_weights(prefix, d, ff, seed):
    w = {}
    w[prefix + "self_attn.in_proj_weight"] = _uniform((3 * d, d), math.sqrt(6.0 / (4 * d)), seed, prefix + "qkv_w")
    w[prefix + "self_attn.in_proj_bias"] = _uniform((3 * d,), 0.02, seed, prefix + "qkv_b")
    w[prefix + "self_attn.out_proj.weight"] = _uniform((d, d), 1 / math.sqrt(d), seed, prefix + "o_w")
    w[prefix + "self_attn.out_proj.bias"] = _uniform((d,), 0.02, seed, prefix + "o_b")
    w[prefix + "linear1.weight"] = _uniform((ff, d), 1 / math.sqrt(d), seed, prefix + "w1")
    w[prefix + "linear1.bias"] = _uniform((ff,), 1 / math.sqrt(d), seed, prefix + "b1")
    w[prefix + "linear2.weight"] = _uniform((d, ff), 1 / math.sqrt(ff), seed, prefix + "w2")
    w[prefix + "linear2.bias"] = _uniform((d,), 1 / math.sqrt(ff), seed, prefix + "b2")
    w[prefix + "norm1.weight"] = 1.0 + _uniform((d,), 0.1, seed, prefix + "ln1_g")
    w[prefix + "norm1.bias"] = _uniform((d,), 0.1, seed, prefix + "ln1_b")
    w[prefix + "norm2.weight"] = 1.0 + _uniform((d,), 0.1, seed, prefix + "ln2_g")
    w[prefix + "norm2.bias"] = _uniform((d,), 0.1, seed, prefix + "ln2_b")
    return w


def mdm_state_dict(n_feats=181, d=512, ff=1024, n_layers=8, clip_dim=512, seed=0, pe_len=5000):
    """state_dict of the reference's MDM (minus clip_model.*) with reproducible values."""
    w = {}
    w["input_process.poseEmbedding.weight"] = _uniform((d, n_feats), 1 / math.sqrt(n_feats), seed, "in_w")
    w["input_process.poseEmbedding.bias"] = _uniform((d,), 1 / math.sqrt(n_feats), seed, "in_b")
    pe = positional_table(d, pe_len)
    w["sequence_pos_encoder.pe"] = pe
    w["embed_timestep.sequence_pos_encoder.pe"] = pe.clone()
    for i in range(n_layers):
        w.update(encoder_layer_weights(f"seqTransEncoder.layers.{i}.", d, ff, seed))
    w["embed_timestep.time_embed.0.weight"] = _uniform((d, d), 1 / math.sqrt(d), seed, "t_w1")
    w["embed_timestep.time_embed.0.bias"] = _uniform((d,), 1 / math.sqrt(d), seed, "t_b1")
    w["embed_timestep.time_embed.2.weight"] = _uniform((d, d), 1 / math.sqrt(d), seed, "t_w2")
    w["embed_timestep.time_embed.2.bias"] = _uniform((d,), 1 / math.sqrt(d), seed, "t_b2")
    w["embed_text.weight"] = _uniform((d, clip_dim), 1 / math.sqrt(clip_dim), seed, "txt_w")
    w["embed_text.bias"] = _uniform((d,), 1 / math.sqrt(clip_dim), seed, "txt_b")
    w["output_process.poseFinal.weight"] = _uniform((n_feats, d), 1 / math.sqrt(d), seed, "out_w")
    w["output_process.poseFinal.bias"] = _uniform((n_feats,), 1 / math.sqrt(d), seed, "out_b")
    return w


def checksum(state):
    """order-independent fingerprint of a state dict (float64 sum of |w| per tensor, then summed)"""
    return float(sum(v.double().abs().sum().item() for k, v in sorted(state.items())))


class NoiseTape:
    """Reproducible sequence of N(0,1) tensors: draw k of a tape seeded s is torch.randn(shape,
    generator=Generator(s, k)).  The golden script injects it into the reference in place of
    torch.randn / randn_like; the oracle and the CUDA path read the same tape."""

    def __init__(self, seed):
        self.seed, self.k = int(seed), 0

    def draw(self, shape):
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed * 7919 + self.k)
        self.k += 1
        return _torch_randn(tuple(shape), generator=g, dtype=torch.float32)


def text_features(texts, clip_dim=512):
    """Stand-in for the (third-party, absent) CLIP text tower: a deterministic [B, clip_dim] feature per
    caption, seeded by crc32 of the string.  The hot path treats CLIP features as an input."""
    rows = []
    for s in texts:
        g = torch.Generator(device="cpu")
        g.manual_seed(zlib.crc32(s.encode()))
        rows.append(_torch_randn(clip_dim, generator=g, dtype=torch.float32))
    return torch.stack(rows)
