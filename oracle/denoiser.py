"""torch-CPU fp32 restatement of the denoiser forward (oracle; test infrastructure only).

Plain tensor algebra - no nn.TransformerEncoder, no fused attention - so it is an independent
statement of what the reference computes through torch's library layers:

* MDM.forward                      reference model/mdm_forstyledataset.py:315-364
* StyleDiffusion.forward           reference model/mdm_forstyledataset.py:602-625 (same math, other encoder weights)
* nn.TransformerEncoderLayer       torch semantics the reference relies on at :231-238: batch_first=False,
                                   norm_first=False (post-norm), activation=gelu (exact erf), LayerNorm eps=1e-5,
                                   softmax(QK^T/sqrt(dh))V, no mask
* ClassifierFreeSampleModel        reference model/cfg_sampler.py:36-43
"""
import math

import torch


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def encoder_layer(x, w, pre, n_heads):
    """x: [S, B, d] -> [S, B, d]; w: state dict, pre: 'seqTransEncoder.layers.i.'"""
    S, B, d = x.shape
    dh = d // n_heads
    qkv = x @ w[pre + "self_attn.in_proj_weight"].T + w[pre + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)

    def heads(t):  # [S,B,d] -> [B,H,S,dh]
        return t.reshape(S, B, n_heads, dh).permute(1, 2, 0, 3)

    q, k, v = heads(q), heads(k), heads(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(S, B, d)
    sa = o @ w[pre + "self_attn.out_proj.weight"].T + w[pre + "self_attn.out_proj.bias"]
    x = layer_norm(x + sa, w[pre + "norm1.weight"], w[pre + "norm1.bias"])
    h = gelu(x @ w[pre + "linear1.weight"].T + w[pre + "linear1.bias"])
    ff = h @ w[pre + "linear2.weight"].T + w[pre + "linear2.bias"]
    return layer_norm(x + ff, w[pre + "norm2.weight"], w[pre + "norm2.bias"])


def time_embedding(w, t):
    """TimestepEmbedder.forward (:421): time_embed(pe[t]) -> [B, d]"""
    pe = w["sequence_pos_encoder.pe"][:, 0, :]
    h = pe[t] @ w["embed_timestep.time_embed.0.weight"].T + w["embed_timestep.time_embed.0.bias"]
    h = h * torch.sigmoid(h)
    return h @ w["embed_timestep.time_embed.2.weight"].T + w["embed_timestep.time_embed.2.bias"]


def mdm_forward(w, x, t, text_feat, uncond=False, n_heads=4, enc_prefix="seqTransEncoder.layers.", enc_w=None):
    """x [B,F,1,T] fp32, t int64 [B], text_feat [B,clip] or None -> [B,F,1,T].
    ``enc_w`` lets StyleDiffusion use its own encoder weights with the frozen MDM's projections."""
    B, F, _, T = x.shape
    emb = time_embedding(w, t)                                         # [B,d]
    if "embed_text.weight" in w and text_feat is not None or uncond and "embed_text.weight" in w:
        feat = torch.zeros(B, w["embed_text.weight"].shape[1]) if uncond else text_feat   # mask_cond(force_mask) :288-291
        emb = emb + feat @ w["embed_text.weight"].T + w["embed_text.bias"]               # :327
    xs = x.permute(3, 0, 1, 2).reshape(T, B, F)                         # :437
    xs = xs @ w["input_process.poseEmbedding.weight"].T + w["input_process.poseEmbedding.bias"]
    seq = torch.cat([emb[None], xs], dim=0)                             # :344
    seq = seq + w["sequence_pos_encoder.pe"][:T + 1]                    # :403 (dropout is identity in eval)
    ew = enc_w if enc_w is not None else w
    n_layers = 1 + max(int(k[len(enc_prefix):].split(".")[0]) for k in ew if k.startswith(enc_prefix))
    for i in range(n_layers):
        seq = encoder_layer(seq, ew, f"{enc_prefix}{i}.", n_heads)
    out = seq[1:] @ w["output_process.poseFinal.weight"].T + w["output_process.poseFinal.bias"]   # :467
    return out.reshape(T, B, F, 1).permute(1, 2, 3, 0).contiguous()     # :476-477


def cfg_forward(w, x, t, text_feat, scale, **kw):
    """ClassifierFreeSampleModel.forward: out_u + scale * (out_c - out_u)"""
    out_c = mdm_forward(w, x, t, text_feat, uncond=False, **kw)
    out_u = mdm_forward(w, x, t, text_feat, uncond=True, **kw)
    return out_u + scale.view(-1, 1, 1, 1) * (out_c - out_u)
