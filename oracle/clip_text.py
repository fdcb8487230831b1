"""torch-CPU restatement of the CLIP text tower (oracle; test infrastructure only).

The reference calls it as ``self.clip_model.encode_text(texts)`` (model/mdm_forstyledataset.py:298-313); the
arithmetic lives in the third-party dependency openai/CLIP pinned at commit a9b1bf59 (requirements.txt:26), which is
not under /root/reference.  This file restates the published algorithm of clip/model.py:

* CLIP.encode_text          x = token_embedding(text) + positional_embedding -> transformer -> ln_final ->
                            x[arange(B), text.argmax(-1)] @ text_projection
* CLIP.build_attention_mask additive mask, -inf strictly above the diagonal (causal)
* ResidualAttentionBlock    x = x + attn(ln_1(x)); x = x + mlp(ln_2(x)); attn = nn.MultiheadAttention (packed in_proj,
                            heads of 64, scale 1/8); mlp = c_proj(QuickGELU(c_fc(x)))
* QuickGELU                 x * sigmoid(1.702 x)

Pinned by tests/golden/make_golden_clip_text.py against an INDEPENDENT implementation of the same model, Hugging
Face transformers' CLIPTextModelWithProjection (quick_gelu, eos pooling by argmax), on shared weights; the reference
repository itself holds no fixtures for this path."""
import torch
import torch.nn.functional as F


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def encode_text(sd, tokens, n_heads=None):
    """sd: CLIP state_dict (text-side keys, fp32); tokens: int64 [B, ctx] -> [B, embed_dim] fp32."""
    tokens = tokens.long()
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]  # [B, S, w]
    B, S, w = x.shape
    H = n_heads or w // 64
    dh = w // H
    n_layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
    mask = torch.full((S, S), float("-inf")).triu_(1)
    for l in range(n_layers):
        p = f"transformer.resblocks.{l}."
        h = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = h @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
        q, k, v = (t.reshape(B, S, H, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        a = torch.softmax(q @ k.transpose(-1, -2) / dh ** 0.5 + mask, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, S, w)
        x = x + a @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"]
        h = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        u = quick_gelu(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        x = x + u @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]
    x = F.layer_norm(x, (w,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(B), tokens.argmax(dim=-1)] @ sd["text_projection"]
