#!/usr/bin/env python
"""tests/golden/clip_text.npz: text features from Hugging Face transformers' CLIPTextModelWithProjection, an
independent implementation of the openai/CLIP text tower, on the seeded weights of clip_text_inputs.py.
openai/CLIP itself (the reference's pinned dependency) is not installed in the build container and there is no
network, so this is the strongest available pin for oracle/clip_text.py."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import clip_text_inputs as CI  # noqa: E402
from oracle import clip_text as OC  # noqa: E402


def hf_model(sd, width=512, layers=12, heads=8, embed_dim=512):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = CLIPTextConfig(vocab_size=CI.VOCAB, hidden_size=width, intermediate_size=4 * width, projection_dim=embed_dim,
                         num_hidden_layers=layers, num_attention_heads=heads, max_position_embeddings=CI.CTX,
                         hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0,
                         pad_token_id=1, bos_token_id=0, eos_token_id=2)  # eos_token_id=2 = "pool at argmax", as CLIP does
    m = CLIPTextModelWithProjection(cfg).eval()
    new = {"text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
           "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
           "text_model.final_layer_norm.weight": sd["ln_final.weight"],
           "text_model.final_layer_norm.bias": sd["ln_final.bias"],
           "text_projection.weight": sd["text_projection"].t().contiguous()}
    for l in range(layers):
        p, q = f"transformer.resblocks.{l}.", f"text_model.encoder.layers.{l}."
        wq, wk, wv = sd[p + "attn.in_proj_weight"].chunk(3, dim=0)
        bq, bk, bv = sd[p + "attn.in_proj_bias"].chunk(3, dim=0)
        new.update({q + "self_attn.q_proj.weight": wq, q + "self_attn.q_proj.bias": bq,
                    q + "self_attn.k_proj.weight": wk, q + "self_attn.k_proj.bias": bk,
                    q + "self_attn.v_proj.weight": wv, q + "self_attn.v_proj.bias": bv,
                    q + "self_attn.out_proj.weight": sd[p + "attn.out_proj.weight"],
                    q + "self_attn.out_proj.bias": sd[p + "attn.out_proj.bias"],
                    q + "layer_norm1.weight": sd[p + "ln_1.weight"], q + "layer_norm1.bias": sd[p + "ln_1.bias"],
                    q + "layer_norm2.weight": sd[p + "ln_2.weight"], q + "layer_norm2.bias": sd[p + "ln_2.bias"],
                    q + "mlp.fc1.weight": sd[p + "mlp.c_fc.weight"], q + "mlp.fc1.bias": sd[p + "mlp.c_fc.bias"],
                    q + "mlp.fc2.weight": sd[p + "mlp.c_proj.weight"], q + "mlp.fc2.bias": sd[p + "mlp.c_proj.bias"]})
    own = m.state_dict()
    extra = {k: v for k, v in own.items() if k not in new}
    assert all("position_ids" in k for k in extra), list(extra)
    missing, unexpected = m.load_state_dict(new, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


def main():
    import transformers
    torch.manual_seed(0)
    sd = CI.state_dict()
    m = hf_model(sd)
    arrays = {}
    for name, lengths in CI.CASES.items():
        tok = CI.tokens(lengths)
        with torch.no_grad():
            ref = m(input_ids=tok).text_embeds.float()
            got = OC.encode_text(sd, tok)
        err = float((got - ref).abs().max() / ref.abs().max())
        print(f"[{name}] oracle vs transformers {transformers.__version__} CLIPTextModelWithProjection: {err:.2e}  "
              f"|feat| max {float(ref.abs().max()):.3f}")
        assert err < 2e-5, err
        arrays[f"{name}/features"] = ref.numpy()
        arrays[f"{name}/tokens"] = tok.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "clip_text.npz"), **arrays)
    print("wrote tests/golden/clip_text.npz")


if __name__ == "__main__":
    main()
