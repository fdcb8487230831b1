"""Seeded weights and token batches shared by make_golden_clip_text.py and the tests (no reference needed)."""
import torch

SOT, EOT, VOCAB, CTX = 49406, 49407, 49408, 77


def state_dict(seed=0, width=512, layers=12, embed_dim=512, vocab=VOCAB, ctx=CTX):
    """CLIP text-side state_dict with the initialisation scales of clip/model.py plus non-trivial biases / LN params."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    attn_std, proj_std, fc_std = width ** -0.5, (width ** -0.5) * ((2 * layers) ** -0.5), (2 * width) ** -0.5
    sd = {"token_embedding.weight": rn(vocab, width, std=0.02), "positional_embedding": rn(ctx, width, std=0.01)}
    for l in range(layers):
        p = f"transformer.resblocks.{l}."
        sd[p + "ln_1.weight"] = rn(width, std=0.1, mean=1.0)
        sd[p + "ln_1.bias"] = rn(width, std=0.05)
        sd[p + "attn.in_proj_weight"] = rn(3 * width, width, std=attn_std)
        sd[p + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
        sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
        sd[p + "ln_2.weight"] = rn(width, std=0.1, mean=1.0)
        sd[p + "ln_2.bias"] = rn(width, std=0.05)
        sd[p + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
    sd["ln_final.weight"] = rn(width, std=0.1, mean=1.0)
    sd["ln_final.bias"] = rn(width, std=0.05)
    sd["text_projection"] = rn(width, embed_dim, std=width ** -0.5)
    return sd


def tokens(lengths, seed=1, ctx=CTX):
    """What clip.tokenize returns: <sot> ids... <eot> then zero padding; `lengths` = number of word tokens per row."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(len(lengths), ctx, dtype=torch.int64)
    for i, n in enumerate(lengths):
        assert 0 <= n <= ctx - 2
        out[i, 0] = SOT
        out[i, 1:1 + n] = torch.randint(1, SOT, (n,), generator=g)
        out[i, 1 + n] = EOT
    return out


CASES = {
    "mixed": [7, 0, 75, 20, 33],   # a typical prompt, the empty string, a truncated 77-token prompt, humanml's 20-word cap
    "single": [12],
}
