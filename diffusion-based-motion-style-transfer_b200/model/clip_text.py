"""CLIP text tower (scope row N1): the frozen text encoder ``MDM.encode_text`` calls
(reference ``model/mdm_forstyledataset.py:275-286`` ``load_and_freeze_clip``, ``:298-313`` ``encode_text``).

The reference gets it from the third-party ``clip`` package (openai/CLIP @ a9b1bf59,
``requirements.txt:26``): ``clip.load('ViT-B/32')`` and then only ever calls ``.encode_text(tokens)``.
``CLIPTextTower`` is that text side as an ``nn.Module`` with the CLIP state_dict's own parameter names, so

    tower = CLIPTextTower.from_clip(clip_model)       # or  tower.load_state_dict(clip_sd, strict=False)
    model.clip_model = tower                           # MDM.encode_text then runs on the native kernels

Its ``encode_text`` runs on libmst_b200.so (csrc/text.cu); there is no torch fallback.  Tokenising stays with
``clip.tokenize`` (it needs the BPE vocabulary file that ships with the ``clip`` package).
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import engine as K


class _Block(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads)  # parameter container only (in_proj_*, out_proj.*)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.ModuleDict({"c_fc": nn.Linear(width, 4 * width), "c_proj": nn.Linear(4 * width, width)})


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([_Block(width, heads) for _ in range(layers)])


class CLIPTextTower(nn.Module):
    """Text side of ``clip.model.CLIP`` (ViT-B/32 defaults: 49408 tokens, 77 positions, width 512, 8 heads, 12 layers)."""

    def __init__(self, embed_dim: int = 512, context_length: int = 77, vocab_size: int = 49408,
                 transformer_width: int = 512, transformer_heads: int = 8, transformer_layers: int = 12,
                 precision: Optional[str] = None):
        super().__init__()
        self.context_length, self.vocab_size = context_length, vocab_size
        self.embed_dim, self.heads = embed_dim, transformer_heads
        self.transformer = _Transformer(transformer_width, transformer_layers, transformer_heads)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.mst_precision = precision
        self._eng = None
        self._eng_key = None
        self.initialize_parameters()
        for p in self.parameters():
            p.requires_grad = False

    def initialize_parameters(self):
        """The initialisation of clip/model.py CLIP.initialize_parameters (text side)."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        w, n = self.transformer.width, self.transformer.layers
        proj_std, attn_std, fc_std = (w ** -0.5) * ((2 * n) ** -0.5), w ** -0.5, (2 * w) ** -0.5
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp["c_fc"].weight, std=fc_std)
            nn.init.normal_(blk.mlp["c_proj"].weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    @classmethod
    def from_clip(cls, clip_model, precision: Optional[str] = None):
        """Build from a loaded ``clip.model.CLIP`` (its visual tower is ignored)."""
        sd = {k: v for k, v in clip_model.state_dict().items() if not k.startswith("visual.")}
        return cls.from_state_dict(sd, precision=precision)

    @classmethod
    def from_state_dict(cls, sd, precision: Optional[str] = None):
        width = sd["ln_final.weight"].shape[0]
        layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
        tower = cls(embed_dim=sd["text_projection"].shape[1], context_length=sd["positional_embedding"].shape[0],
                    vocab_size=sd["token_embedding.weight"].shape[0], transformer_width=width,
                    transformer_heads=width // 64, transformer_layers=layers, precision=precision)
        own = tower.state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"CLIP state_dict lacks text-tower entries: {missing[:4]}...")
        tower.load_state_dict({k: sd[k].float() for k in own})
        return tower

    @property
    def dtype(self):
        return torch.float32

    def mst_weights_changed(self):
        self._eng_key = None

    def _engine(self, device):
        key = (str(device), self.mst_precision or K.default_precision())
        if self._eng is None or self._eng_key != key:
            t = self.transformer
            eng = K.ClipTextEngine(self.vocab_size, self.context_length, t.width, self.heads, t.layers, 4 * t.width,
                                   self.embed_dim, precision=key[1], device=device)
            top = dict(token_embedding=self.token_embedding.weight, positional_embedding=self.positional_embedding,
                       lnf_g=self.ln_final.weight, lnf_b=self.ln_final.bias, text_projection=self.text_projection)
            layers = [dict(ln1_g=b.ln_1.weight, ln1_b=b.ln_1.bias, qkv_w=b.attn.in_proj_weight, qkv_b=b.attn.in_proj_bias,
                           o_w=b.attn.out_proj.weight, o_b=b.attn.out_proj.bias, ln2_g=b.ln_2.weight, ln2_b=b.ln_2.bias,
                           fc_w=b.mlp["c_fc"].weight, fc_b=b.mlp["c_fc"].bias, proj_w=b.mlp["c_proj"].weight,
                           proj_b=b.mlp["c_proj"].bias) for b in t.resblocks]
            eng.load_weights(top, layers)
            self._eng, self._eng_key = eng, key
        return self._eng

    def _apply(self, fn, *a, **k):
        self._eng_key = None  # .to() / .cuda() move the parameters the engine aliases
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        """text: integer tokens [B, context_length] -> features fp32 [B, embed_dim] (clip/model.py CLIP.encode_text)."""
        dev = self.text_projection.device
        if dev.type != "cuda":
            raise RuntimeError("CLIPTextTower.encode_text runs on the mst CUDA kernels only; move the module to a GPU")
        if text.dim() != 2 or text.shape[1] != self.context_length:
            raise ValueError(f"tokens must be [B, {self.context_length}], got {tuple(text.shape)}")
        if not text.is_cuda:  # the usual case (clip.tokenize output): validate ids for free on the host
            if text.numel() and (int(text.min()) < 0 or int(text.max()) >= self.vocab_size):
                raise IndexError("token id outside the vocabulary")
        tok = text.to(dev, torch.int32).contiguous()
        return self._engine(dev).encode(tok)

    def forward(self, text):
        return self.encode_text(text)
