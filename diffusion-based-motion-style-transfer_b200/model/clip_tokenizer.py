"""Byte-pair-encoding tokenizer of the CLIP text tower (scope row N1): ``clip.tokenize`` without the ``clip`` package.

The reference tokenises captions with ``clip.tokenize(raw_text, context_length=..., truncate=True)``
(``model/mdm_forstyledataset.py:298-313``; openai/CLIP @ a9b1bf59, ``requirements.txt:26`` - third-party and absent
here).  This module restates the published algorithm of ``clip/simple_tokenizer.py`` + ``clip.tokenize``:

* vocabulary = the 256 printable byte symbols, the same 256 with the end-of-word marker ``</w>``, one entry per merge
  of the BPE merges file (``bpe_simple_vocab_16e6.txt.gz``: first line is a header, the next 49152-256-2 lines are
  merges), ``<|startoftext|>`` and ``<|endoftext|>`` -> 49408 entries for the stock file;
* text is cleaned (html unescape, whitespace collapse), lower-cased, split with CLIP's regular expression, every
  piece is mapped to byte symbols and merged greedily by merge rank;
* ``tokenize`` frames every caption as ``<sot> tokens <eot>`` in a zero-padded ``[B, context_length]`` LongTensor,
  truncating (keeping ``<eot>`` last) or raising like the original.

The vocabulary file is NOT shipped (it belongs to the ``clip`` package): pass its path, or set ``MST_CLIP_BPE``.
``MDM`` picks the tokenizer up through its ``mst_tokenize`` hook (``attach_tokenizer``).

openai/CLIP also runs ``ftfy.fix_text`` on the caption; ``ftfy`` is used when importable and skipped otherwise (it
only repairs mojibake, plain ASCII captions are unaffected).
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import List, Union

import torch

try:  # CLIP's pattern needs \\p{L} / \\p{N}: the third-party `regex` module (present in this image)
    import regex as _re
    _PATTERN = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"
except ImportError:  # pragma: no cover - ASCII-equivalent fallback
    import re as _re
    _PATTERN = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[^\W\d_]+|\d|[^\s\w]+|_+"


@lru_cache()
def bytes_to_unicode():
    """byte value -> printable unicode character (the reversible byte alphabet of GPT-2 / CLIP BPE)"""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def get_pairs(word):
    """set of adjacent symbol pairs of a word (tuple of symbols)"""
    pairs, prev = set(), word[0]
    for ch in word[1:]:
        pairs.add((prev, ch))
        prev = ch
    return pairs


def basic_clean(text):
    try:
        import ftfy  # type: ignore
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    return html.unescape(html.unescape(text)).strip()


def whitespace_clean(text):
    return _re.sub(r"\s+", " ", text).strip()


def read_merges(bpe_path: str, n_merges: int = 49152 - 256 - 2):
    opener = gzip.open if bpe_path.endswith(".gz") else open
    with opener(bpe_path, "rb") as f:
        lines = f.read().decode("utf-8").split("\n")
    merges = lines[1:n_merges + 1]                     # line 0 is the "#version" header
    return [tuple(m.split()) for m in merges if len(m.split()) == 2]


class SimpleTokenizer:
    def __init__(self, bpe_path: str | None = None, merges=None):
        if merges is None:
            bpe_path = bpe_path or os.environ.get("MST_CLIP_BPE")
            if not bpe_path or not os.path.exists(bpe_path):
                raise FileNotFoundError("CLIP BPE vocabulary not found: pass bpe_path or set MST_CLIP_BPE to "
                                        "bpe_simple_vocab_16e6.txt.gz of the openai/CLIP package")
            merges = read_merges(bpe_path)
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        for m in merges:
            vocab.append("".join(m))
        vocab.extend(["<|startoftext|>", "<|endoftext|>"])
        self.encoder = dict(zip(vocab, range(len(vocab))))
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.bpe_ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = _re.compile(_PATTERN, _re.IGNORECASE)
        self.sot_token, self.eot_token = self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"]

    def bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = get_pairs(word) if len(word) > 1 else set()
        if not pairs:
            return token + "</w>"
        while True:
            bigram = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new_word, i = [], 0
            while i < len(word):
                try:
                    j = word.index(first, i)
                except ValueError:
                    new_word.extend(word[i:])
                    break
                new_word.extend(word[i:j])
                i = j
                if word[i] == first and i < len(word) - 1 and word[i + 1] == second:
                    new_word.append(first + second)
                    i += 2
                else:
                    new_word.append(word[i])
                    i += 1
            word = tuple(new_word)
            if len(word) == 1:
                break
            pairs = get_pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text) -> List[int]:
        ids = []
        text = whitespace_clean(basic_clean(text)).lower()
        for token in _re.findall(self.pat, text):
            token = "".join(self.byte_encoder[b] for b in token.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(token).split(" "))
        return ids

    def decode(self, tokens) -> str:
        text = "".join(self.decoder[int(t)] for t in tokens)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")

    def tokenize(self, texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
        """``clip.tokenize``: [len(texts), context_length] int64, ``<sot> ... <eot>`` zero padded."""
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            toks = [self.sot_token] + self.encode(t) + [self.eot_token]
            if len(toks) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {t} is too long for context length {context_length}")
                toks = toks[:context_length]
                toks[-1] = self.eot_token
            out[i, :len(toks)] = torch.tensor(toks)
        return out

    __call__ = tokenize


def attach_tokenizer(model, bpe_path: str | None = None):
    """Give ``model`` (an ``MDM`` / ``StyleDiffusion`` / CFG wrapper) this tokenizer: captions in ``y['text']`` are then
    tokenised here and encoded by the native text tower, with no ``clip`` package involved."""
    tok = SimpleTokenizer(bpe_path)
    target = getattr(model, "model", model)
    front = target._mst_front() if hasattr(target, "_mst_front") else target
    front.mst_tokenize = tok.tokenize
    return tok
