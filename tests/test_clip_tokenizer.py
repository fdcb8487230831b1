"""CPU: the native CLIP BPE tokenizer (scope row N1) against an INDEPENDENT implementation of the same algorithm,
Hugging Face transformers' ``CLIPTokenizer``, on a shared synthetic vocabulary (the real
``bpe_simple_vocab_16e6.txt.gz`` belongs to the absent openai/CLIP package), plus ``clip.tokenize``'s framing rules."""
import gzip
import json
import os

import pytest
import torch

from mst_b200.model.clip_tokenizer import SimpleTokenizer, bytes_to_unicode

MERGES = [("t", "h"), ("th", "e</w>"), ("a", "n"), ("an", "d</w>"), ("w", "a"), ("wa", "l"), ("wal", "k"),
          ("walk", "s</w>"), ("p", "e"), ("pe", "r"), ("per", "s"), ("pers", "o"), ("perso", "n</w>"), ("i", "n"),
          ("in", "g</w>"), ("l", "i"), ("li", "k"), ("lik", "e</w>"), ("o", "l"), ("ol", "d</w>"), ("m", "a"),
          ("ma", "n</w>"), ("j", "u"), ("ju", "m"), ("jum", "p"), ("jump", "s</w>"), ("h", "a"), ("ha", "p"),
          ("hap", "p"), ("happ", "i"), ("happi", "l"), ("happil", "y</w>"), ("'", "s</w>"),
          ("walk", "ing</w>"), ("!", "!</w>"), ("c", "a"), ("ca", "f"), ("caf", "é</w>")]
CAPTIONS = ["a person walks like an old man", "A person jumps happily!!", "the man's walking, and the 12 steps...",
            "  spaces   and\ttabs  ", "unicode café naïve", "&amp; html &lt;escapes&gt;", "", "x" * 300]


@pytest.fixture(scope="module")
def vocab_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("bpe")
    bpe = os.path.join(d, "bpe.txt.gz")
    with gzip.open(bpe, "wb") as f:
        f.write(("#version: synthetic\n" + "\n".join(" ".join(m) for m in MERGES) + "\n").encode("utf-8"))
    return str(d), bpe


def test_vocabulary_layout(vocab_files):
    _, bpe = vocab_files
    tok = SimpleTokenizer(bpe)
    assert len(tok.encoder) == 512 + len(MERGES) + 2
    assert tok.sot_token == 512 + len(MERGES) and tok.eot_token == tok.sot_token + 1
    b2u = bytes_to_unicode()
    assert len(b2u) == 256 and len(set(b2u.values())) == 256 and b2u[ord("a")] == "a" and b2u[ord(" ")] == "Ġ"
    # with the stock 48894 merges the layout gives CLIP's 49408-entry vocabulary and <sot> = 49406, <eot> = 49407
    assert 512 + (49152 - 256 - 2) + 2 == 49408


def test_matches_transformers_clip_tokenizer(vocab_files):
    transformers = pytest.importorskip("transformers")
    d, bpe = vocab_files
    tok = SimpleTokenizer(bpe)
    with open(os.path.join(d, "vocab.json"), "w") as f:
        json.dump(tok.encoder, f)
    with open(os.path.join(d, "merges.txt"), "w") as f:
        f.write("#version: synthetic\n" + "\n".join(" ".join(m) for m in MERGES) + "\n")
    hf = transformers.CLIPTokenizer(os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt"))
    for text in CAPTIONS:
        if "&" in text:  # openai/CLIP html-unescapes captions (basic_clean); transformers only does with ftfy installed
            assert tok.encode(text) == tok.encode("& html <escapes>")
            continue
        mine = tok.encode(text)
        theirs = hf(text, add_special_tokens=False)["input_ids"]
        assert mine == theirs, (text, mine, theirs)
    framed = tok.tokenize(CAPTIONS[:3], context_length=77, truncate=True)
    ref = hf(CAPTIONS[:3], padding="max_length", max_length=77, truncation=True)
    for row, ids, att in zip(framed, ref["input_ids"], ref["attention_mask"]):
        n = sum(att)
        assert row[:n].tolist() == ids[:n] and int(row[n:].abs().sum()) == 0  # CLIP pads with 0, transformers with <eot>


def test_tokenize_framing_and_truncation(vocab_files):
    _, bpe = vocab_files
    tok = SimpleTokenizer(bpe)
    out = tok.tokenize(["a person walks", ""], context_length=22, truncate=True)
    assert out.dtype == torch.int64 and out.shape == (2, 22)
    assert out[0, 0] == tok.sot_token and out[1, 0] == tok.sot_token and out[1, 1] == tok.eot_token and out[1, 2:].sum() == 0
    n = int((out[0] != 0).sum())
    assert out[0, n - 1] == tok.eot_token and tok.decode(out[0, 1:n - 1].tolist()).strip() == "a person walks"
    assert int(out[0].argmax()) == n - 1  # the text tower pools at arg-max = <eot>
    long = tok.tokenize("x" * 300, context_length=77, truncate=True)
    assert long[0, -1] == tok.eot_token and long[0, 0] == tok.sot_token
    with pytest.raises(RuntimeError, match="too long"):
        tok.tokenize("x" * 300, context_length=77, truncate=False)
    # humanml / kit path of MDM.encode_text: context 22, zero padded to 77 by the caller
    assert tok.tokenize("walks", context_length=22, truncate=True).shape == (1, 22)


def test_missing_vocabulary_is_a_clear_error(monkeypatch):
    monkeypatch.delenv("MST_CLIP_BPE", raising=False)
    with pytest.raises(FileNotFoundError, match="MST_CLIP_BPE"):
        SimpleTokenizer()
