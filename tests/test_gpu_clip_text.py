"""GPU: the CLIP text tower (scope row N1, csrc/text.cu) through the C-ABI against oracle/clip_text.py and the
golden features of tests/golden/clip_text.npz (made with an independent implementation, see make_golden_clip_text.py).

Tolerances: fp32 path <= 1e-4, bf16 tcgen05 path <= 2e-2 (max abs error / max abs feature), BASELINE north_star's."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import relerr
from oracle import clip_text as OC

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import clip_text_inputs as CI  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module")
def sd():
    return CI.state_dict()


@pytest.fixture(scope="module")
def towers(sd):
    from mst_b200.model.clip_text import CLIPTextTower
    torch.cuda.set_device(0)
    return {p: CLIPTextTower.from_state_dict(sd, precision=p).to(DEV) for p in ("fp32", "bf16")}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_features_match_golden_and_oracle(towers, sd, precision):
    gold = np.load(os.path.join(HERE, "golden", "clip_text.npz"))
    for name, lengths in CI.CASES.items():
        tok = CI.tokens(lengths)
        got = towers[precision].encode_text(tok)  # host tokens, as clip.tokenize returns them
        assert got.shape == (len(lengths), 512) and got.dtype == torch.float32 and got.is_cuda
        assert relerr(got, gold[f"{name}/features"]) <= TOL[precision]
        assert relerr(got, OC.encode_text(sd, tok)) <= TOL[precision]


def test_batch_of_64_captions_and_row_independence(towers, sd):
    """BASELINE configs[1] encodes 64 captions per trajectory; a caption's feature must not depend on its batch-mates."""
    g = torch.Generator().manual_seed(5)
    lengths = torch.randint(0, 76, (64,), generator=g).tolist()
    tok = CI.tokens(lengths, seed=7)
    want = OC.encode_text(sd, tok)
    got = towers["fp32"].encode_text(tok.to(DEV))  # device tokens are accepted too
    assert relerr(got, want) <= 1e-4
    single = torch.cat([towers["fp32"].encode_text(tok[i:i + 1]) for i in (0, 17, 63)])
    assert relerr(single, got[[0, 17, 63]]) <= 2e-6  # different GEMM tilings for M=77 and M=4928: fp32 rounding only
    assert relerr(towers["bf16"].encode_text(tok), want) <= 2e-2


def test_causal_mask_and_eot_pooling(towers):
    tok = CI.tokens([9])
    after = tok.clone()
    after[0, 20:30] = 5  # tokens behind the end-of-text token cannot influence it (causal attention)
    a, b = towers["fp32"].encode_text(tok), towers["fp32"].encode_text(after)
    assert torch.equal(a, b)
    moved = tok.clone()
    moved[0, 10], moved[0, 40] = 0, CI.EOT  # pooling follows the arg-max token, so moving <eot> changes the feature
    assert relerr(towers["fp32"].encode_text(moved), a) > 1e-2
    twice = tok.clone()
    twice[0, 50] = CI.EOT  # ties: torch.argmax semantics = the first occurrence, which the later copy cannot influence
    assert torch.equal(towers["fp32"].encode_text(twice), a)


def test_reduced_geometry(sd):
    """width 256 / 4 heads / 3 layers / 40 positions / 300 tokens: nothing is hard-wired to ViT-B/32 except head_dim 64."""
    from mst_b200.model.clip_text import CLIPTextTower
    small = CI.state_dict(seed=3, width=256, layers=3, embed_dim=128, vocab=300, ctx=40)
    g = torch.Generator().manual_seed(2)
    tok = torch.randint(1, 298, (6, 40), generator=g)
    tok[torch.arange(6), torch.tensor([1, 39, 5, 17, 20, 33])] = 299
    want = OC.encode_text(small, tok)
    for precision in ("fp32", "bf16"):
        tower = CLIPTextTower.from_state_dict(small, precision=precision).to(DEV)
        assert relerr(tower.encode_text(tok), want) <= TOL[precision]


def test_bad_tokens_are_refused(towers):
    bad = CI.tokens([3])
    bad[0, 2] = CI.VOCAB
    with pytest.raises(IndexError):
        towers["fp32"].encode_text(bad)
    with pytest.raises(ValueError):
        towers["fp32"].encode_text(torch.zeros(2, 22, dtype=torch.long))


def test_mdm_encode_text_runs_on_the_native_tower(towers, sd):
    """MDM.encode_text (reference model/mdm_forstyledataset.py:298-313) with the native tower attached: tokenising is
    the caller's (clip.tokenize; a deterministic stand-in here), the 20-word cap + zero padding of humanml is the
    reference's, and the features reach the sampler through the per-caption cache."""
    from helpers import Args
    from mst_b200.utils import model_util as mu
    from mst_b200.model.mdm_forstyledataset import MDM

    def fake_tokenize(texts, context_length=77, truncate=True):
        out = torch.zeros(len(texts), context_length, dtype=torch.int64)
        for i, t in enumerate(texts):
            ids = [CI.SOT] + [1 + (hash_ % (CI.SOT - 1)) for hash_ in (sum(map(ord, w)) * 7919 for w in t.split())]
            ids = ids[:context_length - 1] + [CI.EOT]
            out[i, :len(ids)] = torch.tensor(ids)
        return out

    args = Args()
    model = MDM(**mu.get_transfer_args(args)).to(DEV).eval()
    model.clip_model = towers["fp32"]
    model.mst_tokenize = fake_tokenize
    texts = ["a person walks forward proudly", "a person jumps", "a person walks forward proudly"]
    feat = model.encode_text(texts)
    want = OC.encode_text(sd, fake_tokenize(texts))
    assert feat.shape == (3, 512) and relerr(feat, want) <= 1e-4
    cached = model.encode_text_cached(texts, torch.device(DEV))
    assert torch.equal(cached[0], cached[2]) and relerr(cached, want) <= 1e-4
    model.dataset = "humanml"  # 22-token context padded back to 77 (reference :301-309)
    long_text = [" ".join(["word%d" % i for i in range(40)])]
    feat_h = model.encode_text(long_text)
    tok_h = torch.cat([fake_tokenize(long_text, context_length=22), torch.zeros(1, 55, dtype=torch.int64)], dim=1)
    assert relerr(feat_h, OC.encode_text(sd, tok_h)) <= 1e-4


def test_captions_to_features_without_the_clip_package(tmp_path):
    """Row N1 end to end: captions -> native BPE tokenizer (a vocabulary FILE, synthetic here) -> native text tower ->
    MDM.encode_text, with no `clip` import anywhere; checked against the oracle tower on the same tokens."""
    import gzip
    from helpers import Args
    from mst_b200.model.clip_text import CLIPTextTower
    from mst_b200.model.clip_tokenizer import SimpleTokenizer, attach_tokenizer
    from mst_b200.model.mdm_forstyledataset import MDM
    from mst_b200.utils import model_util as mu
    merges = [("t", "h"), ("th", "e</w>"), ("w", "a"), ("wa", "l"), ("wal", "k"), ("walk", "s</w>"), ("p", "e"), ("pe", "r"),
              ("per", "s"), ("pers", "o"), ("perso", "n</w>"), ("j", "u"), ("ju", "m"), ("jum", "p"), ("jump", "s</w>")]
    bpe = os.path.join(tmp_path, "bpe_simple_vocab.txt.gz")
    with gzip.open(bpe, "wb") as f:
        f.write(("#version: synthetic\n" + "\n".join(" ".join(m) for m in merges) + "\n").encode())
    tok = SimpleTokenizer(bpe)
    vocab = len(tok.encoder)
    small = CI.state_dict(seed=4, width=512, layers=2, embed_dim=512, vocab=vocab, ctx=77)
    model = MDM(**mu.get_transfer_args(Args())).to(DEV).eval()
    model.clip_model = CLIPTextTower.from_state_dict(small, precision="fp32").to(DEV)
    attach_tokenizer(model, bpe)
    texts = ["a person walks", "the person jumps and walks", "a person walks"]
    feat = model.encode_text(texts)
    want = OC.encode_text(small, tok.tokenize(texts, truncate=True))
    assert feat.shape == (3, 512) and relerr(feat, want) <= 1e-4 and torch.equal(feat[0], feat[2])
