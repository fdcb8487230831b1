"""CPU: host-side logic of the product (schedules, respacing, masks, factories, module layout) against the
reference's hashes, and the C-ABI library: loads, exports every symbol of include/mst.h, struct layouts agree."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import Args, sha16

import mst_b200
from mst_b200 import _lib as L
from mst_b200.data_loaders import bandai_posrot_utils, humanml_utils, stylexia_posrot_utils
from mst_b200.diffusion import gaussian_diffusion as gd
from mst_b200.diffusion.inpainting_gaussian_diffusion import InpaintingGaussianDiffusion
from mst_b200.diffusion.respace import SpacedDiffusion, _WrappedModel, space_timesteps
from mst_b200.utils import model_util as mu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schedule_tables_bit_exact(golden_hashes):
    d = mu.create_gaussian_diffusion(Args())
    for name in ["betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_variance",
                 "posterior_log_variance_clipped", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                 "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"]:
        assert sha16(getattr(d, name)) == golden_hashes["sched/cosine1000/" + name], name
    assert d.num_timesteps == 1000 and d.timestep_map == list(range(1000))
    assert d.model_mean_type == gd.ModelMeanType.START_X and d.model_var_type == gd.ModelVarType.FIXED_SMALL
    assert d.loss_type == gd.LossType.MSE


@pytest.mark.parametrize("spec", ["ddim20", "50", "ddim50", "100", "10,10,10", "ddim10", "25"])
def test_respacing_bit_exact(golden_hashes, spec):
    d = mu.create_gaussian_diffusion(Args(), SpacedDiffusion, timestep_respacing=spec)
    assert sha16(np.array(d.timestep_map, dtype=np.int64)) == golden_hashes[f"space/{spec}/map"]
    assert sha16(d.betas) == golden_hashes[f"space/{spec}/betas"]
    assert d.num_timesteps == len(d.timestep_map) and d.original_num_steps == 1000


def test_space_timesteps_edge_cases():
    assert space_timesteps(1000, [1000]) == set(range(1000))
    assert space_timesteps(10, "ddim5") == {0, 2, 4, 6, 8}
    assert space_timesteps(300, [10, 15, 20]) == space_timesteps(300, "10,15,20")
    assert space_timesteps(7, "1") == {0}
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim37")
    with pytest.raises(ValueError):
        space_timesteps(10, "20")
    with pytest.raises(NotImplementedError):
        gd.get_named_beta_schedule("quadratic", 10)


def test_masks_bit_exact(golden_hashes):
    mods = {"stylexia": stylexia_posrot_utils, "humanml": humanml_utils, "bandai": bandai_posrot_utils}
    n = 0
    for key, want in golden_hashes.items():
        if not key.startswith("mask/"):
            continue
        _, ds, name, shp = key.split("/")
        shape = tuple(int(v) for v in shp.split("x"))
        kw = dict(lengths=[40, 30], prefix_end=0.25, suffix_end=0.75) if name == "in_between" else {}
        m = mods[ds].get_inpainting_mask(name, shape, **kw)
        assert m.dtype == np.float64 and m.shape == shape
        assert f"{sha16(m)}:{m.sum():.0f}" == want, key
        n += 1
    assert n >= 15
    assert stylexia_posrot_utils.NUM_HML_FEATS == 181 and bandai_posrot_utils.NUM_HML_FEATS == 190
    assert humanml_utils.NUM_HML_FEATS == 263


def test_factories_and_model_layout():
    args = Args()
    kw = mu.get_transfer_args(args)
    assert kw["njoints"] == 181 and kw["nfeats"] == 1 and kw["data_rep"] == "hml_vec" and kw["cond_mode"] == "text"
    model, d1, d2 = mu.creat_ddpm_ddim_diffusion(args, ModelClass=mu.StyleDiffusion, timestep_respacing="ddim20")
    assert isinstance(d1, InpaintingGaussianDiffusion) and d1.num_timesteps == 20
    assert isinstance(d2, InpaintingGaussianDiffusion) and d2.num_timesteps == 1000
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 16822272
    assert len(trainable) == 96 and all(n.startswith("seqTransEncoder.layers.") for n in trainable)
    model2, s1, s2 = mu.creat_serval_diffusion(args, ModelClass=mu.StyleDiffusion, timestep_respacing="ddim20")
    assert isinstance(s1, InpaintingGaussianDiffusion) and type(s2) is SpacedDiffusion

    class _Data:
        class dataset:
            pass
    m, d = mu.create_model_and_diffusion(args, _Data())
    n_params = sum(p.numel() for n, p in m.named_parameters())
    assert n_params == 17796277  # non-CLIP parameters of the reference's stylexia MDM
    keys = set(m.state_dict().keys())
    for k in ["input_process.poseEmbedding.weight", "sequence_pos_encoder.pe", "embed_timestep.sequence_pos_encoder.pe",
              "embed_timestep.time_embed.0.weight", "embed_timestep.time_embed.2.bias", "embed_text.weight",
              "output_process.poseFinal.bias", "seqTransEncoder.layers.7.self_attn.in_proj_weight",
              "seqTransEncoder.layers.0.self_attn.out_proj.bias", "seqTransEncoder.layers.3.linear1.weight",
              "seqTransEncoder.layers.3.linear2.bias", "seqTransEncoder.layers.5.norm1.weight",
              "seqTransEncoder.layers.5.norm2.bias"]:
        assert k in keys, k
    from oracle.weights import mdm_state_dict
    missing, unexpected = m.load_state_dict(mdm_state_dict(181, seed=0), strict=False)
    assert not missing and not unexpected


def test_cfg_wrapper_contract():
    from mst_b200.model.cfg_sampler import ClassifierFreeSampleModel
    args = Args()

    class _Data:
        class dataset:
            pass
    m, _ = mu.create_model_and_diffusion(args, _Data())
    w = ClassifierFreeSampleModel(m)
    assert w.model is m and w.njoints == 181 and w.nfeats == 1 and w.data_rep == "hml_vec" and w.cond_mode == "text"
    m.cond_mask_prob = 0.0
    with pytest.raises(AssertionError):
        ClassifierFreeSampleModel(m)


def test_no_cpu_fallback():
    d = mu.create_gaussian_diffusion(Args(), InpaintingGaussianDiffusion, timestep_respacing="ddim20")
    x = torch.zeros(1, 181, 1, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        d.q_sample(x, torch.zeros(1, dtype=torch.long), noise=torch.zeros_like(x), model_kwargs={'y': {'inpainting_mask': x}})
    with pytest.raises(RuntimeError, match="CUDA"):
        d.p_sample(lambda *a, **k: x, x, torch.zeros(1, dtype=torch.long), model_kwargs={'y': {}})

    class _Data:
        class dataset:
            pass
    m, _ = mu.create_model_and_diffusion(Args(), _Data())
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, torch.zeros(1, dtype=torch.long), y={'text_feat': torch.zeros(1, 512)})
    with pytest.raises(RuntimeError, match="CUDA"):
        d.p_sample_loop(m, (1, 181, 1, 8), model_kwargs={'y': {}})


def test_wrapped_model_remaps_timesteps():
    seen = {}

    def fake(x, ts, **kw):
        seen["ts"] = ts
        return x
    w = _WrappedModel(fake, [0, 50, 100, 150], False, 1000)
    w(torch.zeros(3), torch.tensor([3, 0, 2]))
    assert seen["ts"].tolist() == [150, 0, 100]


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "mst.h")).read()
    declared = sorted(set(re.findall(r"\b(mst_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(L.EXPORTED), (declared, sorted(L.EXPORTED))
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.mst_version()


def test_abi_struct_sizes_agree():
    lib = L.load()
    sizes = [ctypes.c_size_t() for _ in range(4)]
    assert lib.mst_abi_sizes(*[ctypes.byref(s) for s in sizes]) == 0
    assert [s.value for s in sizes] == [ctypes.sizeof(L.ModelDesc), ctypes.sizeof(L.Weights),
                                        ctypes.sizeof(L.ForwardArgs), ctypes.sizeof(L.UpdateArgs)]


def test_argument_validation_without_gpu():
    lib = L.load()
    desc = L.ModelDesc(181, 512, 4, 1024, 8, 512, 5000, 7)
    h = ctypes.c_void_p()
    assert lib.mst_engine_create(ctypes.byref(desc), ctypes.byref(h)) == 1
    assert b"precision" in lib.mst_last_error()
    desc = L.ModelDesc(181, 256, 4, 1024, 8, 512, 5000, L.PREC_BF16)
    assert lib.mst_engine_create(ctypes.byref(desc), ctypes.byref(h)) == 3  # unsupported shape for the tcgen05 path
    desc = L.ModelDesc(181, 512, 4, 1024, 8, 512, 5000, L.PREC_BF16)
    assert lib.mst_engine_create(ctypes.byref(desc), ctypes.byref(h)) == 0
    nbytes = ctypes.c_size_t()
    assert lib.mst_engine_workspace_bytes(h, 128, 196, ctypes.byref(nbytes)) == 0 and nbytes.value > 100e6
    assert lib.mst_engine_packed_weight_bytes(h, ctypes.byref(nbytes)) == 0
    # bf16 [N,K] packs of the six GEMM weights per layer + the in/out projections, plus the transposed [K,N] packs
    # of the layer weights that the training backward (dX = dY W) reads
    # ... plus the fp16 copies of the three weights that multiply the sampler's fp16 residual stream (QKV, linear1, final)
    layer_w = 16822272 - 8 * 6656
    f16_w = 8 * (3 * 512 * 512 + 1024 * 512) + 192 * 512
    # ... plus the transposed in-/out-projection packs of the training backward
    assert abs(nbytes.value - (2 * layer_w + 4 * 192 * 512 + f16_w) * 2) < 64 * 1024
    assert lib.mst_engine_destroy(h) == 0
    a = L.UpdateArgs()
    assert lib.mst_update_step(ctypes.byref(a), None) == 1 and b"empty shape" in lib.mst_last_error()


def test_install_overlay_registers_reference_names():
    import sys
    names = mst_b200.install()
    assert "diffusion.respace" in names
    assert sys.modules["diffusion.respace"].space_timesteps is space_timesteps
    assert sys.modules["utils.model_util"].creat_ddpm_ddim_diffusion is mu.creat_ddpm_ddim_diffusion
    for n in names:
        sys.modules.pop(n, None)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout only exists in the build container")
def test_install_overlay_next_to_the_real_reference():
    """install('/root/reference'): the hot-path modules resolve to this package, everything else (here the reference's
    own data_loaders/tensors.py collate, which the demo script uses) keeps coming from the reference - and the mirror's
    collate builds the same model_kwargs."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np, torch; np.float = float; np.int = int\n"
        f"sys.path.insert(0, {REPO!r})\n"
        "import mst_b200\n"
        "mst_b200.install('/root/reference')\n"
        "from data_loaders.tensors import collate as ref_collate\n"
        "from utils import model_util\n"
        "from diffusion.respace import SpacedDiffusion\n"
        "from mst_b200.data_loaders.tensors import collate\n"
        "assert ref_collate.__module__ == 'data_loaders.tensors' and '/root/reference' in sys.modules['data_loaders.tensors'].__file__\n"
        "assert model_util.__name__.startswith('mst_b200') and SpacedDiffusion.__module__.startswith('mst_b200')\n"
        "items = [{'inp': torch.randn(181, 1, 76), 'tokens': None, 'lengths': 76, 'text': 'a'},\n"
        "         {'inp': torch.randn(181, 1, 60), 'tokens': None, 'lengths': 60, 'text': 'b'}]\n"
        "m0, k0 = ref_collate(items); m1, k1 = collate(items)\n"
        "assert torch.equal(m0, m1) and sorted(k0['y']) == sorted(k1['y'])\n"
        "assert torch.equal(k0['y']['mask'].float(), k1['y']['mask'].float()) and torch.equal(k0['y']['lengths'], k1['y']['lengths'])\n"
        "assert k0['y']['text'] == k1['y']['text']\n"
        "print('overlay ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "overlay ok" in out.stdout, out.stderr[-2000:]


def test_text_feature_cache_encodes_each_caption_once():
    """Scope row N1 (cached-feature API): CLIP is frozen, so a caption is encoded once per process."""
    import torch
    from mst_b200.model.mdm_forstyledataset import NativeDenoiser

    class Fake(NativeDenoiser):
        cond_mode, training, cond_mask_prob = "text", False, 0.0

        def __init__(self):
            torch.nn.Module.__init__(self)
            self.calls = []

        def encode_text(self, raw_text):
            self.calls.append(list(raw_text))
            return torch.stack([torch.full((4,), float(len(t))) for t in raw_text])

    m = Fake()
    a = m.encode_text_cached(["walk", "jump", "walk"], "cpu")
    assert m.calls == [["walk", "jump"]] and a.shape == (3, 4) and torch.equal(a[0], a[2])
    b = m.encode_text_cached(["jump", "run"], "cpu")
    assert m.calls == [["walk", "jump"], ["run"]] and float(b[1, 0]) == 3.0
    m.mst_text_cache_clear()
    m.encode_text_cached(["jump"], "cpu")
    assert m.calls[-1] == ["jump"]


def test_clip_text_abi_and_module_layout():
    """Scope row N1 without a GPU: struct sizes, argument validation, and the CLIP state_dict layout of the host mirror."""
    lib = L.load()
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.mst_abi_sizes_clip_text(ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (ctypes.sizeof(L.ClipTextDesc), ctypes.sizeof(L.ClipTextWeights))
    h = ctypes.c_void_p()
    bad = L.ClipTextDesc(49408, 77, 512, 4, 12, 2048, 512, L.PREC_FP32)  # head_dim 128
    assert lib.mst_clip_text_create(ctypes.byref(bad), ctypes.byref(h)) == 3 and b"head_dim" in lib.mst_last_error()
    bad = L.ClipTextDesc(49408, 128, 512, 8, 12, 2048, 512, L.PREC_FP32)
    assert lib.mst_clip_text_create(ctypes.byref(bad), ctypes.byref(h)) == 3 and b"context" in lib.mst_last_error()
    ok = L.ClipTextDesc(49408, 77, 512, 8, 12, 2048, 512, L.PREC_BF16)
    assert lib.mst_clip_text_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    n = ctypes.c_size_t()
    assert lib.mst_clip_text_packed_weight_bytes(h, ctypes.byref(n)) == 0
    assert n.value == 12 * (3 * 512 * 512 + 512 * 512 + 2 * 2048 * 512) * 2
    assert lib.mst_clip_text_workspace_bytes(h, 64, ctypes.byref(n)) == 0 and n.value > 64 * 77 * 512 * 4 * 4
    assert lib.mst_clip_text_encode(h, None, 1, None, None, 0, None) == 1  # null arguments are refused before any launch
    assert lib.mst_clip_text_destroy(h) == 0

    from mst_b200.model.clip_text import CLIPTextTower
    tower = CLIPTextTower(transformer_layers=2)
    keys = set(tower.state_dict())
    assert {"token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection",
            "transformer.resblocks.1.attn.in_proj_weight", "transformer.resblocks.1.attn.out_proj.bias",
            "transformer.resblocks.0.mlp.c_fc.weight", "transformer.resblocks.0.mlp.c_proj.bias",
            "transformer.resblocks.0.ln_1.weight", "transformer.resblocks.0.ln_2.bias"} <= keys and len(keys) == 29
    assert not any(p.requires_grad for p in tower.parameters())
    rebuilt = CLIPTextTower.from_state_dict({**tower.state_dict(), "visual.proj": torch.zeros(1)})
    assert rebuilt.transformer.layers == 2 and rebuilt.heads == 8
    with pytest.raises(RuntimeError, match="CUDA"):  # no CPU fallback
        tower.encode_text(torch.zeros(1, 77, dtype=torch.long))
