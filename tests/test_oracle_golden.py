"""CPU: the oracle against the fixtures generated from the real reference (tests/golden/make_golden.py)
and against the reference's own hashes listed in SURVEY section 8(c)."""
import numpy as np
import pytest
import torch

from helpers import relerr, sha16
from oracle import denoiser as OD
from oracle import sampler as OS
from oracle import schedule as OSch
from oracle.weights import NoiseTape, checksum, mdm_state_dict

SURVEY_HASHES = {  # SURVEY.md section 8(c), probed from the reference
    "betas": "9e50a88ff4dcc6a4", "abar": "3c1e39abad8af881", "coef1": "e87e8c966f98e943", "coef2": "478925d26e51386a",
    "post_var": "35e7d283e41f7ccf", "post_logvar_clipped": "dce018658b9eaed2", "sqrt_abar": "a7aff4f06c4035aa",
    "sqrt_1m_abar": "1cf6cb7eeaa20a8b",
}


@pytest.fixture(scope="module")
def state():
    return mdm_state_dict(n_feats=181, seed=0)


def test_schedule_tables_match_survey_hashes():
    s = OSch.Schedule(OSch.cosine_betas(1000))
    for name, h in SURVEY_HASHES.items():
        assert sha16(getattr(s, name)) == h, name
    assert s.betas[0] == 4.128422482196914e-05 and s.betas[999] == 0.999
    assert s.coef1[0] == 1 and s.coef1[1] == 0.5277814093344871 and s.coef2[0] == 0
    assert s.post_logvar_clipped[0] == s.post_logvar_clipped[1] == -10.734082532465003


def test_space_timesteps_match_survey_hashes():
    f = lambda spec: sha16(np.array(OSch.space_timesteps(1000, spec), dtype=np.int64))
    assert f("ddim20") == "327628f6d1264892"
    assert f("50") == "c5bdf9b959c7973b"
    assert f("ddim50") == "ebc60f0baaa7a6bc"
    assert f("100") == "432f07a847b37dff"
    assert f("10,10,10") == "6b000b2b11a4f157"
    assert f([1000]) == "702746827e553786"
    with pytest.raises(ValueError):  # no integer stride yields 37 steps (the reference agrees: golden hashes.txt)
        OSch.space_timesteps(1000, "ddim37")
    assert len(OSch.space_timesteps(1000, "ddim30")) == 30  # SURVEY lists this as an error; the reference returns stride 34


def test_respaced_schedules_match_reference(golden_hashes):
    assert golden_hashes["space/ddim30"] == "len=30" and golden_hashes["space/ddim37"] == "ValueError"
    for spec in ["ddim20", "50", "ddim50", "100", "10,10,10", "ddim10", "25"]:
        s = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, spec))
        assert sha16(np.array(s.timestep_map, dtype=np.int64)) == golden_hashes[f"space/{spec}/map"]
        assert sha16(s.betas) == golden_hashes[f"space/{spec}/betas"]
    # create_gaussian_diffusion returns a SpacedDiffusion: its betas are re-derived as 1 - abar_i/abar_{i-1}
    assert sha16(OSch.Schedule(OSch.linear_betas(1000)).betas) == golden_hashes["sched/linear1000/betas"]


def test_weight_recipe_is_reproducible(state, golden_hashes):
    assert repr(checksum(state)) == golden_hashes["weights/seed0/checksum"]


def test_denoiser_forward_matches_reference(state, golden):
    x, t, feat = (torch.from_numpy(golden[k]) for k in ("fwd_x", "fwd_t", "fwd_feat"))
    assert relerr(OD.mdm_forward(state, x, t, feat), golden["fwd_out_cond"]) < 2e-5
    assert relerr(OD.mdm_forward(state, x, t, feat, uncond=True), golden["fwd_out_uncond"]) < 2e-5


@pytest.mark.parametrize("clip", [False, True])
def test_cfg_inpainting_trajectory_matches_reference(state, golden, clip):
    x_inp, mask, scale, feat = (torch.from_numpy(golden[k]) for k in ("traj_x_inp", "traj_mask", "traj_scale", "traj_feat"))
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, "8"))
    final, xs = OS.sample_loop(sch, lambda xx, tt: OD.cfg_forward(state, xx, tt, feat, scale), tuple(x_inp.shape),
                               NoiseTape(5), mask=mask, x_inp=x_inp, clip=clip)
    ref = golden[f"traj_ddpm8_clip{int(clip)}_xstart"]
    assert len(xs) == ref.shape[0] == 8
    assert max(relerr(a, b) for a, b in zip(xs, ref)) < 5e-5
    assert relerr(final, golden[f"traj_ddpm8_clip{int(clip)}_final"]) < 5e-5


def test_demo_ddim_path_matches_reference(state, golden):
    content, style, mask, feat = (torch.from_numpy(golden[k]) for k in ("demo_content", "demo_style", "demo_mask", "demo_feat"))
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, "ddim20"))
    _, xs = OS.sample_loop(sch, lambda xx, tt: OD.mdm_forward(state, xx, tt, feat), tuple(content.shape), NoiseTape(9),
                           ddim=True, mask=mask, x_inp=style, skip_timesteps=14, init_image=content)
    assert len(xs) == 6
    assert max(relerr(a, b) for a, b in zip(xs, golden["demo_xstart"])) < 5e-5


def test_oracle_decode_matches_reference_golden():
    """oracle/decode.py (inv_transform + recover_from_ric) against the real reference's outputs (tests/golden/decode.npz)."""
    import os
    import numpy as np
    import torch
    from oracle import decode as OD
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode.npz"))
    for name, F, T, J in [("stylexia", 181, 76, 20), ("humanml", 263, 196, 22), ("bandai", 190, 60, 21)]:
        g = torch.Generator().manual_seed(0)
        sample = torch.randn(3, F, 1, T, generator=g)
        mean = torch.randn(F, generator=g) * 0.3
        std = torch.rand(F, generator=g) * 0.5 + 0.05
        got = OD.decode_motion(sample, mean, std, J)
        want = torch.from_numpy(gold[f"{name}/joints"])
        assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max())


def test_oracle_clip_text_matches_independent_golden():
    """oracle/clip_text.py (the restated openai/CLIP text tower, scope row N1) against tests/golden/clip_text.npz, which
    tests/golden/make_golden_clip_text.py produced with Hugging Face transformers' CLIPTextModelWithProjection on the
    same seeded weights.  Tolerance: 2e-5 of the largest feature (two fp32 implementations, 12 layers)."""
    import os
    import sys
    import numpy as np
    import torch
    from oracle import clip_text as OC
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import clip_text_inputs as CI
    gold = np.load(os.path.join(here, "golden", "clip_text.npz"))
    sd = CI.state_dict()
    for name, lengths in CI.CASES.items():
        tok = CI.tokens(lengths)
        assert np.array_equal(tok.numpy().astype(np.int32), gold[f"{name}/tokens"])  # the generator is reproducible
        got = OC.encode_text(sd, tok)
        want = torch.from_numpy(gold[f"{name}/features"])
        assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
    # the causal mask: features of a prompt do not depend on what follows its end-of-text token
    tok = CI.tokens([9])
    tok2 = tok.clone()
    tok2[0, 20:30] = 5
    assert torch.equal(OC.encode_text(sd, tok), OC.encode_text(sd, tok2))


def test_oracle_matches_headline_shape_golden(state):
    """B=64, F=181, T=196, CFG scale 2.5 + root_horizontal inpainting: one DDPM step of the REAL reference
    (tests/golden/make_golden_b64.py); four samples in full and three digests of all 64."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from b64_inputs import KEEP, b64_digests, b64_inputs
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "b64_step.npz")))
    inp = b64_inputs()
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    with torch.no_grad():
        _, xs = OS.sample_loop(sch, lambda xx, tt: OD.cfg_forward(state, xx, tt, inp["feat"], inp["scale"]), inp["shape"],
                               NoiseTape(3), mask=inp["mask"], x_inp=inp["x_inp"], stop_timesteps=999)
    assert relerr(xs[0][KEEP], gold["x0_keep"]) < 5e-5
    dig = b64_digests(xs[0])
    for k in ("sum", "l2", "probe"):
        scale = np.abs(gold["l2"]) if k != "probe" else np.abs(gold["l2"]) * 190.0  # |probe| ~ sqrt(F*T) = 188
        assert np.max(np.abs(dig[k].numpy() - gold[k]) / scale) < 1e-4, k
