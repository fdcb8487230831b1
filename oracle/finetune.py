"""torch-CPU restatement of the few-shot style finetune step (oracle; test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Plain tensor algebra on top of oracle.denoiser; the gradients come from torch autograd over that algebra, which
is what the reference itself does (it differentiates nn.TransformerEncoder through autograd).

* few_shot_style_finetune_losses   reference diffusion/gaussian_diffusion.py:1317-1399
* masked_l2                        reference diffusion/gaussian_diffusion.py:223-235
* ddim_sample_with_grad            reference diffusion/inpainting_gaussian_diffusion.py:176-239
* p_sample_with_grad               reference diffusion/inpainting_gaussian_diffusion.py:66-123
* MotionEncoder.forward            reference model/mdm_forstyledataset.py:89-124
* AdamW (lr, wd=0) + norms         reference train/training_loop.py:97-99, diffusion/fp16_util.py:208-223

Pinned by tests/golden/make_golden_finetune.py, which runs the unmodified reference (StyleDiffusion +
InpaintingGaussianDiffusion.few_shot_style_finetune_losses + loss.backward()) on the same weights / noise tape and
asserts agreement before writing tests/golden/finetune.npz.
"""
import math

import torch

from . import denoiser as OD
from . import sampler as OS


def masked_l2(a, b, mask):
    """a, b: [R, J, Jdim, T]; mask: [R, 1, 1, T] -> [R]"""
    loss = ((a - b) ** 2 * mask.float()).flatten(1).sum(dim=1)
    n_entries = a.shape[1] * a.shape[2]
    return loss / (mask.flatten(1).sum(dim=1) * n_entries)


def motion_encoder_forward(w_front, w_enc, x, frame_mask, n_heads=4, enc_prefix="seqTransEncoder.layers."):
    """mu of MotionEncoder.forward.  w_front: the frozen mdm_model's state dict (input_process, pe);
    w_enc: dict with 'muQuery', 'sigmaQuery' [1,d] and the encoder's own layers; frame_mask: bool [B,T]."""
    B, F, _, T = x.shape
    xs = x.permute(3, 0, 1, 2).reshape(T, B, F)
    xs = xs @ w_front["input_process.poseEmbedding.weight"].T + w_front["input_process.poseEmbedding.bias"]
    mu_q = w_enc["muQuery"][:1][None].repeat(1, B, 1)
    sg_q = w_enc["sigmaQuery"][:1][None].repeat(1, B, 1)
    seq = torch.cat((mu_q, sg_q, xs), dim=0)
    seq = seq + w_front["sequence_pos_encoder.pe"][:T + 2]
    valid = torch.cat((torch.ones(B, 2, dtype=torch.bool), frame_mask), dim=1)     # [B, S]
    n_layers = 1 + max(int(k[len(enc_prefix):].split(".")[0]) for k in w_enc if k.startswith(enc_prefix))
    for i in range(n_layers):
        seq = _encoder_layer_masked(seq, w_enc, f"{enc_prefix}{i}.", n_heads, valid)
    return seq[0]


def _encoder_layer_masked(x, w, pre, n_heads, valid):
    """oracle.denoiser.encoder_layer with src_key_padding_mask = ~valid (keys only)."""
    S, B, d = x.shape
    dh = d // n_heads
    qkv = x @ w[pre + "self_attn.in_proj_weight"].T + w[pre + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)

    def heads(t):
        return t.reshape(S, B, n_heads, dh).permute(1, 2, 0, 3)

    q, k, v = heads(q), heads(k), heads(v)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    sc = sc.masked_fill(~valid[:, None, None, :], float("-inf"))
    att = torch.softmax(sc, dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(S, B, d)
    sa = o @ w[pre + "self_attn.out_proj.weight"].T + w[pre + "self_attn.out_proj.bias"]
    x = OD.layer_norm(x + sa, w[pre + "norm1.weight"], w[pre + "norm1.bias"])
    h = OD.gelu(x @ w[pre + "linear1.weight"].T + w[pre + "linear1.bias"])
    ff = h @ w[pre + "linear2.weight"].T + w[pre + "linear2.bias"]
    return OD.layer_norm(x + ff, w[pre + "norm2.weight"], w[pre + "norm2.bias"])


def finetune_losses(sch, w_front, w_enc, w_menc, x_start, t, x_content, x_style, text_feat_style, frame_mask_style,
                    inp_mask, tape, *, text_feat_t2m=None, frame_mask_t2m=None, inp_mask_t2m=None, skip_steps=700,
                    semantic_guidance=0, use_ddim=1, Ls=10.0, noise_t2m=None):
    """Loss terms of the reference function.  sch: oracle.schedule.Schedule of the (respaced) process; w_front /
    w_enc: frozen MDM front-end and the trainable StyleDiffusion encoder (tensors requiring grad); w_menc: the
    MotionEncoder's muQuery/sigmaQuery/layers; tape: NoiseTape supplying the N(0,1) draws in the reference's order
    (the q_sample noise of the t2m batch is passed explicitly: the reference draws it with th.rand_like, :1334)."""
    terms = {}
    tmap = torch.tensor(sch.timestep_map, dtype=torch.long)

    def model(x, ti, feat):
        return OD.mdm_forward(w_front, x, tmap[ti], feat, enc_w=w_enc)

    mu = None
    if semantic_guidance:
        x_t = OS.q_sample(sch, x_start, t, noise_t2m, inp_mask_t2m)
        model_output = model(x_t, t, text_feat_t2m)
        mu = motion_encoder_forward(w_front, w_menc, model_output, frame_mask_t2m)
    skip = int(skip_steps / 1000 * 20) if use_ddim else skip_steps
    indices = list(range(sch.N - skip))[::-1]
    B = x_content.shape[0]
    img = tape.draw(x_content.shape)                                   # th.randn(*shape), :744
    img = OS.q_sample(sch, x_content, torch.full((B,), indices[0], dtype=torch.long), img, inp_mask)
    xs = []
    for i in indices:
        ti = torch.full((B,), i, dtype=torch.long)
        out = model(img.detach(), ti, text_feat_style)                 # x detached at entry of every *_with_grad step
        noise = tape.draw(img.shape)
        step = OS.ddim_sample if use_ddim else OS.p_sample
        nxt, x0 = step(sch, out, img, ti, noise, inp_mask, x_style, False)
        img = nxt.detach()
        xs.append(x0)                                                  # pred_xstart_in_graph=True
    num_step = len(xs)
    sample = torch.cat(xs, dim=0)
    target = x_style.expand(num_step, -1, -1, -1)
    m = frame_mask_style.view(-1, 1, 1, frame_mask_style.shape[-1]).float().expand(num_step, -1, -1, -1)
    terms["rot_mse"] = masked_l2(target, sample, m)
    if semantic_guidance:
        fn = text_feat_t2m / text_feat_t2m.norm(dim=-1, keepdim=True)
        mn = mu / mu.norm(dim=-1, keepdim=True)
        cos = torch.nn.functional.cosine_similarity(fn, mn, dim=1, eps=1e-6)
        terms["text_cosine"] = (1 - cos).mean()
        terms["loss"] = terms["rot_mse"].mean() + terms["text_cosine"] * Ls
    else:
        terms["loss"] = terms["rot_mse"].mean()
    terms["xstart"] = [x.detach() for x in xs]
    return terms


def adamw_step(p, g, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.AdamW update of one tensor, written out (in place on p, m, v)."""
    p.mul_(1 - lr * wd)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
