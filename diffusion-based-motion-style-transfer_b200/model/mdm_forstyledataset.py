"""Denoisers of the reference's ``model/mdm_forstyledataset.py`` as thin
``nn.Module`` shells around the hand-written sm_100a forward.

The modules keep the reference's constructor signatures, attribute names and
``state_dict`` layout (so reference checkpoints load with the same
``load_model_wo_clip`` / ``load_model_wo_moenc`` discipline,
``utils/model_util.py:9-23``), but their ``forward`` does not execute any torch
layer: parameters are handed to the C-ABI engine, which packs them (bf16 for
the tcgen05 path) and runs InputProcess -> 8 x TransformerEncoderLayer ->
OutputProcess as fused CUDA kernels.

Scope: ``arch='trans_enc'`` with a single Linear in/out projection
(``data_rep`` in rot6d / xyz / hml_vec) - the only configuration the reference's
scripts instantiate (``utils/model_util.py:108-167``).  Other arches raise.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn

from ..engine import Engine, TapeSlot, default_precision


# ----------------------------------------------------------------------------------
# parameter containers with the reference's state_dict key names
# ----------------------------------------------------------------------------------
class _SelfAttnParams(nn.Module):
    """Keys of nn.MultiheadAttention: in_proj_weight, in_proj_bias, out_proj.{weight,bias}."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _EncoderLayerParams(nn.Module):
    """Keys of nn.TransformerEncoderLayer (post-norm): self_attn, linear1, linear2, norm1, norm2."""

    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = _SelfAttnParams(d)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d, eps=1e-5)
        self.norm2 = nn.LayerNorm(d, eps=1e-5)

    def mst_tensors(self):
        return {
            "qkv_w": self.self_attn.in_proj_weight, "qkv_b": self.self_attn.in_proj_bias,
            "o_w": self.self_attn.out_proj.weight, "o_b": self.self_attn.out_proj.bias,
            "w1": self.linear1.weight, "b1": self.linear1.bias, "w2": self.linear2.weight, "b2": self.linear2.bias,
            "ln1_g": self.norm1.weight, "ln1_b": self.norm1.bias, "ln2_g": self.norm2.weight, "ln2_b": self.norm2.bias,
        }


class _EncoderParams(nn.Module):
    """Keys of nn.TransformerEncoder: layers.{i}.*"""

    def __init__(self, d, ff, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayerParams(d, ff) for _ in range(n_layers)])


class PositionalEncoding(nn.Module):
    """Sinusoidal table, persisted buffer 'pe' of shape [max_len, 1, d] (reference :387-404)."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout_p = dropout
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0).transpose(0, 1))


class TimestepEmbedder(nn.Module):
    """time_embed = Linear -> SiLU -> Linear applied to pe[t] (reference :408-422)."""

    def __init__(self, latent_dim, sequence_pos_encoder):
        super().__init__()
        self.latent_dim = latent_dim
        self.sequence_pos_encoder = sequence_pos_encoder
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.SiLU(), nn.Linear(latent_dim, latent_dim))


class InputProcess(nn.Module):
    def __init__(self, data_rep, input_feats, latent_dim):
        super().__init__()
        self.data_rep, self.input_feats, self.latent_dim = data_rep, input_feats, latent_dim
        self.poseEmbedding = nn.Linear(input_feats, latent_dim)


class OutputProcess(nn.Module):
    def __init__(self, data_rep, input_feats, latent_dim, njoints, nfeats):
        super().__init__()
        self.data_rep, self.input_feats, self.latent_dim = data_rep, input_feats, latent_dim
        self.njoints, self.nfeats = njoints, nfeats
        self.poseFinal = nn.Linear(latent_dim, input_feats)


class _NullRot2xyz:
    """Placeholder for the SMPL forward wrapper the reference instantiates but never calls on the
    hot path (mdm_forstyledataset.py:270); rendering is out of scope."""
    smpl_model = None


def _load_clip(clip_version):
    """Frozen CLIP text tower, if the ``clip`` package is importable (it is third-party and absent
    from the sandbox).  Without it callers must put ``y['text_feat']`` ([B, clip_dim]) in model_kwargs."""
    try:
        import clip  # type: ignore
    except Exception:
        return None
    clip_model, _ = clip.load(clip_version, device='cpu', jit=False)
    clip.model.convert_weights(clip_model)
    clip_model.eval()
    for p in clip_model.parameters():
        p.requires_grad = False
    if os.environ.get("MST_NATIVE_CLIP", "1") != "0":
        # scope row N1: the text side of the loaded model on the mst kernels (csrc/text.cu); same .encode_text()
        from .clip_text import CLIPTextTower
        return CLIPTextTower.from_clip(clip_model)
    return clip_model


# ----------------------------------------------------------------------------------
# autograd bridges: the forward records an activation tape in the C-ABI engine, the backward is the explicit
# CUDA backward (csrc/train.cu).  torch only routes the gradient tensors (plumbing).
# ----------------------------------------------------------------------------------
def _flat_layer_grads(params_by_layer, needs):
    """One zero arena for the gradients of every trainable encoder parameter + per-layer dicts of views."""
    total = sum(p.numel() for lp in params_by_layer for k, p in lp.items() if needs[id(p)])
    ref = next(p for lp in params_by_layer for p in lp.values())
    arena = torch.zeros(total, dtype=torch.float32, device=ref.device)
    off, out = 0, []
    for lp in params_by_layer:
        d = {}
        for k, p in lp.items():
            if needs[id(p)]:
                d[k] = arena[off:off + p.numel()].view(p.shape)
                off += p.numel()
            else:
                d[k] = None
        out.append(d)
    return out


class _DenoiserGradFn(torch.autograd.Function):
    """x0 prediction with a gradient path to the encoder-layer parameters (and to x when it requires grad).

    Two modes.  Pooled (``native.mst_tape_pool`` active - the trainer switches it on): the forward runs on a persistent
    TapeSlot and the backward accumulates straight into the parameters' existing ``.grad`` buffers, so every pointer is
    the same from one training step to the next and both directions replay as CUDA graphs.  Plain: fresh tape, fresh
    gradient tensors handed back to autograd."""

    @staticmethod
    def forward(ctx, native, x, temb, text_emb, uncond, *params):
        eng = native.__dict__.pop("_mst_eng_hint", None) or native.mst_engine(x.device, precision=native.mst_train_prec())
        B = x.shape[0]
        slot, k = native.__dict__.pop("_mst_slot_hint", None) or native._mst_tape_acquire(eng, B, x.shape[-1], text_emb is not None)
        ctx.native, ctx.eng, ctx.slot, ctx.k = native, eng, slot, k
        p = ctx.drop_p = native.mst_dropout_p()
        key = native.mst_draw_dropout_key() if p > 0 else 0
        if slot is not None:
            r = slot.rows(k)
            slot.x[r].copy_(x)
            slot.temb[r].copy_(temb)
            if text_emb is not None:
                slot.text[r].copy_(text_emb)
            if p > 0:
                slot.seed[r].copy_(torch.arange(key, key + B, dtype=torch.int64), non_blocking=True)
            eng.forward_train(slot.x[r], slot.temb[r], slot.text[r] if text_emb is not None else None, uncond=uncond,
                              tape=slot.tape, out=slot.out[r], use_graph=True, dropout_p=p, dropout_seed=slot.seed[r],
                              tape_seqs=slot.tape_seqs, tape_seq_offset=k * B)
            ctx.tape, ctx.epoch = slot.tape, slot.epoch
            return slot.out[r].clone()
        ctx.drop_seed = torch.arange(key, key + B, dtype=torch.int64, device=x.device) if p > 0 else None
        out, tape = eng.forward_train(x, temb, text_emb, uncond=uncond, dropout_p=p, dropout_seed=ctx.drop_seed)
        ctx.tape = tape
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.tape is None:
            raise RuntimeError("the mst activation tape of this forward was already consumed by a backward pass "
                               "(retain_graph is not supported: run the forward again)")
        layers = ctx.native._mst_cached_tensors()[1]
        slot = ctx.slot
        if slot is not None:
            if slot.epoch != ctx.epoch:
                raise RuntimeError("this forward's pooled activation tape was recycled by a later training step "
                                   "(call backward() before the next zero_grad(), or disable the pool)")
            # the six single-sequence forwards of a finetune step share one slot and ask the same question about the
            # same ~100 parameters: walk them once per step (p.grad is a slow property), not once per backward call
            want = ctx.needs_input_grad[5:]
            cache = slot.__dict__.get("_bw_cache")
            probe = layers[0]["qkv_w"].grad  # the trainer may point .grad at a second arena for this backward
            arena = None if probe is None else probe.data_ptr()
            if (cache is not None and cache[0] == slot.epoch and cache[1] is layers and cache[2] == want
                    and cache[6] == arena):
                direct, grads, n_flat = cache[3], cache[4], cache[5]
            else:
                flat = [p for lp in layers for p in lp.values()]
                needs = {id(p): bool(want[i]) for i, p in enumerate(flat)}
                direct = all((not needs[id(p)]) or (p.grad is not None and p.grad.is_contiguous() and
                                                     p.grad.dtype == torch.float32) for p in flat)
                grads = ([{k_: (p.grad if needs[id(p)] else None) for k_, p in lp.items()} for lp in layers]
                         if direct else None)
                n_flat = len(flat)
                slot._bw_cache = (slot.epoch, layers, want, direct, grads, n_flat, arena)
            if direct and not ctx.needs_input_grad[1]:
                # accumulate straight into .grad (what autograd's AccumulateGrad would do with a returned tensor); the
                # kernels run once every forward recorded on this tape has reported its output gradient
                slot.stage_backward(ctx.k, d_out, ctx.drop_p, grads)
                ctx.tape = None
                return (None,) * (5 + n_flat)
        flat = [p for lp in layers for p in lp.values()]
        needs = {id(p): bool(ctx.needs_input_grad[5 + i]) for i, p in enumerate(flat)}
        if slot is not None:
            # a gradient w.r.t. x (or parameters without .grad buffers) was requested: plain backward on the pooled tape
            B = d_out.shape[0]
            grads = _flat_layer_grads(layers, needs)
            d_x = ctx.eng.backward(d_out.float().contiguous(), slot.tape, grads, want_dx=bool(ctx.needs_input_grad[1]),
                                   dropout_p=ctx.drop_p, dropout_seed=slot.seed[slot.rows(ctx.k)],
                                   tape_seqs=slot.tape_seqs, tape_seq_offset=ctx.k * B)
            slot.done.add(ctx.k)
            if len(slot.pending) + len(slot.done) == slot.used:
                slot.flush()
            ctx.tape = None
            return (None, d_x, None, None, None) + tuple(g for lg in grads for g in lg.values())
        grads = _flat_layer_grads(layers, needs)
        d_x = ctx.eng.backward(d_out.float().contiguous(), ctx.tape, grads, want_dx=bool(ctx.needs_input_grad[1]),
                               dropout_p=ctx.drop_p, dropout_seed=ctx.drop_seed)
        ctx.tape = None
        return (None, d_x, None, None, None) + tuple(g for lg in grads for g in lg.values())


class _DenoiserPooledFn(torch.autograd.Function):
    """The trainer's fast lane of ``_DenoiserGradFn``: pooled tape, gradients accumulated straight into the parameters'
    arena ``.grad`` buffers, no gradient w.r.t. x.  The ~100 encoder parameters are NOT autograd inputs here (one anchor
    parameter keeps the node alive): with them, every apply() and every backward spends ~0.1 ms of host time wrapping,
    unwrapping and validating tensors whose gradients are written by the kernels anyway."""

    @staticmethod
    def forward(ctx, native, x, temb, text_emb, uncond, anchor, where):
        eng, slot, k = where
        B = x.shape[0]
        ctx.native, ctx.slot, ctx.k = native, slot, k
        p = ctx.drop_p = native.mst_dropout_p()
        key = native.mst_draw_dropout_key() if p > 0 else 0
        r = slot.rows(k)
        slot.x[r].copy_(x)
        slot.temb[r].copy_(temb)
        if text_emb is not None:
            slot.text[r].copy_(text_emb)
        if p > 0:
            slot.seed[r].copy_(torch.arange(key, key + B, dtype=torch.int64), non_blocking=True)
        eng.forward_train(slot.x[r], slot.temb[r], slot.text[r] if text_emb is not None else None, uncond=uncond,
                          tape=slot.tape, out=slot.out[r], use_graph=True, dropout_p=p, dropout_seed=slot.seed[r],
                          tape_seqs=slot.tape_seqs, tape_seq_offset=k * B)
        ctx.epoch, ctx.live = slot.epoch, True
        return slot.out[r].clone()

    @staticmethod
    def backward(ctx, d_out):
        if not ctx.live:
            raise RuntimeError("the mst activation tape of this forward was already consumed by a backward pass "
                               "(retain_graph is not supported: run the forward again)")
        slot = ctx.slot
        if slot.epoch != ctx.epoch:
            raise RuntimeError("this forward's pooled activation tape was recycled by a later training step "
                               "(call backward() before the next zero_grad(), or disable the pool)")
        layers = ctx.native._mst_cached_tensors()[1]
        cache = slot.__dict__.get("_bw_cache")
        probe = layers[0]["qkv_w"].grad  # the trainer may point .grad at a second arena for this backward
        arena = None if probe is None else probe.data_ptr()
        if cache is not None and cache[0] == slot.epoch and cache[1] is layers and cache[2] == "pooled" and cache[6] == arena:
            grads = cache[4]
        else:
            for lp in layers:
                for p in lp.values():
                    if p.requires_grad and (p.grad is None or not p.grad.is_contiguous() or p.grad.dtype != torch.float32):
                        raise RuntimeError("mst_direct_grads promises an fp32 .grad buffer on every trainable encoder "
                                           "parameter (MixedPrecisionTrainer keeps them in its arena)")
            grads = [{k_: (p.grad if p.requires_grad else None) for k_, p in lp.items()} for lp in layers]
            slot._bw_cache = (slot.epoch, layers, "pooled", True, grads, 0, arena)
        slot.stage_backward(ctx.k, d_out, ctx.drop_p, grads)
        ctx.live = False
        return (None,) * 7


class _MencSlot:
    """Persistent buffers of one MotionEncoder forward / backward (trainer mode: CUDA-graph replay)."""

    def __init__(self, eng, shape):
        dev, B, T = eng.device, shape[0], shape[-1]
        tape_bytes, _ = eng.train_sizes(B, T + 2)
        self.tape = torch.empty(tape_bytes, dtype=torch.uint8, device=dev)
        self.x = torch.empty(shape, dtype=torch.float32, device=dev)
        self.key_valid = torch.empty(B, T + 2, dtype=torch.uint8, device=dev)
        self.mu = torch.empty(B, eng.d_model, dtype=torch.float32, device=dev)
        self.d_mu = torch.empty(B, eng.d_model, dtype=torch.float32, device=dev)
        self.d_x = torch.empty(shape, dtype=torch.float32, device=dev)
        self.seed = torch.zeros(B, dtype=torch.int64, device=dev)
        self.used = False


class _MotionEncoderGradFn(torch.autograd.Function):
    """mu of MotionEncoder.forward with a gradient path to its input motion only (its parameters are frozen)."""

    @staticmethod
    def forward(ctx, enc, x, key_valid):
        eng = enc.mst_engine(x.device, precision=enc.mst_train_prec())
        p = ctx.drop_p = enc.mst_dropout_p()
        key = enc.mst_draw_dropout_key() if p > 0 else 0
        mq, sq = enc.muQuery.detach().reshape(-1).contiguous(), enc.sigmaQuery.detach().reshape(-1).contiguous()
        slot = ctx.slot = enc._mst_menc_acquire(eng, tuple(x.shape))
        ctx.eng, ctx.shape = eng, tuple(x.shape)
        if slot is not None:
            slot.x.copy_(x)
            slot.key_valid.copy_(key_valid)
            if p > 0:
                slot.seed.copy_(torch.arange(key, key + x.shape[0], dtype=torch.int64), non_blocking=True)
            ctx.queries = (mq, sq)   # keep the operand tensors alive: the captured graph reads their addresses
            eng.motion_encoder_forward(slot.x, slot.key_valid, mq, sq, dropout_p=p, dropout_seed=slot.seed,
                                       tape=slot.tape, mu=slot.mu, use_graph=True)
            ctx.tape, ctx.drop_seed = slot.tape, slot.seed
            return slot.mu.clone()
        ctx.drop_seed = torch.arange(key, key + x.shape[0], dtype=torch.int64, device=x.device) if p > 0 else None
        mu, tape = eng.motion_encoder_forward(x, key_valid, mq, sq, dropout_p=p, dropout_seed=ctx.drop_seed)
        ctx.tape = tape
        return mu

    @staticmethod
    def backward(ctx, d_mu):
        if ctx.tape is None:
            raise RuntimeError("the mst activation tape of this forward was already consumed by a backward pass")
        d_x = None
        if ctx.needs_input_grad[1]:
            slot = ctx.slot
            if slot is not None:
                slot.d_mu.copy_(d_mu)
                ctx.eng.motion_encoder_backward(slot.d_mu, slot.tape, ctx.shape, dropout_p=ctx.drop_p,
                                                dropout_seed=ctx.drop_seed, d_x=slot.d_x, use_graph=True)
                d_x = slot.d_x.clone()
            else:
                d_x = ctx.eng.motion_encoder_backward(d_mu.float().contiguous(), ctx.tape, ctx.shape, dropout_p=ctx.drop_p,
                                                      dropout_seed=ctx.drop_seed)
        ctx.tape = None
        return None, d_x, None


# ----------------------------------------------------------------------------------
class NativeDenoiser(nn.Module):
    """Common engine plumbing of MDM and StyleDiffusion."""

    mst_precision = None  # None -> MST_PRECISION env (bf16 default) ; or 'fp32' / 'bf16'
    # precision of the linear layers on the training / differentiable-sampling path: 'fp32' (SIMT, parity mode, <=2e-4
    # of the reference's gradients) or 'bf16' (tcgen05, bf16 operands + fp32 accumulation, <=3e-2).
    # None -> MST_TRAIN_PRECISION env, else the same default as the sampler.
    mst_train_precision = None

    def mst_train_prec(self) -> str:
        p = self.mst_train_precision or os.environ.get("MST_TRAIN_PRECISION") or self.mst_precision or default_precision()
        if p not in ("bf16", "fp32"):
            raise ValueError(f"training precision must be 'bf16' or 'fp32', got {p!r}")
        return p

    # subclasses provide these views -------------------------------------------------
    def _mst_front(self):
        """module that owns input_process / sequence_pos_encoder / embed_timestep / embed_text / output_process"""
        raise NotImplementedError

    def _mst_encoder(self):
        """the _EncoderParams whose layers run between in- and out-projection"""
        raise NotImplementedError

    # ---------------------------------------------------------------------------------
    def _mst_all_tensors(self):
        f = self._mst_front()
        top = {
            "in_w": f.input_process.poseEmbedding.weight, "in_b": f.input_process.poseEmbedding.bias,
            "pe": f.sequence_pos_encoder.pe,
            "t_w1": f.embed_timestep.time_embed[0].weight, "t_b1": f.embed_timestep.time_embed[0].bias,
            "t_w2": f.embed_timestep.time_embed[2].weight, "t_b2": f.embed_timestep.time_embed[2].bias,
            "txt_w": f.embed_text.weight if hasattr(f, "embed_text") else None,
            "txt_b": f.embed_text.bias if hasattr(f, "embed_text") else None,
            "out_w": f.output_process.poseFinal.weight, "out_b": f.output_process.poseFinal.bias,
        }
        layers = [l.mst_tensors() for l in self._mst_encoder().layers]
        return top, layers

    def __setattr__(self, name, value):
        if isinstance(value, (nn.Module, nn.Parameter)):  # a replaced sub-module / parameter invalidates the cached walk
            self.__dict__.pop("_mst_tensor_cache", None)
            self.__dict__.pop("_mst_grad_cache", None)
        super().__setattr__(name, value)

    def _mst_cached_tensors(self):
        """(top, layers, flat list) of the weight tensors.  Walking the module tree costs ~0.2 ms and every forward /
        embedding call asks for the engine (18 times per finetune step), so the walk is cached; the per-call check is
        (data_ptr, _version) of the cached tensors, which sees in-place updates, load_state_dict and .data swaps.  The
        cache is dropped by .to()/.cuda()/.float() (``_apply``), ``load_state_dict``, ``mst_weights_changed`` and when a
        sub-module or Parameter is assigned on a module of this tree through ``add_module`` / attribute assignment
        on the denoiser itself; after deeper module surgery call ``mst_weights_changed()``."""
        c = self.__dict__.get("_mst_tensor_cache")
        if c is None:
            top, layers = self._mst_all_tensors()
            flat = [t for d in [top] + layers for t in d.values() if t is not None]
            c = (top, layers, flat)
            self.__dict__["_mst_tensor_cache"] = c
        return c

    def _mst_drop_tensor_cache(self):
        self.__dict__.pop("_mst_tensor_cache", None)
        self.__dict__.pop("_mst_grad_cache", None)

    def _apply(self, fn, *args, **kwargs):
        self._mst_drop_tensor_cache()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._mst_drop_tensor_cache()
        return super().load_state_dict(*args, **kwargs)

    def requires_grad_(self, requires_grad=True):
        self._mst_drop_tensor_cache()
        return super().requires_grad_(requires_grad)

    def _mst_signature(self, top, layers):
        sig = []
        for d in [top] + layers:
            for t in d.values():
                if t is not None:
                    sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def mst_ready(self, x) -> bool:
        """True when ``x`` and the parameters live on the same CUDA device (the only supported case)."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            return False
        p = self._mst_front().input_process.poseEmbedding.weight
        if not p.is_cuda or p.device != x.device:
            raise RuntimeError(f"model parameters are on {p.device} but the input is on {x.device}: move the model "
                               "to the CUDA device first (there is no CPU fallback)")
        return True

    # -- dropout of the training forward ---------------------------------------------------------------------
    # nn.TransformerEncoderLayer(dropout=p) and PositionalEncoding(dropout=p) are active in the reference as soon as
    # model.train() was called (train/finetune_style_diffusion.py:256).  None -> follow self.training like torch does;
    # a float forces that probability (0.0 = deterministic gradients, what the parity fixtures pin).
    mst_train_dropout = None

    def mst_dropout_p(self) -> float:
        if self.mst_train_dropout is not None:
            return float(self.mst_train_dropout)
        return float(getattr(self, "dropout", 0.0)) if self.training else 0.0

    @staticmethod
    def mst_draw_dropout_key() -> int:
        """Philox key of one forward's masks, drawn from torch's CPU generator (so torch.manual_seed controls it)."""
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    # -- pooled activation tapes (CUDA-graph replay of the training forward / backward) ------------------------
    mst_tape_pool = False   # switched on by the trainer (MixedPrecisionTrainer); MST_TRAIN_GRAPH=0 disables replay

    def mst_tape_reset(self):
        """Start of a training step: pending backward work is flushed and every pooled slot may be handed out again, in
        the same order as last step."""
        for ent in self.__dict__.get("_mst_tape_slots", {}).values():
            for slot in ent[0]:
                slot.reset()
            ent[1] = 0
        for slot in self.__dict__.get("_mst_menc_slots", {}).values():
            slot.used = False

    def _mst_menc_acquire(self, eng, shape):
        """Persistent MotionEncoder buffers for this input shape (trainer mode, one forward per step), else None."""
        if not self.mst_tape_pool:
            return None
        slots = self.__dict__.setdefault("_mst_menc_slots", {})
        slot = slots.get((id(eng), shape))
        if slot is None:
            if len(slots) >= 4:
                return None
            slot = slots[(id(eng), shape)] = _MencSlot(eng, shape)
        if slot.used:
            return None      # a second forward of this shape in the same step: plain path
        slot.used = True
        return slot

    def mst_flush_backward(self):
        """Run the batched backward passes that are still waiting for a sibling forward's gradient (a forward whose
        output never reached the loss).  The trainer calls this after loss.backward() and before reading gradients."""
        for ent in self.__dict__.get("_mst_tape_slots", {}).values():
            for slot in ent[0]:
                slot.flush()

    def _mst_tape_acquire(self, eng, B, T, has_text):
        """(slot, call index) of a pooled tape, or (None, 0).  Single-sequence forwards (the differentiable sampling
        steps) share one tape of 8 calls so that their backward passes run as one batch."""
        if not self.mst_tape_pool:
            return None, 0
        pools = self.__dict__.setdefault("_mst_tape_slots", {})
        ent = pools.setdefault((id(eng), B, T, has_text), [[], 0])
        slots, cur = ent
        capacity = 8 if B == 1 else 1
        if cur < len(slots) and slots[cur].used == capacity:
            cur = ent[1] = cur + 1
        if cur >= 4:        # more live forwards of one shape than a finetune step ever has: plain path
            return None, 0
        if cur == len(slots):
            slots.append(TapeSlot(eng, B, T, has_text, capacity))
        slot = slots[cur]
        k = slot.used
        slot.used += 1
        return slot, k

    def mst_weights_changed(self, structure=True):
        """Call after parameters were updated behind torch's back (the fused optimizer writes through raw
        pointers, which does not bump the tensors' version counters): forces a re-pack on next use.
        ``structure=False`` (the optimizer's per-step call) keeps the cached walk of the module tree."""
        if structure:
            self._mst_drop_tensor_cache()
        for ent in self.__dict__.get("_mst_engines", {}).values():
            ent[1] = None

    def mst_engine(self, device, precision=None) -> Engine:
        """Engine for this module on ``device``; weights are (re)packed whenever a parameter changed."""
        prec = precision or self.mst_precision or default_precision()
        key = (str(device), prec)
        cache = self.__dict__.setdefault("_mst_engines", {})
        top, layers, flat = self._mst_cached_tensors()
        sig = tuple([(t.data_ptr(), t._version) for t in flat])
        ent = cache.get(key)
        if ent is None:
            f = self._mst_front()
            eng = Engine(n_feats=self.input_feats, d_model=self.latent_dim, n_heads=self.num_heads,
                         d_ff=self.ff_size, n_layers=len(layers), clip_dim=self.clip_dim,
                         pe_len=f.sequence_pos_encoder.pe.shape[0], precision=prec, device=device)
            ent = [eng, None]
            cache[key] = ent
        if ent[1] != sig:
            top = dict(top)
            top["pe"] = top["pe"].reshape(top["pe"].shape[0], -1)
            ent[0].load_weights(top, layers)
            ent[1] = sig
        return ent[0]

    # -- conditioning ------------------------------------------------------------------
    def encode_text(self, raw_text):
        """CLIP text features [B, clip_dim] fp32 (reference :298-313).  CLIP itself is third-party and
        outside the hot path; when it is not installed, pass ``y['text_feat']`` instead."""
        f = self._mst_front()
        clip_model = getattr(f, "clip_model", None)
        if clip_model is None:
            raise RuntimeError("no CLIP text encoder is attached to this model (the `clip` package is not installed); "
                               "provide precomputed features as y['text_feat'] with shape [B, clip_dim]")
        tokenize = getattr(f, "mst_tokenize", None)  # hook: any callable with clip.tokenize's signature
        if tokenize is None:
            try:
                import clip  # type: ignore
                tokenize = clip.tokenize
            except ImportError:  # the native BPE tokenizer (needs the vocabulary file: MST_CLIP_BPE or attach_tokenizer)
                from .clip_tokenizer import SimpleTokenizer
                tokenize = f.mst_tokenize = SimpleTokenizer().tokenize
        device = next(self.parameters()).device
        max_text_len = 20 if self.dataset in ['humanml', 'kit'] else None
        if max_text_len is not None:
            default_context_length = 77
            context_length = max_text_len + 2
            texts = tokenize(raw_text, context_length=context_length, truncate=True).to(device)
            zero_pad = torch.zeros([texts.shape[0], default_context_length - context_length], dtype=texts.dtype,
                                   device=texts.device)
            texts = torch.cat([texts, zero_pad], dim=1)
        else:
            texts = tokenize(raw_text, truncate=True).to(device)
        return clip_model.encode_text(texts).float()

    def mask_cond(self, cond, force_mask=False):
        bs, d = cond.shape
        if force_mask:
            return torch.zeros_like(cond)
        elif self.training and self.cond_mask_prob > 0.:
            mask = torch.bernoulli(torch.ones(bs, device=cond.device) * self.cond_mask_prob).view(bs, 1)
            return cond * (1. - mask)
        return cond

    def text_features(self, y, device):
        """[B, clip_dim] fp32 CUDA: y['text_feat'] when given, else CLIP(y['text'])."""
        if 'text' not in self.cond_mode:
            return None
        feat = y.get('text_feat', None)
        if feat is None:
            feat = self.encode_text_cached(y['text'], device)
        return self.mask_cond(feat.to(device).float()).contiguous()

    def encode_text_cached(self, raw_text, device):
        """CLIP features of a list of captions through a per-model cache keyed by the caption string (scope row N1, the
        "cached-feature API" half): the frozen text tower is deterministic, so a caption is encoded once per process -
        the reference re-encodes the same strings at every denoising step and twice under CFG
        (model/mdm_forstyledataset.py:326, model/cfg_sampler.py:36-43).  ``mst_text_cache_clear()`` drops the cache."""
        cache = self.__dict__.setdefault("_mst_text_cache", {})
        missing = [t for t in dict.fromkeys(raw_text) if (t, str(device)) not in cache]
        if missing:
            feats = self.encode_text(missing).detach().to(device).float()
            for t, f in zip(missing, feats):
                cache[(t, str(device))] = f.clone()
            while len(cache) > 65536:  # bound the memory: 512 floats per caption
                cache.pop(next(iter(cache)))
        return torch.stack([cache[(t, str(device))] for t in raw_text])

    def mst_text_cache_clear(self):
        self.__dict__.pop("_mst_text_cache", None)

    def text_embedding(self, y, device, precision=None, eng=None):
        """embed_text(mask_cond(clip(text))) [B, d]; computed once per trajectory by the sampler."""
        feat = self.text_features(y, device)
        if feat is None:
            return None
        return (eng or self.mst_engine(device, precision)).text_embed(feat)

    @staticmethod
    def compact_mask(mask):
        """[B,F,1,T] mask that does not vary over batch and time -> [F] vector (all named masks of
        data_loaders/*_utils.py except in_between/prefix).  Saves one state-sized read per step."""
        if mask is None or mask.dim() != 4:
            return mask
        col = mask[:1, :, :, :1]
        if bool((mask == col).all()):
            return col.reshape(-1).contiguous()
        row = mask[:1]
        if bool((mask == row).all()):
            return row.reshape(mask.shape[1] * mask.shape[2], mask.shape[3]).contiguous()
        return mask

    # -- forward -----------------------------------------------------------------------
    def forward(self, x, timesteps, y=None):
        """x: [B, njoints, nfeats, T] (x_t); timesteps: [B] int; y: dict with 'text' or 'text_feat',
        optional 'uncond'.  Returns the x_0 prediction [B, njoints, nfeats, T] (reference :315-364)."""
        if not self.mst_ready(x):
            raise RuntimeError("the mst denoiser runs on CUDA tensors only (no CPU fallback)")
        y = y if y is not None else {}
        force_mask = bool(y.get('uncond', False))
        xc = x.float().contiguous()
        if self._mst_wants_grad(xc):
            # training / differentiable-sampling path (fp32 engine with an activation tape)
            enc_params = [p for lp in self._mst_cached_tensors()[1] for p in lp.values()]
            with torch.no_grad():
                eng = self.mst_engine(x.device, precision=self.mst_train_prec())
                temb = eng.time_embed(timesteps)
                text_emb = None if force_mask else self.text_embedding(y, x.device, eng=eng)
            if 'text' in self.cond_mode and text_emb is None and not force_mask:
                raise RuntimeError("text-conditioned model called without text")
            uncond = force_mask or text_emb is None
            if self.mst_tape_pool and self.__dict__.get("mst_direct_grads") and not xc.requires_grad:
                where = self._mst_tape_acquire(eng, xc.shape[0], xc.shape[-1], text_emb is not None)
                anchor = next((p for p in enc_params if p.requires_grad), None)
                if where[0] is not None and anchor is not None:
                    return _DenoiserPooledFn.apply(self, xc, temb, text_emb, uncond, anchor, (eng,) + tuple(where))
                self.__dict__["_mst_slot_hint"] = where
            self.__dict__["_mst_eng_hint"] = eng  # the bridge below would otherwise fingerprint the weights a third time
            return _DenoiserGradFn.apply(self, xc, temb, text_emb, uncond, *enc_params)
        eng = self.mst_engine(x.device)
        temb = eng.time_embed(timesteps)
        text_emb = None if force_mask else self.text_embedding(y, x.device)
        if 'text' in self.cond_mode and text_emb is None and not force_mask:
            raise RuntimeError("text-conditioned model called without text")
        return eng.forward(xc, temb, text_emb, cfg=False, uncond=force_mask or text_emb is None)

    def _mst_wants_grad(self, x) -> bool:
        """True when autograd is recording and something upstream of the output needs a gradient.  Only the encoder
        stack is differentiable w.r.t. its parameters (the reference finetunes StyleDiffusion.seqTransEncoder with
        every other module frozen, mdm_forstyledataset.py:562-567); trainable projections / embedders raise."""
        if not torch.is_grad_enabled():
            return False
        gc = self.__dict__.get("_mst_grad_cache")
        if gc is None:  # parameter lists of the encoder stack / the front end (walked once; the flags are read per call)
            f = self._mst_front()
            front = [f.input_process, f.output_process, f.embed_timestep] + ([f.embed_text] if hasattr(f, "embed_text") else [])
            gc = ([p for l in self._mst_encoder().layers for p in l.parameters()], [p for m in front for p in m.parameters()])
            self.__dict__["_mst_grad_cache"] = gc
        enc = any([p.requires_grad for p in gc[0]])
        if not (enc or x.requires_grad):
            return False
        if any([p.requires_grad for p in gc[1]]):
            # e.g. a freshly constructed MDM evaluated without torch.no_grad(): run the inference kernels; the result
            # carries no grad_fn, so an attempted backward() fails loudly in torch instead of training half a model
            if not getattr(self, "_mst_warned_front_grad", False):
                import warnings
                warnings.warn("mst denoiser called with autograd enabled while its in/out projections or time/text "
                              "embedders require grad: gradients are only built for the encoder stack (the reference's "
                              "finetune path, StyleDiffusion with a frozen front end); this call is NOT differentiable. "
                              "Freeze those modules to train, or wrap inference in torch.no_grad().")
                self._mst_warned_front_grad = True
            return False
        return True

    def forward_cfg(self, x, timesteps, y):
        """Both passes of ClassifierFreeSampleModel.forward batched: returns (out_cond, out_uncond)."""
        eng = self.mst_engine(x.device)
        xc = x.float().contiguous()
        temb = eng.time_embed(timesteps)
        text_emb = self.text_embedding(y, x.device)
        if text_emb is None:
            raise RuntimeError("classifier-free guidance needs a text-conditioned model")
        return eng.forward(xc, temb, text_emb, cfg=True)


def _common_init(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot, latent_dim,
                 ff_size, num_layers, num_heads, dropout, ablation, activation, legacy, data_rep, dataset, clip_dim,
                 arch, emb_trans_dec, clip_version, kargs):
    self.legacy = legacy
    self.modeltype = modeltype
    self.njoints = njoints
    self.nfeats = nfeats
    self.num_actions = num_actions
    self.data_rep = data_rep
    self.dataset = dataset
    self.pose_rep = pose_rep
    self.glob = glob
    self.glob_rot = glob_rot
    self.translation = translation
    self.latent_dim = latent_dim
    self.ff_size = ff_size
    self.num_layers = num_layers
    self.num_heads = num_heads
    self.dropout = dropout
    self.ablation = ablation
    self.activation = activation
    self.clip_dim = clip_dim
    self.action_emb = kargs.get('action_emb', None)
    self.input_feats = self.njoints * self.nfeats
    self.normalize_output = kargs.get('normalize_encoder_output', False)
    self.cond_mode = kargs.get('cond_mode', 'no_cond')
    self.cond_mask_prob = kargs.get('cond_mask_prob', 0.)
    self.arch = arch
    self.gru_emb_dim = self.latent_dim if self.arch == 'gru' else 0
    self.emb_trans_dec = emb_trans_dec
    self.clip_version = clip_version
    if arch != 'trans_enc':
        raise NotImplementedError(f"arch={arch!r}: only 'trans_enc' is on the reference's hot path "
                                  "(sample/demo_style_transfer.py and train/finetune_style_diffusion.py)")
    if data_rep not in ('rot6d', 'xyz', 'hml_vec'):
        raise NotImplementedError(f"data_rep={data_rep!r}: the 'rot_vel' two-projection variant is not built")
    if activation != "gelu":
        raise NotImplementedError("only the GELU feed-forward the reference configures is built")
    if 'action' in self.cond_mode:
        raise NotImplementedError("action conditioning is not used by the reference's scripts and is not built")


class MDM(NativeDenoiser):
    """Reference ``MDM`` (mdm_forstyledataset.py:183-384): text-conditioned transformer-encoder denoiser."""

    def __init__(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                 latent_dim=256, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1,
                 ablation=None, activation="gelu", legacy=False, data_rep='rot6d', dataset='amass', clip_dim=512,
                 arch='trans_enc', emb_trans_dec=False, clip_version=None, **kargs):
        super().__init__()
        _common_init(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot, latent_dim,
                     ff_size, num_layers, num_heads, dropout, ablation, activation, legacy, data_rep, dataset,
                     clip_dim, arch, emb_trans_dec, clip_version, kargs)
        self.input_process = InputProcess(self.data_rep, self.input_feats + self.gru_emb_dim, self.latent_dim)
        self.sequence_pos_encoder = PositionalEncoding(self.latent_dim, self.dropout)
        self.seqTransEncoder = _EncoderParams(self.latent_dim, self.ff_size, self.num_layers)
        self.embed_timestep = TimestepEmbedder(self.latent_dim, self.sequence_pos_encoder)
        if self.cond_mode != 'no_cond' and 'text' in self.cond_mode:
            self.embed_text = nn.Linear(self.clip_dim, self.latent_dim)
            clip_model = _load_clip(clip_version) if kargs.get('load_clip', True) else None
            if clip_model is not None:
                self.clip_model = clip_model
        self.output_process = OutputProcess(self.data_rep, self.input_feats, self.latent_dim, self.njoints, self.nfeats)
        self.rot2xyz = _NullRot2xyz()

    def _mst_front(self):
        return self

    def _mst_encoder(self):
        return self.seqTransEncoder

    def parameters_wo_clip(self):
        return [p for name, p in self.named_parameters() if not name.startswith('clip_model.')]


class MotionEncoder(NativeDenoiser):
    """The reference's frozen semantic discriminator (mdm_forstyledataset.py:11-124): two learned query tokens
    (mu, sigma) in front of InputProcess(x), its own encoder stack with a key-padding mask, mu = output token 0.
    It also carries ``mdm_model``, whose projections and embedders ``StyleDiffusion`` borrows."""

    def __init__(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                 latent_dim=256, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1,
                 ablation=None, activation="gelu", legacy=False, data_rep='rot6d', dataset='amass', clip_dim=512,
                 arch='trans_enc', emb_trans_dec=False, clip_version=None, **kargs):
        super().__init__()
        self.latent_dim, self.ff_size, self.num_layers, self.num_heads = latent_dim, ff_size, num_layers, num_heads
        self.njoints, self.nfeats = njoints, nfeats
        self.input_feats = njoints * nfeats
        self.clip_dim = clip_dim
        self.dropout = dropout
        self.cond_mode = kargs.get('cond_mode', 'no_cond')
        self.dataset = dataset
        self.cond_mask_prob = kargs.get('cond_mask_prob', 0.)
        self.muQuery = nn.Parameter(torch.randn(1, latent_dim))
        self.sigmaQuery = nn.Parameter(torch.randn(1, latent_dim))
        self.seqTransEncoder = _EncoderParams(latent_dim, ff_size, num_layers)
        self.mdm_model = MDM(modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                             latent_dim, ff_size, num_layers, num_heads, dropout, ablation, activation, legacy,
                             data_rep, dataset, clip_dim, arch, emb_trans_dec, clip_version, **kargs)
        mdm_path = kargs.get("mdm_path", "")
        if mdm_path:
            print("load mdm_model from checkpoint {}".format(mdm_path))
            self.load_model_wo_clip(self.mdm_model, torch.load(mdm_path, map_location='cpu'))
        self.mdm_model.eval()
        for p in self.mdm_model.parameters():
            p.requires_grad = False

    def parameters_wo_clip(self):
        return [p for name, p in self.named_parameters() if not name.startswith('mdm_model.')]

    @staticmethod
    def load_model_wo_clip(model, state_dict):
        missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
        assert len(unexpected_keys) == 0
        assert all([k.startswith('clip_model.') for k in missing_keys])

    def _mst_front(self):
        return self.mdm_model

    def _mst_encoder(self):
        return self.seqTransEncoder

    def mask_cond(self, cond, force_mask=False):
        return cond  # reference :126-127

    def encode_text(self, raw_text):
        return self.mdm_model.encode_text(raw_text)

    def forward(self, x, y=None):
        """x [B, njoints, nfeats, T] -> (mu [B, d], CLIP text features [B, clip_dim] or None)  (reference :89-124).
        Differentiable w.r.t. x (the finetune loss back-propagates the cosine term into the denoiser's output)."""
        if not self.mst_ready(x):
            raise RuntimeError("the mst MotionEncoder runs on CUDA tensors only (no CPU fallback)")
        bs, nframes = x.shape[0], x.shape[-1]
        enc_text = None
        if y is not None:
            mask = y.get("mask").to(x.device).squeeze(1).squeeze(1).bool()
            if y.get('text_feat', None) is not None:
                enc_text = y['text_feat'].to(x.device).float()
            elif y.get('text', None) is not None:
                enc_text = self.mdm_model.encode_text(y['text'])
        else:
            mask = torch.ones((bs, nframes), dtype=torch.bool, device=x.device)
        key_valid = torch.cat((torch.ones((bs, 2), dtype=torch.bool, device=x.device), mask), dim=1).to(torch.uint8)
        mu = _MotionEncoderGradFn.apply(self, x.float().contiguous(), key_valid.contiguous())
        return mu, enc_text


class StyleDiffusion(NativeDenoiser):
    """Reference ``StyleDiffusion`` (mdm_forstyledataset.py:494-625): its own trainable ``seqTransEncoder``
    between the frozen ``motion_enc.mdm_model``'s projections and embedders."""

    def __init__(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                 latent_dim=256, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1,
                 ablation=None, activation="gelu", legacy=False, data_rep='rot6d', dataset='amass', clip_dim=512,
                 arch='trans_enc', emb_trans_dec=False, clip_version=None, **kargs):
        super().__init__()
        _common_init(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot, latent_dim,
                     ff_size, num_layers, num_heads, dropout, ablation, activation, legacy, data_rep, dataset,
                     clip_dim, arch, emb_trans_dec, clip_version, kargs)
        self.kargs = kargs
        self.seqTransEncoder = _EncoderParams(self.latent_dim, self.ff_size, self.num_layers)
        self.motion_enc = MotionEncoder(modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                                        latent_dim, ff_size, num_layers, num_heads, dropout, ablation, activation,
                                        legacy, data_rep, dataset, clip_dim, arch, emb_trans_dec, clip_version, **kargs)
        self.load_motion_enc()

    def load_motion_enc(self):
        path = self.kargs.get("semantic_discriminator_path", "")
        if path:
            print("load motion_enc from checkpoint {}".format(path))
            self.load_model(self.motion_enc, torch.load(path, map_location='cpu'))
        self.motion_enc = self.motion_enc.eval()
        for p in self.motion_enc.parameters():
            p.requires_grad = False
        assert all([not para.requires_grad for para in self.motion_enc.parameters()])

    @staticmethod
    def load_model(model, state_dict):
        missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
        assert len(unexpected_keys) == 0
        assert all([k.startswith('mdm_model.') for k in missing_keys])

    def parameters_wo_enc(self):
        return [p for name, p in self.named_parameters() if not name.startswith('motion_enc.')]

    def _mst_front(self):
        return self.motion_enc.mdm_model

    def _mst_encoder(self):
        return self.seqTransEncoder

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name == "mst_train_precision" and "motion_enc" in self._modules:
            self.motion_enc.mst_train_precision = value  # the semantic discriminator runs in the same mode

    def encode_text(self, raw_text):
        return self.motion_enc.mdm_model.encode_text(raw_text)


# name the reference exports for the (unused) style/content-code variant; utils/model_util.py imports it
class DiffuseTrasnfer(StyleDiffusion):
    def __init__(self, *args, **kargs):
        raise NotImplementedError("DiffuseTrasnfer is defined but never instantiated by the reference's scripts "
                                  "(SURVEY section 2); use StyleDiffusion")
