"""Python handle on the C-ABI denoiser engine and the fused update kernels.

PyTorch is used only for device memory and streams; every device operation is
a call into libmst_b200.so.  All tensors handed to these wrappers must be
CUDA, contiguous and of the documented dtype - anything else raises (there is
no CPU path).
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from typing import Optional

import torch

from . import _lib as L


def default_precision() -> str:
    p = os.environ.get("MST_PRECISION", "bf16").lower()
    if p not in ("bf16", "fp32"):
        raise ValueError(f"MST_PRECISION must be 'bf16' or 'fp32', got {p!r}")
    return p


# torch.cuda.current_stream() / current_device() build Python objects through several layers (~20 us and ~5 us a call); the
# finetune step asks ~90 times.  The raw accessors below are what torch's own generated code (inductor) calls.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _current_device() -> int:
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream_ptr() -> int:
    if _raw_stream is not None:
        return _raw_stream(_current_device())
    return torch.cuda.current_stream().cuda_stream


def _engine_device(fn):
    """Run an Engine method with the engine's GPU current: kernels launch on the CUDA runtime's current device and on
    its current stream, and the reference picks its GPU with ``--device N`` without ever calling ``set_device`` - the
    model and its tensors may live on cuda:N while the process's current device is still 0."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if _current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


def _tensor_device(fn):
    """The same for the engine-less kernels: the first CUDA tensor among the arguments names the device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for v in list(args) + list(kwargs.values()):
            if isinstance(v, torch.Tensor) and v.is_cuda:
                dev = v.device
                break
            if isinstance(v, torch.device) and v.type == "cuda":
                dev = v
                break
        if dev is None or dev.index is None or _current_device() == dev.index:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32, name="tensor") -> Optional[int]:
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: must live on a CUDA device (the mst kernels have no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: must be contiguous")
    return t.data_ptr()


class Engine:
    """One denoiser (MDM / StyleDiffusion) instance on one GPU."""

    LAYER_KEYS = L._LAYER_FIELDS

    def __init__(self, n_feats: int, d_model: int = 512, n_heads: int = 4, d_ff: int = 1024, n_layers: int = 8,
                 clip_dim: int = 512, pe_len: int = 5000, precision: Optional[str] = None,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("mst Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.precision = precision or default_precision()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.desc = L.ModelDesc(n_feats, d_model, n_heads, d_ff, n_layers, clip_dim, pe_len,
                                L.PREC_BF16 if self.precision == "bf16" else L.PREC_FP32)
        h = C.c_void_p()
        L.check(self.lib.mst_engine_create(C.byref(self.desc), C.byref(h)), "mst_engine_create")
        self._h = h
        self._packed = None
        self._keep = None  # tensors whose storage the engine aliases
        self._ws = None
        self.n_feats, self.d_model = n_feats, d_model
        self.weights_generation = 0  # bumped by every load_weights: captured graphs bake weight pointers in

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.mst_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- weights -----------------------------------------------------------------
    def load_weights(self, top: dict, layers: list):
        """top: in_w,in_b,pe,t_w1,t_b1,t_w2,t_b2,txt_w,txt_b,out_w,out_b -> fp32 CUDA tensors (txt_* may be None);
        layers: list of dicts keyed by LAYER_KEYS."""
        w = L.Weights()
        keep = []
        with torch.cuda.device(self.device):
            for k in ("in_w", "in_b", "pe", "t_w1", "t_b1", "t_w2", "t_b2", "txt_w", "txt_b", "out_w", "out_b"):
                t = top.get(k)
                if t is not None:
                    t = t.detach()
                    if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                        t = t.to(self.device, torch.float32).contiguous()
                    keep.append(t)
                setattr(w, k, _ptr(t, name=k))
            if len(layers) != self.desc.n_layers:
                raise ValueError(f"expected {self.desc.n_layers} layers, got {len(layers)}")
            for i, lw in enumerate(layers):
                for k in self.LAYER_KEYS:
                    t = lw[k].detach()
                    if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                        t = t.to(self.device, torch.float32).contiguous()
                    keep.append(t)
                    setattr(w.layers[i], k, _ptr(t, name=f"layer{i}.{k}"))
            nbytes = C.c_size_t()
            L.check(self.lib.mst_engine_packed_weight_bytes(self._h, C.byref(nbytes)))
            if self._packed is None or self._packed.numel() < nbytes.value:
                self._packed = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            L.check(self.lib.mst_engine_load_weights(self._h, C.byref(w), self._packed.data_ptr(), nbytes.value,
                                                     _stream_ptr()), "mst_engine_load_weights")
        self._keep = keep
        self.weights_generation += 1

    # -- scratch -----------------------------------------------------------------
    def workspace_bytes(self, n_seqs: int, n_frames: int) -> int:
        nbytes = C.c_size_t()
        L.check(self.lib.mst_engine_workspace_bytes(self._h, n_seqs, n_frames, C.byref(nbytes)))
        return int(nbytes.value)

    @_engine_device
    def workspace(self, n_seqs: int, n_frames: int) -> torch.Tensor:
        """Engine-owned scratch for eager calls (grown on demand; captured graphs bring their own)."""
        nbytes = C.c_size_t()
        L.check(self.lib.mst_engine_workspace_bytes(self._h, n_seqs, n_frames, C.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        return self._ws

    # -- small embeddings ----------------------------------------------------------
    @_engine_device
    def time_embed(self, t: torch.Tensor) -> torch.Tensor:
        """rows of time_embed(pe[t]) (reference TimestepEmbedder.forward, mdm_forstyledataset.py:421)."""
        t = t.to(self.device, torch.int64).contiguous()
        n = t.numel()
        out = torch.empty(n, self.d_model, dtype=torch.float32, device=self.device)
        scratch = torch.empty(n, self.d_model, dtype=torch.float32, device=self.device)
        L.check(self.lib.mst_time_embed(self._h, t.data_ptr(), n, out.data_ptr(), scratch.data_ptr(),
                                        scratch.numel() * 4, _stream_ptr()), "mst_time_embed")
        return out

    @_engine_device
    def text_embed(self, feat: torch.Tensor) -> torch.Tensor:
        """embed_text(feat) (reference mdm_forstyledataset.py:327)."""
        n = feat.shape[0]
        out = torch.empty(n, self.d_model, dtype=torch.float32, device=self.device)
        L.check(self.lib.mst_text_embed(self._h, _ptr(feat, name="text_feat"), n, out.data_ptr(), _stream_ptr()),
                "mst_text_embed")
        return out

    # -- forward -----------------------------------------------------------------
    @_engine_device
    def forward(self, x: torch.Tensor, temb: torch.Tensor, text_emb: Optional[torch.Tensor], *, cfg: bool = False,
                uncond: bool = False, out_cond: Optional[torch.Tensor] = None,
                out_uncond: Optional[torch.Tensor] = None, temb_row_dev: Optional[torch.Tensor] = None,
                temb_row_offset: int = 0, workspace: Optional[torch.Tensor] = None):
        """x [B,F,1,T] fp32 -> model output(s) [B,F,1,T].  With cfg=True both the
        conditional and the unconditional pass run batched and two tensors return."""
        B, T = x.shape[0], x.shape[-1]
        if x.numel() != B * self.n_feats * T:
            raise ValueError(f"x has shape {tuple(x.shape)}, expected [B,{self.n_feats},1,T]")
        if out_cond is None:
            out_cond = torch.empty_like(x)
        if cfg and out_uncond is None:
            out_uncond = torch.empty_like(x)
        ws = workspace if workspace is not None else self.workspace(B * (2 if cfg else 1), T)
        a = L.ForwardArgs()
        a.batch, a.n_frames, a.cfg, a.uncond = B, T, int(cfg), int(uncond)
        a.x = _ptr(x, name="x")
        a.temb = _ptr(temb, name="temb")
        a.temb_row_dev = _ptr(temb_row_dev, torch.int32, "temb_row_dev")
        a.temb_row_offset = temb_row_offset
        a.text_emb = _ptr(text_emb, name="text_emb")
        a.out_cond = _ptr(out_cond, name="out_cond")
        a.out_uncond = _ptr(out_uncond, name="out_uncond") if cfg else None
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        L.check(self.lib.mst_denoiser_forward(self._h, C.byref(a), _stream_ptr()), "mst_denoiser_forward")
        return (out_cond, out_uncond) if cfg else out_cond


    # -- training path (fp32 engines) ------------------------------------------------
    def train_sizes(self, n_seqs: int, seq_len: int):
        """(tape bytes, backward scratch bytes) for n_seqs sequences of seq_len tokens."""
        tb, sb = C.c_size_t(), C.c_size_t()
        L.check(self.lib.mst_train_sizes(self._h, n_seqs, seq_len, C.byref(tb), C.byref(sb)), "mst_train_sizes")
        return int(tb.value), int(sb.value)

    def _scratch(self, nbytes: int) -> torch.Tensor:
        """backward scratch, one buffer PER STREAM: the trainer back-propagates the text-to-motion batch on a side stream
        while the style steps run on the main one, through the same engine"""
        pool = self.__dict__.setdefault("_bwd_scratch", {})
        key = _stream_ptr()
        buf = pool.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = pool[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return buf

    @_engine_device
    def forward_train(self, x: torch.Tensor, temb: torch.Tensor, text_emb: Optional[torch.Tensor], *,
                      uncond: bool = False, tape: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                      use_graph: bool = False, dropout_p: float = 0.0, dropout_seed: Optional[torch.Tensor] = None,
                      tape_seqs: int = 0, tape_seq_offset: int = 0):
        """Denoiser forward that records the activation tape.  Returns (out [B,F,1,T], tape).
        use_graph=True promises that x / temb / text_emb / out / tape are the same buffers on every call with this
        shape (a TapeSlot): the launch sequence is then captured once and replayed as a CUDA graph."""
        B, T = x.shape[0], x.shape[-1]
        if x.numel() != B * self.n_feats * T:
            raise ValueError(f"x has shape {tuple(x.shape)}, expected [B,{self.n_feats},1,T]")
        if tape is None:
            tape_bytes, _ = self.train_sizes(max(B, tape_seqs), T + 1)
            tape = torch.empty(tape_bytes, dtype=torch.uint8, device=self.device)
        if out is None:
            out = torch.empty_like(x)
        a = L.ForwardArgs()
        a.use_graph = int(bool(use_graph))
        a.dropout_p = float(dropout_p)
        a.dropout_seed = _ptr(dropout_seed, torch.int64, "dropout_seed") if dropout_p > 0 else None
        a.tape_seqs, a.tape_seq_offset = int(tape_seqs), int(tape_seq_offset)
        if dropout_p > 0 and dropout_seed.numel() < B:
            raise ValueError("dropout_seed needs one key per sequence of the call")
        a.batch, a.n_frames, a.cfg, a.uncond = B, T, 0, int(uncond)
        a.x = _ptr(x, name="x")
        a.temb = _ptr(temb, name="temb")
        a.temb_row_dev, a.temb_row_offset = None, 0
        a.text_emb = _ptr(text_emb, name="text_emb")
        a.out_cond, a.out_uncond = _ptr(out, name="out"), None
        a.workspace, a.workspace_bytes = None, 0
        L.check(self.lib.mst_denoiser_forward_train(self._h, C.byref(a), tape.data_ptr(), tape.numel(), _stream_ptr()),
                "mst_denoiser_forward_train")
        return out, tape

    @_engine_device
    def backward(self, d_out: torch.Tensor, tape: torch.Tensor, layer_grads: list, want_dx: bool = False,
                 use_graph: bool = False, dropout_p: float = 0.0, dropout_seed: Optional[torch.Tensor] = None,
                 tape_seqs: int = 0, tape_seq_offset: int = 0):
        """Back-propagate d_out [B,F,1,T] through the taped forward(s) of sequences [tape_seq_offset, +B) of the tape.  layer_grads: per layer a dict keyed by LAYER_KEYS
        of fp32 CUDA tensors (or None) that the gradients are ACCUMULATED into.  Returns d_x or None."""
        B, T = d_out.shape[0], d_out.shape[-1]
        _, scratch_bytes = self.train_sizes(B, T + 1)
        scratch = self._scratch(scratch_bytes)
        n_layers = self.desc.n_layers
        if len(layer_grads) != n_layers:
            raise ValueError(f"expected {n_layers} layer gradient dicts, got {len(layer_grads)}")
        arr = (L.LayerGrads * n_layers)()
        for i, lg in enumerate(layer_grads):
            for k in self.LAYER_KEYS:
                setattr(arr[i], k, _ptr(lg.get(k), name=f"grad.layer{i}.{k}"))
        d_x = torch.empty_like(d_out) if want_dx else None
        a = L.BackwardArgs()
        a.use_graph = int(bool(use_graph))
        a.dropout_p = float(dropout_p)
        a.dropout_seed = _ptr(dropout_seed, torch.int64, "dropout_seed") if dropout_p > 0 else None
        a.tape_seqs, a.tape_seq_offset = int(tape_seqs), int(tape_seq_offset)
        a.batch, a.n_frames = B, T
        a.d_out, a.d_x = _ptr(d_out, name="d_out"), _ptr(d_x, name="d_x")
        a.layer_grads = arr
        a.tape, a.tape_bytes = tape.data_ptr(), tape.numel()
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        L.check(self.lib.mst_denoiser_backward(self._h, C.byref(a), _stream_ptr()), "mst_denoiser_backward")
        return d_x

    @_engine_device
    def motion_encoder_forward(self, x: torch.Tensor, key_valid: Optional[torch.Tensor], mu_query: torch.Tensor,
                               sigma_query: torch.Tensor, dropout_p: float = 0.0,
                               dropout_seed: Optional[torch.Tensor] = None, tape: Optional[torch.Tensor] = None,
                               mu: Optional[torch.Tensor] = None, use_graph: bool = False):
        """MotionEncoder.forward on this engine's stack.  Returns (mu [B,d], tape)."""
        B, T = x.shape[0], x.shape[-1]
        if tape is None:
            tape_bytes, _ = self.train_sizes(B, T + 2)
            tape = torch.empty(tape_bytes, dtype=torch.uint8, device=self.device)
        if mu is None:
            mu = torch.empty(B, self.d_model, dtype=torch.float32, device=self.device)
        L.check(self.lib.mst_motion_encoder_forward(
            self._h, _ptr(x, name="x"), _ptr(key_valid, torch.uint8, "key_valid"), _ptr(mu_query, name="muQuery"),
            _ptr(sigma_query, name="sigmaQuery"), B, T, mu.data_ptr(), tape.data_ptr(), tape.numel(), float(dropout_p),
            _ptr(dropout_seed, torch.int64, "dropout_seed") if dropout_p > 0 else None, int(bool(use_graph)),
            _stream_ptr()), "mst_motion_encoder_forward")
        return mu, tape

    @_engine_device
    def motion_encoder_backward(self, d_mu: torch.Tensor, tape: torch.Tensor, shape, dropout_p: float = 0.0,
                                dropout_seed: Optional[torch.Tensor] = None, d_x: Optional[torch.Tensor] = None,
                                use_graph: bool = False):
        B, T = shape[0], shape[-1]
        _, scratch_bytes = self.train_sizes(B, T + 2)
        scratch = self._scratch(scratch_bytes)
        if d_x is None:
            d_x = torch.empty(shape, dtype=torch.float32, device=self.device)
        L.check(self.lib.mst_motion_encoder_backward(
            self._h, _ptr(d_mu, name="d_mu"), B, T, d_x.data_ptr(), tape.data_ptr(), tape.numel(), scratch.data_ptr(),
            scratch.numel(), float(dropout_p), _ptr(dropout_seed, torch.int64, "dropout_seed") if dropout_p > 0 else None,
            int(bool(use_graph)), _stream_ptr()), "mst_motion_encoder_backward")
        return d_x


class ClipTextEngine:
    """The CLIP text tower on one GPU (`mst_clip_text_*`, csrc/text.cu)."""

    LAYER_KEYS = L._CLIP_LAYER_FIELDS
    TOP_KEYS = ("token_embedding", "positional_embedding", "lnf_g", "lnf_b", "text_projection")

    def __init__(self, vocab: int = 49408, ctx: int = 77, width: int = 512, n_heads: int = 8, n_layers: int = 12,
                 d_ff: Optional[int] = None, d_out: int = 512, precision: Optional[str] = None,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("mst ClipTextEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.precision = precision or default_precision()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.desc = L.ClipTextDesc(vocab, ctx, width, n_heads, n_layers, d_ff or 4 * width, d_out,
                                   L.PREC_BF16 if self.precision == "bf16" else L.PREC_FP32)
        h = C.c_void_p()
        L.check(self.lib.mst_clip_text_create(C.byref(self.desc), C.byref(h)), "mst_clip_text_create")
        self._h = h
        self._packed = None
        self._keep = None
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.mst_clip_text_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def load_weights(self, top: dict, layers: list):
        """top: TOP_KEYS -> fp32 CUDA tensors; layers: list of dicts keyed by LAYER_KEYS."""
        w = L.ClipTextWeights()
        keep = []

        def dev(t):
            t = t.detach()
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(self.device, torch.float32).contiguous()
            keep.append(t)
            return t

        if len(layers) != self.desc.n_layers:
            raise ValueError(f"expected {self.desc.n_layers} layers, got {len(layers)}")
        with torch.cuda.device(self.device):
            for k in self.TOP_KEYS:
                setattr(w, k, _ptr(dev(top[k]), name=k))
            for i, lw in enumerate(layers):
                for k in self.LAYER_KEYS:
                    setattr(w.layers[i], k, _ptr(dev(lw[k]), name=f"layer{i}.{k}"))
            nbytes = C.c_size_t()
            L.check(self.lib.mst_clip_text_packed_weight_bytes(self._h, C.byref(nbytes)))
            if nbytes.value and (self._packed is None or self._packed.numel() < nbytes.value):
                self._packed = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            L.check(self.lib.mst_clip_text_load_weights(self._h, C.byref(w),
                                                        self._packed.data_ptr() if nbytes.value else None, nbytes.value,
                                                        _stream_ptr()), "mst_clip_text_load_weights")
        self._keep = keep

    def encode(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int32 [B, ctx] CUDA -> text features fp32 [B, d_out]."""
        if tokens.dim() != 2 or tokens.shape[1] != self.desc.ctx:
            raise ValueError(f"tokens: expected [B, {self.desc.ctx}], got {tuple(tokens.shape)}")
        B = tokens.shape[0]
        with torch.cuda.device(self.device):
            nbytes = C.c_size_t()
            L.check(self.lib.mst_clip_text_workspace_bytes(self._h, B, C.byref(nbytes)))
            if self._ws is None or self._ws.numel() < nbytes.value:
                self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            out = torch.empty(B, self.desc.d_out, dtype=torch.float32, device=self.device)
            L.check(self.lib.mst_clip_text_encode(self._h, _ptr(tokens, torch.int32, "tokens"), B, out.data_ptr(),
                                                  self._ws.data_ptr(), self._ws.numel(), _stream_ptr()),
                    "mst_clip_text_encode")
        return out


class TapeSlot:
    """Persistent buffers for up to ``capacity`` taped forwards of ``B`` sequences x ``T`` frames each, recorded into
    slices of ONE tape.  Every pointer is stable from one training step to the next, so the C side replays the launch
    sequences as CUDA graphs (mst_forward_args.use_graph), and forwards recorded one by one - the six differentiable
    DDIM steps of the finetune loss - are back-propagated by ONE batched backward over the whole tape."""

    def __init__(self, eng: "Engine", B: int, T: int, has_text: bool, capacity: int = 1):
        dev, f32 = eng.device, torch.float32
        self.eng, self.B, self.T, self.capacity = eng, B, T, capacity
        n = B * capacity
        self.tape_seqs = n
        tape_bytes, _ = eng.train_sizes(n, T + 1)
        self.tape = torch.empty(tape_bytes, dtype=torch.uint8, device=dev)
        self.x = torch.empty(n, eng.n_feats, 1, T, dtype=f32, device=dev)
        self.temb = torch.empty(n, eng.d_model, dtype=f32, device=dev)
        self.text = torch.empty(n, eng.d_model, dtype=f32, device=dev) if has_text else None
        self.out = torch.empty(n, eng.n_feats, 1, T, dtype=f32, device=dev)
        self.d_out = torch.empty(n, eng.n_feats, 1, T, dtype=f32, device=dev)
        self.seed = torch.zeros(n, dtype=torch.int64, device=dev)   # one Philox key per sequence (dropout masks)
        self.epoch = 0      # bumped at every reset: a stale autograd node can tell its tape is gone
        self.used = 0       # forwards handed out since the last reset
        self.pending = {}   # call index -> (dropout_p, layer_grads) whose d_out is staged, backward not yet run
        self.done = set()

    def rows(self, k):
        return slice(k * self.B, (k + 1) * self.B)

    def reset(self):
        self.flush()
        self.used, self.done = 0, set()
        self.epoch += 1

    def stage_backward(self, k, d_out, dropout_p, grads):
        """Record the output gradient of call k; run the batched backward once every call of this step has reported."""
        self.d_out[self.rows(k)].copy_(d_out)
        self.pending[k] = (dropout_p, grads)
        if len(self.pending) + len(self.done) == self.used:
            self.flush()

    def flush(self):
        """Back-propagate every staged call: one launch sequence (graph) per run of consecutive calls."""
        ks = sorted(self.pending)
        i = 0
        while i < len(ks):
            j = i
            while j + 1 < len(ks) and ks[j + 1] == ks[j] + 1 and self.pending[ks[j + 1]][0] == self.pending[ks[i]][0]:
                j += 1
            k0, k1 = ks[i], ks[j] + 1
            p, grads = self.pending[k0]
            lo, hi = k0 * self.B, k1 * self.B
            self.eng.backward(self.d_out[lo:hi], self.tape, grads, want_dx=False, use_graph=True, dropout_p=p,
                              dropout_seed=self.seed[lo:hi], tape_seqs=self.tape_seqs, tape_seq_offset=lo)
            i = j + 1
        self.done.update(ks)
        self.pending = {}


# ---------------------------------------------------------------------------------
# training kernels that need no engine
# ---------------------------------------------------------------------------------
@_tensor_device
def masked_l2_forward(a, b, mask):
    """masked_l2 rows (reference gaussian_diffusion.py:223-235): a [Ra,F,1,T] (rows repeat), b [R,F,1,T], mask [Rm,1,1,T]."""
    R, F, T = b.shape[0], b.shape[1] * b.shape[2], b.shape[3]
    loss = torch.empty(R, dtype=torch.float32, device=b.device)
    L.check(L.load().mst_masked_l2(_ptr(a, name="a"), _ptr(b, name="b"), _ptr(mask, name="mask"), loss.data_ptr(), None, None,
                                   R, a.shape[0], mask.shape[0], F, T, _stream_ptr()), "mst_masked_l2")
    return loss


@_tensor_device
def masked_l2_backward(a, b, mask, grad_loss):
    R, F, T = b.shape[0], b.shape[1] * b.shape[2], b.shape[3]
    gb = torch.empty_like(b)
    L.check(L.load().mst_masked_l2(_ptr(a, name="a"), _ptr(b, name="b"), _ptr(mask, name="mask"), None,
                                   _ptr(grad_loss, name="grad_loss"), gb.data_ptr(), R, a.shape[0], mask.shape[0], F, T,
                                   _stream_ptr()), "mst_masked_l2")
    return gb


@_tensor_device
def update_step_backward(d_x0, d_sample, k_table, t_vec, mask, pred_xstart, clip_denoised, shape):
    B, F, T = shape[0], shape[1] * shape[2], shape[3]
    ref = d_x0 if d_x0 is not None else d_sample
    d_out = torch.empty(shape, dtype=torch.float32, device=ref.device)
    L.check(L.load().mst_update_step_backward(
        _ptr(d_x0, name="d_pred_xstart"), _ptr(d_sample, name="d_sample"), _ptr(k_table, name="k_table"),
        _ptr(t_vec, torch.int64, "t"), mask_kind_of(mask, shape), _ptr(mask, name="inpainting_mask"),
        _ptr(pred_xstart, name="pred_xstart") if clip_denoised else None, int(bool(clip_denoised)), d_out.data_ptr(),
        B, F, T, _stream_ptr()), "mst_update_step_backward")
    return d_out


@_tensor_device
def recover_from_ric(x, joints_num, mean=None, std=None):
    """x [B,F,1,T] (sampler layout) -> joints [B,1,T,J,3]: inv_transform + recover_from_ric fused (scope row N2)."""
    B, F, T = x.shape[0], x.shape[1] * x.shape[2], x.shape[3]
    out = torch.empty(B, 1, T, joints_num, 3, dtype=torch.float32, device=x.device)
    L.check(L.load().mst_recover_from_ric(_ptr(x, name="x"), _ptr(mean, name="mean"), _ptr(std, name="std"), out.data_ptr(),
                                          B, F, T, int(joints_num), _stream_ptr()), "mst_recover_from_ric")
    return out


@_tensor_device
def dropout_scale(n, p, seed, site):
    """Test hook: the 0 / 1/(1-p) multipliers of dropout site ``site`` for the key in ``seed`` (int64 [1] device)."""
    out = torch.empty(n, dtype=torch.float32, device=seed.device)
    L.check(L.load().mst_test_dropout_scale(out.data_ptr(), n, float(p), _ptr(seed, torch.int64, "seed"), int(site),
                                            _stream_ptr()), "mst_test_dropout_scale")
    return out


@_tensor_device
def adamw_step(params, grads, exp_avg, exp_avg_sq, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1,
               grad_scale=1.0):
    """torch.optim.AdamW step over flat fp32 arenas, one launch."""
    n = params.numel()
    if not (grads.numel() == exp_avg.numel() == exp_avg_sq.numel() == n):
        raise ValueError("adamw_step: arenas differ in size")
    L.check(L.load().mst_adamw_step(_ptr(params, name="params"), _ptr(grads, name="grads"), _ptr(exp_avg, name="exp_avg"),
                                    _ptr(exp_avg_sq, name="exp_avg_sq"), n, lr, beta1, beta2, eps, weight_decay, int(step),
                                    grad_scale, _stream_ptr()), "mst_adamw_step")


@_tensor_device
def sumsq2(x, y=None):
    """device float64 [2] = (sum x^2, sum y^2)."""
    out = torch.empty(2, dtype=torch.float64, device=x.device)
    L.check(L.load().mst_sumsq2(_ptr(x, name="x"), _ptr(y, name="y"), x.numel(), out.data_ptr(), _stream_ptr()), "mst_sumsq2")
    return out


# ---------------------------------------------------------------------------------
# fused diffusion kernels (no engine needed)
# ---------------------------------------------------------------------------------
def mask_kind_of(mask: Optional[torch.Tensor], shape) -> int:
    if mask is None:
        return L.MASK_NONE
    B, F, T = shape[0], shape[1] * shape[2], shape[3]
    n = mask.numel()
    if n == B * F * T:
        return L.MASK_FULL
    if n == F * T:
        return L.MASK_FT
    if n == F:
        return L.MASK_F
    raise ValueError(f"inpainting mask with {n} elements does not match state shape {tuple(shape)}")


@_tensor_device
def update_step(*, sampler: int, out_cond, x_t, x_prev, coef1, coef2, sigma=None, recip=None, recipm1=None,
                out_uncond=None, cfg_scale=None, pred_xstart=None, mask=None, x_inpaint=None, mask_noise=True,
                clip_denoised=False, t_vec=None, t_scalar_dev=None, t_imm=0, advance_t=False, block_counter=None,
                noise_kind=L.NOISE_NONE, noise=None, const_noise=False, philox_seed=0, philox_sample_offset=0):
    lib = L.load()
    B, F, T = x_t.shape[0], x_t.shape[1] * x_t.shape[2], x_t.shape[3]
    a = L.UpdateArgs()
    a.batch, a.n_feats, a.n_frames = B, F, T
    a.sampler, a.clip_denoised = sampler, int(bool(clip_denoised))
    a.out_cond = _ptr(out_cond, name="out_cond")
    a.out_uncond = _ptr(out_uncond, name="out_uncond")
    a.cfg_scale = _ptr(cfg_scale, name="cfg_scale")
    a.x_t = _ptr(x_t, name="x_t")
    a.x_prev = _ptr(x_prev, name="x_prev")
    a.pred_xstart = _ptr(pred_xstart, name="pred_xstart")
    a.mask_kind = mask_kind_of(mask, x_t.shape)
    a.mask = _ptr(mask, name="inpainting_mask")
    a.x_inpaint = _ptr(x_inpaint, name="inpainted_motion")
    a.mask_noise = int(bool(mask_noise))
    a.t_vec = _ptr(t_vec, torch.int64, "t")
    a.t_scalar_dev = _ptr(t_scalar_dev, torch.int32, "t_scalar_dev")
    a.t_imm, a.advance_t = int(t_imm), int(bool(advance_t))
    a.block_counter = _ptr(block_counter, torch.int32, "block_counter")
    a.coef1, a.coef2 = _ptr(coef1, name="coef1"), _ptr(coef2, name="coef2")
    a.sigma, a.recip, a.recipm1 = _ptr(sigma, name="sigma"), _ptr(recip, name="recip"), _ptr(recipm1, name="recipm1")
    a.noise_kind = noise_kind
    a.noise = _ptr(noise, name="noise")
    a.const_noise = int(bool(const_noise))
    a.philox_seed, a.philox_sample_offset = int(philox_seed) & (2**64 - 1), int(philox_sample_offset)
    L.check(lib.mst_update_step(C.byref(a), _stream_ptr()), "mst_update_step")
    return x_prev


@_tensor_device
def q_sample(x_start, noise, mask, t_vec, t_imm, sqrt_ab, sqrt_1m_ab, out=None):
    lib = L.load()
    B, F, T = x_start.shape[0], x_start.shape[1] * x_start.shape[2], x_start.shape[3]
    if out is None:
        out = torch.empty_like(x_start)
    L.check(lib.mst_q_sample(_ptr(x_start, name="x_start"), _ptr(noise, name="noise"), mask_kind_of(mask, x_start.shape),
                             _ptr(mask, name="inpainting_mask"), _ptr(t_vec, torch.int64, "t"), int(t_imm),
                             _ptr(sqrt_ab, name="sqrt_ab"), _ptr(sqrt_1m_ab, name="sqrt_1m_ab"), out.data_ptr(), B, F, T,
                             _stream_ptr()), "mst_q_sample")
    return out


@_tensor_device
def cfg_combine(out_cond, out_uncond, scale, out=None):
    lib = L.load()
    if out is None:
        out = torch.empty_like(out_cond)
    B = out_cond.shape[0]
    L.check(lib.mst_cfg_combine(_ptr(out_cond, name="out_cond"), _ptr(out_uncond, name="out_uncond"),
                                _ptr(scale, name="scale"), out.data_ptr(), B, out_cond.numel() // B, _stream_ptr()),
            "mst_cfg_combine")
    return out


@_tensor_device
def philox_normal(shape, seed: int, sample_offset: int, t: int, device) -> torch.Tensor:
    lib = L.load()
    out = torch.empty(shape, dtype=torch.float32, device=device)
    B = shape[0]
    L.check(lib.mst_philox_normal(out.data_ptr(), B, out.numel() // B, int(seed) & (2**64 - 1), int(sample_offset), int(t),
                                  _stream_ptr()), "mst_philox_normal")
    return out


# ---------------------------------------------------------------------------------
# launch accounting / per-launch timing (mst_launch_count, mst_profile_begin/end)
# ---------------------------------------------------------------------------------
_graph_replays = 0


def count_graph_replay(n: int = 1):
    global _graph_replays
    _graph_replays += n


def graph_replays() -> int:
    """CUDA-graph replays of captured denoise steps issued by this process (each re-launches
    ``diffusion.last_plan_launches`` kernels without passing through the host library)."""
    return _graph_replays


def launch_count() -> int:
    """Kernels launched by libmst_b200.so in this process so far (graph replays excluded, see mst.h)."""
    return int(L.load().mst_launch_count())


class profile:
    """``with profile() as p: ...`` -> ``p.records`` = [(kernel name, ms)] for every launch the block made
    through the library on the current stream (CUDA events between launches; synchronises on exit)."""

    def __init__(self, cap: int = 4096, deferred: bool = False):
        """deferred=True: the block is being captured into a CUDA graph; call ``collect()`` after replaying it."""
        self.cap, self.records, self.deferred = cap, [], deferred

    def __enter__(self):
        L.check(L.load().mst_profile_begin(_stream_ptr()), "mst_profile_begin")
        return self

    def _read(self, fn, what):
        ms = (C.c_float * self.cap)()
        names = C.create_string_buffer(self.cap * 32)
        n = C.c_int32()
        L.check(fn(ms, names, self.cap, len(names), C.byref(n)), what)
        nm = names.value.decode().split("\n")[:-1]
        self.records = [(nm[i] if i < len(nm) else "?", float(ms[i])) for i in range(min(n.value, self.cap))]
        return self.records

    def __exit__(self, *exc):
        if self.deferred:
            n = C.c_int32()
            L.check(L.load().mst_profile_end(None, None, 0, 0, C.byref(n)), "mst_profile_end")
        else:
            self._read(L.load().mst_profile_end, "mst_profile_end")
        return False

    def collect(self):
        """per-launch times of the most recent replay of the graph the block was captured into"""
        return self._read(L.load().mst_profile_collect, "mst_profile_collect")
