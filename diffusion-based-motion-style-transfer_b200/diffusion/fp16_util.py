"""``MixedPrecisionTrainer`` of the reference (``diffusion/fp16_util.py:148-235``), fp32 only (the reference
hard-codes ``use_fp16 = False``, ``train/training_loop.py:51``), re-laid for one GPU kernel per operation:

* every parameter of the model is moved into ONE flat fp32 arena (trainable tensors first); ``p.data`` become views;
* ``p.grad`` of the trainable tensors are views of a second flat arena, so ``zero_grad`` is one memset, the
  data-parallel gradient exchange is one all-reduce, and the optimizer (``train.training_loop.FusedAdamW``) is
  one ``mst_adamw_step`` launch over the arena;
* ``_compute_norms`` is one ``mst_sumsq2`` launch and ONE device-to-host read instead of 192 ``.item()`` syncs.
"""
from __future__ import annotations

import numpy as np
import torch as th

from .. import engine as K


class FlatParams:
    """Flat arenas behind a model's parameters."""

    def __init__(self, model):
        params = [p for n, p in model.named_parameters() if not n.startswith("clip_model.") and ".clip_model." not in n]
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.trainable = [p for p in uniq if p.requires_grad]
        self.frozen = [p for p in uniq if not p.requires_grad]
        if not self.trainable:
            raise ValueError("the model has no trainable parameters")
        dev = self.trainable[0].device
        if dev.type != "cuda":
            raise RuntimeError("the mst trainer needs the model on a CUDA device (there is no CPU path)")
        if any(p.dtype != th.float32 or p.device != dev for p in uniq):
            raise RuntimeError("all parameters must be fp32 on one CUDA device")
        self.n_train = sum(p.numel() for p in self.trainable)
        n_all = self.n_train + sum(p.numel() for p in self.frozen)
        self.params = th.empty(n_all, dtype=th.float32, device=dev)
        self.grads = th.zeros(self.n_train, dtype=th.float32, device=dev)
        off = 0
        with th.no_grad():
            for p in self.trainable + self.frozen:
                n = p.numel()
                view = self.params[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                if p.requires_grad:
                    p.grad = self.grads[off:off + n].view(p.shape)
                off += n

    @property
    def train_params(self):
        return self.params[:self.n_train]

    def ensure_grad_views(self):
        """Re-attach the arena views if someone replaced / dropped ``p.grad`` (e.g. zero_grad(set_to_none=True))."""
        views = self.__dict__.get("_grad_views")
        if views is None:
            views, off = [], 0
            for p in self.trainable:
                n = p.numel()
                views.append(self.grads[off:off + n].view(p.shape))
                off += n
            self._grad_views = views
        for p, view in zip(self.trainable, views):
            g = p.grad
            if g is view:
                continue
            if g is None:
                p.grad = view
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
            else:
                p.grad = view


class MixedPrecisionTrainer:
    def __init__(self, *, model, use_fp16=False, fp16_scale_growth=1e-3, initial_lg_loss_scale=20.0):
        if use_fp16:
            raise NotImplementedError("fp16 master-weight training is deprecated in the reference "
                                      "(train/training_loop.py:51) and not built")
        self.model = model
        self.use_fp16 = False
        self.flat = FlatParams(model)
        self.model_params = self.flat.trainable + self.flat.frozen
        self.master_params = self.model_params
        self.lg_loss_scale = initial_lg_loss_scale
        self._norms_dev = None
        # gradients live in the arena from now on: the model may run its forward / backward on pooled tapes and
        # accumulate straight into them (CUDA-graph replay, model/mdm_forstyledataset.py::_DenoiserGradFn)
        # the native denoisers of the model (walked once: model.modules() costs ~0.1 ms per call and a step asks 5 times)
        self._denoisers = [m for m in model.modules() if hasattr(m, "mst_tape_pool")]
        for m in self._denoisers:
            m.mst_tape_pool = True
            m.__dict__["mst_direct_grads"] = True  # every trainable parameter keeps an arena view as .grad (ensure_grad_views)

    def master_params_to_state_dict(self, master_params=None):
        """reference :225-228: the model's state_dict (the parameters ARE the master params in fp32 mode)"""
        return self.model.state_dict()

    def state_dict_to_master_params(self, state_dict):
        return [state_dict[name] for name, _ in self.model.named_parameters()]

    def zero_grad(self):
        self.flat.ensure_grad_views()
        self.flat.grads.zero_()
        for m in self._denoisers:
            m.mst_tape_reset()

    def _flush(self):
        for m in self._denoisers:
            m.mst_flush_backward()

    def backward(self, loss: th.Tensor):
        loss.backward()
        self._flush()   # batched backward passes still waiting for a forward whose output never reached the loss
        self.flat.ensure_grad_views()

    def backward_into(self, loss: th.Tensor, grad_views):
        """Back-propagate ``loss`` with every trainable parameter's ``.grad`` pointing at ``grad_views`` (a second
        arena, same layout as ``flat.grads``) and put the arena views back afterwards: lets one loss term accumulate its
        gradient apart from the others (on another stream, to be all-reduced on its own)."""
        own = self.flat.__dict__.get("_grad_views")
        if own is None:
            self.flat.ensure_grad_views()
            own = self.flat._grad_views
        params = self.flat.trainable
        for p, v in zip(params, grad_views):
            p.grad = v
        try:
            loss.backward()
            self._flush()
        finally:
            for p, v in zip(params, own):
                p.grad = v

    def optimize(self, opt):
        # after a SUM all-reduce the arena holds world_size x the mean gradient; the optimiser applies it scaled by
        # opt.grad_scale = 1 / world_size, and the logged norm is that of the applied (mean) gradient
        scale = getattr(opt, "grad_scale", 1.0)
        # the two sums of squares stay on the device: reading them here would drain the GPU once per step (the host then
        # cannot issue step i+1 while step i still runs); `last_norms` fetches them when somebody looks
        self._flush()
        self._norms_dev = (K.sumsq2(self.flat.grads, None), K.sumsq2(self.flat.params, None),
                           1.0 / scale if scale else 1.0)
        opt.step()
        return True

    @property
    def last_norms(self):
        """(||grad||_2 of the applied gradient, ||param||_2 before the step) of the most recent optimize() (reference
        :215-223); synchronises with the device."""
        dev = self.__dict__.get("_norms_dev")
        if dev is None:
            return (0.0, 0.0)
        both = th.stack([dev[0], dev[1]]).cpu()  # ONE host read
        return float(np.sqrt(both[0, 0].item())) / dev[2], float(np.sqrt(both[1, 0].item()))

    def _compute_norms(self, grad_scale=1.0):
        """(||grad||_2 / grad_scale, ||param||_2) over all master params (reference :215-223)."""
        self._flush()
        both = th.stack([K.sumsq2(self.flat.grads, None), K.sumsq2(self.flat.params, None)]).cpu()  # ONE host read
        return float(np.sqrt(both[0, 0].item())) / grad_scale, float(np.sqrt(both[1, 0].item()))
