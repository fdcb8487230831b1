"""Few-shot style finetune loop (reference ``train/training_loop.py:36-303``) on the hand-written CUDA path.

Kept from the reference: the constructor's argument names, ``run_step`` / ``forward_backward`` / ``_anneal_lr`` and
their order of operations (zero_grad -> few_shot_style_finetune_losses -> backward -> norms -> AdamW -> lr anneal).
Also kept: the epoch-bounded ``run_loop`` (``num_steps // len(data) + 1`` epochs), ``save()`` every ``save_interval`` steps
and at the end in the reference's on-disk format (``modelNNNNNNNNN.pt`` = the model ``state_dict`` minus the frozen
``motion_enc.`` / ``controlmdm.`` / ``clip_model.`` entries, ``optNNNNNNNNN.pt`` = a ``torch.optim.AdamW`` state dict indexed
like ``list(model.parameters())``) and the resume path (``resume_checkpoint`` file or directory), so that
``train/finetune_style_diffusion.py`` leaves the same files behind.  Not kept: logger / train-platform reporting.

Data parallelism: when ``torch.distributed`` is initialised with more than one rank, every rank
takes a contiguous shard of the text-to-motion batch (the only batched term of the loss), keeps the B=1 style term
replicated, and the flat gradient arena is summed with ONE all-reduce (NCCL over NVLink on GPUs) before the fused
AdamW step applies it scaled by 1/world_size.
"""
from __future__ import annotations

import functools
import os
from collections import OrderedDict

import torch
import torch.distributed as dist

from .. import engine as K
from ..diffusion.fp16_util import MixedPrecisionTrainer
from ..diffusion.resample import LossAwareSampler, create_named_schedule_sampler


class FusedAdamW:
    """torch.optim.AdamW(lr, weight_decay) over the trainer's flat arenas: one kernel per step."""

    def __init__(self, flat, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, model=None):
        self.flat, self.model = flat, model
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        self.exp_avg = torch.zeros_like(flat.grads)
        self.exp_avg_sq = torch.zeros_like(flat.grads)
        self.step_count = 0
        self.grad_scale = 1.0

    def step(self):
        g = self.param_groups[0]
        self.step_count += 1
        if self.model is not None:
            den = self.__dict__.get("_denoisers")
            if den is None:
                den = self._denoisers = [m for m in self.model.modules() if hasattr(m, "mst_flush_backward")]
            for m in den:
                m.mst_flush_backward()  # deferred (batched) backward passes must have written their gradients
        K.adamw_step(self.flat.train_params, self.flat.grads, self.exp_avg, self.exp_avg_sq, lr=g["lr"],
                     beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"], weight_decay=g["weight_decay"],
                     step=self.step_count, grad_scale=self.grad_scale)
        if self.model is not None and hasattr(self.model, "mst_weights_changed"):
            self.model.mst_weights_changed(structure=False)  # the kernel wrote through raw pointers: engines must re-pack

    def zero_grad(self, set_to_none=False):
        self.flat.grads.zero_()

    def _index_map(self):
        """(index in list(model.parameters()), offset, numel) of every trainable tensor, arena order."""
        order = {id(p): i for i, p in enumerate(self.model.parameters())} if self.model is not None else {}
        out, off = [], 0
        for k, p in enumerate(self.flat.trainable):
            out.append((order.get(id(p), k), off, p.numel(), tuple(p.shape)))
            off += p.numel()
        return out

    def state_dict(self):
        """``torch.optim.AdamW.state_dict()`` layout over ``list(model.parameters())`` (what the reference writes to
        ``optNNNNNNNNN.pt``, train/training_loop.py:343-348): per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``."""
        g = self.param_groups[0]
        n_all = len(list(self.model.parameters())) if self.model is not None else len(self.flat.trainable)
        state = {}
        if self.step_count > 0:
            for idx, off, n, shape in self._index_map():
                state[idx] = {"step": self.step_count, "exp_avg": self.exp_avg[off:off + n].view(shape).clone(),
                              "exp_avg_sq": self.exp_avg_sq[off:off + n].view(shape).clone()}
        group = {"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": g["weight_decay"],
                 "amsgrad": False, "params": list(range(n_all))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        if "state" in sd:  # torch.optim.AdamW layout (ours or the reference's)
            steps = []
            for idx, off, n, shape in self._index_map():
                ent = sd["state"].get(idx)
                if ent is None:
                    continue
                self.exp_avg[off:off + n].copy_(ent["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(ent["exp_avg_sq"].reshape(-1))
                steps.append(int(ent["step"]))
            self.step_count = max(steps) if steps else 0
            g = sd["param_groups"][0]
            self.param_groups[0].update(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], weight_decay=g["weight_decay"])
            return
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])
        self.param_groups = sd["param_groups"]


def shard_batch(batch, cond, rank, world):
    """Contiguous shard of the t2m batch for this rank: tensors with a leading batch dimension and lists of
    per-sample entries are sliced; everything else is shared."""
    B = batch.shape[0]
    if B < world or B % world != 0:
        # equal shards only: the all-reduce averages the per-rank means with equal weight (a ragged split would bias
        # the text-to-motion term, an empty shard would hang the collective).  The replicated B=1 style term also
        # assumes that every rank draws the same noise / timesteps: seed the ranks identically.
        raise ValueError(f"the text-to-motion batch ({B}) must be a positive multiple of the world size ({world})")
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    y = {}
    for k, v in cond["y"].items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B:
            y[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == B:
            y[k] = v[lo:hi]
        else:
            y[k] = v
    return batch[lo:hi], {"y": y}, (lo, hi)


class TrainInpaintingLoop:
    def __init__(self, args, train_platform, model, data, diffusion=None, style_data=None):
        self.args = args
        self.dataset = getattr(args, "dataset", None)
        self.train_platform = train_platform
        self.model = model
        self.data = data
        self.style_data = style_data
        self.batch_size = args.batch_size
        self.microbatch = args.batch_size
        self.lr = args.lr
        self.weight_decay = getattr(args, "weight_decay", 0.0)
        self.lr_anneal_steps = getattr(args, "lr_anneal_steps", 0)
        self.style_finetune = getattr(args, "style_finetune", 0)
        self.semantic_guidance = getattr(args, "semantic_guidance", 0) if hasattr(args, "style_finetune") else 0
        self.skip_steps = getattr(args, "skip_steps", 0)
        self.log_interval = getattr(args, "log_interval", 0)
        self.save_interval = getattr(args, "save_interval", 0)
        self.save_dir = getattr(args, "save_dir", None)
        self.resume_checkpoint = getattr(args, "resume_checkpoint", "") or ""
        self.overwrite = getattr(args, "overwrite", False)
        self.step = 0
        self.resume_step = 0
        self.num_steps = getattr(args, "num_steps", 0)
        try:
            self.num_epochs = self.num_steps // max(1, len(data)) + 1  # reference :76
        except TypeError:
            self.num_epochs = 1
        self._load_and_sync_parameters()
        self.use_ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.use_ddp else 1
        self.rank = dist.get_rank() if self.use_ddp else 0
        self.mp_trainer = MixedPrecisionTrainer(model=self.model, use_fp16=False)
        self.diffusion = diffusion
        if diffusion is not None:
            self.schedule_sampler_type = 'uniform'
            self.schedule_sampler = create_named_schedule_sampler(self.schedule_sampler_type, diffusion)
        self.opt = FusedAdamW(self.mp_trainer.flat, lr=self.lr, weight_decay=self.weight_decay, model=self.model)
        self.opt.grad_scale = 1.0 / self.world
        if self.resume_step:
            self._load_optimizer_state()
        self.device = next(model.parameters()).device
        self.last_losses = None

    # ------------------------------------------------------------------ checkpoints (reference :108-140, :309-348)
    def _load_and_sync_parameters(self):
        ckpt = self.resume_checkpoint
        if ckpt and os.path.isdir(ckpt):
            ckpt = find_resume_checkpoint(ckpt, "model")
        if not ckpt:
            return
        self.resume_step = parse_resume_step_from_filename(ckpt)
        dev = next(self.model.parameters()).device
        missing, unexpected = self.model.load_state_dict(torch.load(ckpt, map_location=dev), strict=False)
        assert len(unexpected) == 0, unexpected
        assert all(k.startswith(("motion_enc.", "controlmdm.", "clip_model.")) for k in missing), missing
        if hasattr(self.model, "mst_weights_changed"):
            self.model.mst_weights_changed()

    def _load_optimizer_state(self):
        main = self.resume_checkpoint
        if os.path.isdir(main):
            main = find_resume_checkpoint(main, "opt")
        path = os.path.join(os.path.dirname(main), f"opt{self.resume_step:09}.pt")
        if os.path.exists(path):
            try:
                self.opt.load_state_dict(torch.load(path, map_location=next(self.model.parameters()).device))
            except Exception:  # the reference swallows a mismatching optimizer file too (:136-139)
                pass

    def ckpt_file_name(self, step=None):
        return f"model{((self.step if step is None else step) + self.resume_step):09d}.pt"

    def save(self, step=None):
        """``modelNNNNNNNNN.pt`` + ``optNNNNNNNNN.pt`` in ``save_dir`` (rank 0 only under data parallelism)."""
        if self.save_dir is None or self.rank != 0:
            return None
        step = self.step if step is None else step
        os.makedirs(self.save_dir, exist_ok=True)
        drop = ("controlmdm.", "clip_model.") if self.dataset == "humanml" else ("motion_enc.", "clip_model.")
        sd = OrderedDict((k, v.detach().clone()) for k, v in self.mp_trainer.master_params_to_state_dict().items()
                         if not k.startswith(drop))
        path = os.path.join(self.save_dir, self.ckpt_file_name(step))
        torch.save(sd, path)
        torch.save(self.opt.state_dict(), os.path.join(self.save_dir, f"opt{(step + self.resume_step):09d}.pt"))
        return path

    # ------------------------------------------------------------------ one optimisation step
    def run_step(self, batch, cond, style_batch=None, style_cond=None):
        self.forward_backward(batch, cond, style_batch, style_cond)
        self.sync_gradients()
        self.mp_trainer.optimize(self.opt)
        self._anneal_lr()
        self.step += 1

    # ---- the text-to-motion term on its own stream (semantic guidance on) ------------------------------------------
    # The step has two independent branches until the optimizer: the text-to-motion batch (denoiser + MotionEncoder forward
    # and backward: throughput work) and the six sequential B=1 style steps (a latency chain that occupies a handful of
    # SMs).  few_shot_style_finetune_losses runs the first branch on `diffusion.t2m_stream` and hands its loss term to
    # _early_t2m_backward as soon as it exists: the term is back-propagated THERE, into a second gradient arena, and -
    # data parallel - all-reduced asynchronously (only this term differs between ranks: the style term is replicated, same
    # seed, same inputs), all while the main stream works through the style steps.  sync_gradients joins the streams and
    # forms  style gradient + mean(t2m gradients)  for the fused AdamW.
    # MST_OVERLAP_ALLREDUCE=0 restores the serial step with one blocking all-reduce of the whole arena.
    def _early_t2m_backward(self, loss_t2m):
        flat = self.mp_trainer.flat
        if self.__dict__.get("_g_t2m") is None:
            self._g_t2m = torch.zeros_like(flat.grads)
            views, off = [], 0
            for p in flat.trainable:
                n = p.numel()
                views.append(self._g_t2m[off:off + n].view(p.shape))
                off += n
            self._g_t2m_views = views
        self._g_t2m.zero_()
        self.mp_trainer.backward_into(loss_t2m, self._g_t2m_views)
        ev = self.__dict__.get("_ov_events")
        if ev is None:
            ev = self._ov_events = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()                                   # (this stream) gradient complete, all-reduce handed to NCCL here
        self._t2m_work = (dist.all_reduce(self._g_t2m, op=dist.ReduceOp.SUM, async_op=True) if self.use_ddp else None)
        self._t2m_pending = True

    def sync_gradients(self):
        if self.__dict__.pop("_t2m_pending", False):
            work = self.__dict__.pop("_t2m_work", None)
            ev = self._ov_events
            main = torch.cuda.current_stream()
            ev[1].record()                               # the style term's gradient is complete
            if work is not None:
                work.wait()                              # this stream waits for NCCL's stream (a no-op if it is done)
            main.wait_stream(self.diffusion.t2m_stream)
            ev[2].record()
            self.mp_trainer.flat.grads.add_(self._g_t2m, alpha=1.0 / self.world)
            self.opt.grad_scale = 1.0                    # the arena already holds the mean gradient
            extra = self.last_losses.pop("loss_t2m", None)
            if extra is not None:                        # the logged loss is the whole objective again
                self.last_losses["loss"] = self.last_losses["loss"] + extra
            self._overlapped = True
            return
        self.opt.grad_scale = 1.0 / self.world
        self._overlapped = False
        if self.use_ddp:
            ev = self.__dict__.get("_ar_events")
            if ev is None and self.mp_trainer.flat.grads.is_cuda:
                ev = self._ar_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            if ev is not None:
                ev[0].record()
            dist.all_reduce(self.mp_trainer.flat.grads, op=dist.ReduceOp.SUM)
            if ev is not None:
                ev[1].record()

    def last_overlap(self):
        """(window ms between handing the all-reduce to NCCL and needing its result, exposed wait ms) of the most recent
        overlapped step, or None"""
        if not self.__dict__.get("_overlapped"):
            return None
        ev = self._ov_events
        ev[2].synchronize()
        return float(ev[0].elapsed_time(ev[1])), float(ev[1].elapsed_time(ev[2]))

    def last_allreduce_ms(self):
        """device time of the most recent gradient all-reduce (CUDA events around the collective; synchronises)"""
        ev = self.__dict__.get("_ar_events")
        if ev is None:
            return 0.0
        ev[1].synchronize()
        return float(ev[0].elapsed_time(ev[1]))

    def forward_backward(self, batch, cond, style_batch, style_cond):
        self.mp_trainer.zero_grad()
        assert self.diffusion is not None
        if not self.style_finetune:
            raise NotImplementedError("only the style-finetune objective is built (few_shot_style_finetune_losses); "
                                      "the reference's generic training_losses path is outside the scope table")
        if self.use_ddp:
            batch, cond, _ = shard_batch(batch, cond, self.rank, self.world)
        overlap = bool(self.semantic_guidance) and os.environ.get("MST_OVERLAP_ALLREDUCE", "1") != "0"
        self.diffusion.early_t2m_backward = self._early_t2m_backward if overlap else None
        if overlap and getattr(self.diffusion, "t2m_stream", None) is None:
            self.diffusion.t2m_stream = torch.cuda.Stream(device=self.device)
        args = self.args
        if getattr(args, "use_ddim", 0):
            rng = range(int((args.diffusion_steps - args.skip_steps) / args.diffusion_steps * 20))
        else:
            rng = range(args.diffusion_steps - args.skip_steps)
        t, weights = self.schedule_sampler.sample(batch.shape[0], self.device, rng)
        compute_losses = functools.partial(
            self.diffusion.few_shot_style_finetune_losses, self.model, batch, t, style_batch,
            style_cond["y"]["inpainted_motion"], skip_steps=args.skip_steps, model_kwargs=style_cond,
            model_t2m_kwargs=cond, semantic_guidance=self.semantic_guidance, use_ddim=getattr(args, "use_ddim", 0),
            Ls=getattr(args, "Ls", 10))
        losses = compute_losses()
        if isinstance(self.schedule_sampler, LossAwareSampler):
            self.schedule_sampler.update_with_local_losses(t, losses["loss"].detach())
        loss = (losses["loss"] * weights).mean() if style_batch is None else losses["loss"]
        self.last_losses = {k: v.detach() for k, v in losses.items()}
        self.mp_trainer.backward(loss)

    def _anneal_lr(self):
        if not self.lr_anneal_steps:
            return
        frac_done = (self.step + self.resume_step) / self.lr_anneal_steps
        lr = self.lr * (1 - frac_done)
        for param_group in self.opt.param_groups:
            param_group["lr"] = lr

    # ------------------------------------------------------------------ the epoch loop (reference :143-190)
    def run_loop(self, max_steps=None):
        """``num_epochs`` passes over ``self.data`` (t2m batches) against the style example; checkpoints every
        ``save_interval`` steps and at the end.  ``max_steps`` (not in the reference) bounds the number of steps."""
        iter_style = iter(self.style_data) if self.style_finetune else None
        n = 0
        for _epoch in range(self.num_epochs):
            if self.style_finetune:
                try:
                    content_motion, cond_style = next(iter_style)
                except StopIteration:
                    iter_style = iter(self.style_data)
                    content_motion, cond_style = next(iter_style)
                cond_style['y'] = {k: v.to(self.device) if torch.is_tensor(v) else v for k, v in cond_style['y'].items()}
                if torch.is_tensor(content_motion):
                    content_motion = content_motion.to(self.device)
            else:
                content_motion, cond_style = None, None
            for motion, cond in self.data:
                if self.lr_anneal_steps and self.step + self.resume_step >= self.lr_anneal_steps:
                    break
                motion = motion.to(self.device)
                cond['y'] = {k: v.to(self.device) if torch.is_tensor(v) else v for k, v in cond['y'].items()}
                self.run_step(motion, cond, content_motion, cond_style)   # advances self.step
                done = self.step - 1                                         # the reference saves before its increment
                if self.save_interval and done % self.save_interval == 0:
                    self.save(step=done)
                    if os.environ.get("DIFFUSION_TRAINING_TEST", "") and done > 0:
                        return
                n += 1
                if max_steps is not None and n >= max_steps:
                    break
            if max_steps is not None and n >= max_steps:
                break
            if self.lr_anneal_steps and self.step + self.resume_step >= self.lr_anneal_steps:
                break
        if self.save_interval and self.step > 0 and (self.step - 1) % self.save_interval != 0:
            self.save()  # the last checkpoint if it was not saved already (reference :188-190)


def parse_resume_step_from_filename(filename):
    """path/to/modelNNNNNNNNN.pt -> NNNNNNNNN (reference :352-365)"""
    split = filename.split("model")
    if len(split) < 2:
        return 0
    try:
        return int(split[-1].split(".")[0])
    except ValueError:
        return 0


def find_resume_checkpoint(save_dir, mode='model'):
    """latest ``{mode}NNNNNNNNN.pt`` of a directory (reference :375-383)"""
    files = [f for f in os.listdir(save_dir) if f.endswith('.pt') and f.startswith(mode)]
    steps = [int(f[len(mode):len(mode) + 9]) for f in files]
    return os.path.join(save_dir, f"{mode}{sorted(steps)[-1]:09d}.pt")
