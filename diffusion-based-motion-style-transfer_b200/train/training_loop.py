"""Few-shot style finetune loop (reference ``train/training_loop.py:36-303``) on the hand-written CUDA path.

Kept from the reference: the constructor's argument names, ``run_step`` / ``forward_backward`` / ``_anneal_lr`` and
their order of operations (zero_grad -> few_shot_style_finetune_losses -> backward -> norms -> AdamW -> lr anneal).
Not kept: logger / checkpoint / dataset plumbing (control plane, out of scope - see DESIGN.md).

Data parallelism: when ``torch.distributed`` is initialised with more than one rank, every rank
takes a contiguous shard of the text-to-motion batch (the only batched term of the loss), keeps the B=1 style term
replicated, and the flat gradient arena is summed with ONE all-reduce (NCCL over NVLink on GPUs) before the fused
AdamW step applies it scaled by 1/world_size.
"""
from __future__ import annotations

import functools

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

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count,
                "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])
        self.param_groups = sd["param_groups"]


def shard_batch(batch, cond, rank, world):
    """Contiguous shard of the t2m batch for this rank: tensors with a leading batch dimension and lists of
    per-sample entries are sliced; everything else is shared."""
    B = batch.shape[0]
    per = (B + world - 1) // world
    lo, hi = min(B, rank * per), min(B, (rank + 1) * per)
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
        self.step = 0
        self.resume_step = 0
        self.num_steps = getattr(args, "num_steps", 0)
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
        self.device = next(model.parameters()).device
        self.last_losses = None

    # ------------------------------------------------------------------ one optimisation step
    def run_step(self, batch, cond, style_batch=None, style_cond=None):
        self.forward_backward(batch, cond, style_batch, style_cond)
        self.sync_gradients()
        self.mp_trainer.optimize(self.opt)
        self._anneal_lr()
        self.step += 1

    def sync_gradients(self):
        if self.use_ddp:
            dist.all_reduce(self.mp_trainer.flat.grads, op=dist.ReduceOp.SUM)

    def forward_backward(self, batch, cond, style_batch, style_cond):
        self.mp_trainer.zero_grad()
        assert self.diffusion is not None
        if not self.style_finetune:
            raise NotImplementedError("only the style-finetune objective is built (few_shot_style_finetune_losses); "
                                      "the reference's generic training_losses path is outside the scope table")
        if self.use_ddp:
            batch, cond, _ = shard_batch(batch, cond, self.rank, self.world)
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

    # ------------------------------------------------------------------ a plain loop over the given iterables
    def run_loop(self, max_steps=None):
        """Iterate ``self.data`` (t2m batches) against the single style example (training_loop.py:143-190) without the
        reference's logging / checkpoint side effects."""
        iter_style = iter(self.style_data)
        content_motion, cond_style = next(iter_style)
        n = 0
        while max_steps is None or n < max_steps:
            for motion, cond in self.data:
                if self.lr_anneal_steps and self.step + self.resume_step >= self.lr_anneal_steps:
                    return
                motion = motion.to(self.device)
                cond['y'] = {k: v.to(self.device) if torch.is_tensor(v) else v for k, v in cond['y'].items()}
                self.run_step(motion, cond, content_motion, cond_style)
                n += 1
                if max_steps is not None and n >= max_steps:
                    return
            if max_steps is None and self.num_steps and self.step >= self.num_steps:
                return
