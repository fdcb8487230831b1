#!/usr/bin/env python
"""BASELINE configs[3]: few-shot style finetune step (finetune_style_diffusion loop: fwd + bwd + AdamW) on synthetic data.

    python tools/bench_finetune.py [--steps 10] [--warmup 3] [--batch 64] [--frames 76] [--sg 0|1] [--profile]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_finetune.py ...        (data-parallel over the t2m batch, NCCL all-reduce of the flat gradient arena)

One step = TrainInpaintingLoop.run_step: zero_grad -> few_shot_style_finetune_losses (t2m batch B x T through the
denoiser + MotionEncoder when --sg 1; 6 differentiable DDIM steps of the B=1 style example) -> backward -> norms ->
all-reduce -> fused AdamW.  Prints one JSON line on rank 0 (ms per step = max over ranks, CUDA events)."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def build(dev, B, T, sg, precision="bf16"):
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    from mst_b200.model.mdm_forstyledataset import MDM, MotionEncoder, StyleDiffusion
    from mst_b200.train.training_loop import TrainInpaintingLoop
    from mst_b200.utils import model_util as mu

    class A(bench.Args):
        batch_size, lr, weight_decay, lr_anneal_steps, style_finetune, semantic_guidance = B, 1e-4, 0.0, 0, 1, sg
        skip_steps, use_ddim, Ls, num_steps = 700, 1, 10, 0

    torch.manual_seed(0)
    tmp = tempfile.mkdtemp(prefix="mst_ft_")
    args = A()
    kw = mu.get_transfer_args(args)
    mdm = MDM(load_clip=False, **kw)
    torch.save({k: v for k, v in mdm.state_dict().items() if not k.startswith("clip_model.")}, os.path.join(tmp, "mdm.pt"))
    menc = MotionEncoder(load_clip=False, **kw)
    torch.save({k: v for k, v in menc.state_dict().items() if not k.startswith("mdm_model.")}, os.path.join(tmp, "menc.pt"))
    args.mdm_path, args.semantic_discriminator_path = os.path.join(tmp, "mdm.pt"), os.path.join(tmp, "menc.pt")
    model = StyleDiffusion(load_clip=False, **mu.get_transfer_args(args)).to(dev)
    model.train()
    model.mst_train_precision = precision
    diffusion = mu.create_gaussian_diffusion(args, mu.InpaintingGaussianDiffusion, timestep_respacing="ddim20")
    g = torch.Generator().manual_seed(1)
    F = 181
    x_start = torch.randn(B, F, 1, T, generator=g).to(dev)
    content, style = torch.randn(1, F, 1, T, generator=g).to(dev), torch.randn(1, F, 1, T, generator=g).to(dev)
    lengths = torch.randint(T // 2, T + 1, (B,), generator=g)
    fmask = (torch.arange(T)[None, :] < lengths[:, None])[:, None, None, :].to(dev)
    m1 = torch.from_numpy(get_inpainting_mask("root_horizontal", (1, F, 1, T))).float().to(dev)
    mB = torch.from_numpy(get_inpainting_mask("root_horizontal", (B, F, 1, T))).float().to(dev)
    style_cond = {"y": {"text": ["s"], "text_feat": torch.randn(1, 512, generator=g).to(dev),
                        "mask": torch.ones(1, 1, 1, T, dtype=torch.bool, device=dev), "lengths": torch.tensor([T]),
                        "inpainted_motion": style, "inpainting_mask": m1}}
    cond = {"y": {"text": ["c"] * B, "text_feat": torch.randn(B, 512, generator=g).to(dev), "mask": fmask,
                  "lengths": lengths, "inpainting_mask": mB}}
    loop = TrainInpaintingLoop(args, None, model, [(x_start, cond)], diffusion=diffusion, style_data=((content, style_cond),))
    return loop, (x_start, cond, content, style_cond)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=76)
    ap.add_argument("--sg", type=int, default=1)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = os.dup(1)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        sys.stdout.flush()
        os.dup2(2, 1)  # NCCL's banner / log go to stderr; the JSON line is written to the real stdout below
        dist.init_process_group("nccl", device_id=dev)
    np.random.seed(0)
    from mst_b200 import engine as K
    loop, batch = build(dev, a.batch, a.frames, a.sg, a.precision)
    for _ in range(a.warmup):
        loop.run_step(*batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    h0 = time.perf_counter()
    for _ in range(a.steps):
        loop.run_step(*batch)
    host_ms = (time.perf_counter() - h0) * 1e3 / a.steps  # host time to ISSUE a step (no final sync): ~ value => host-bound
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = (K.launch_count() - n0) // a.steps
    out = {"metric": "finetune_ms_per_step", "value": float(ms.item()), "unit": "ms", "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "higher_is_better": False, "dtype": "bf16" if a.precision == "bf16" else "f32", "data": "synthetic",
           "config": {"workload": f"few-shot style finetune step: t2m batch B={a.batch} x T={a.frames} (sharded over "
                                  f"{world} GPU), style example B=1 x 6 DDIM steps with grad, semantic_guidance={a.sg}, "
                                  "AdamW lr 1e-4 (BASELINE configs[3])"},
           "gpu_launches_per_step": int(launches), "host_issue_ms_per_step": round(host_ms, 3), "loss": float(loop.last_losses["loss"]),
           "grad_norm": loop.mp_trainer.last_norms[0], "param_norm": loop.mp_trainer.last_norms[1]}
    if a.profile and rank == 0:
        with K.profile(cap=8192) as p:
            loop.run_step(*batch)
        agg = {}
        for name, t in p.records:
            c = agg.setdefault(name, [0, 0.0])
            c[0] += 1
            c[1] += t
        tot = sum(v[1] for v in agg.values())
        out["profile_ms"] = round(tot, 3)
        out["stages"] = [{"kernel": k, "launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / tot, 4)}
                         for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]]
    if rank == 0:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
