#!/usr/bin/env python
"""Headline benchmark of the hot path (BASELINE.json: "sampled motion frames/sec at 1/2/4/8 B200; ms per
denoise step (B=64)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One *step* = one complete pass of the hot path over one batch: a full ``--traj-steps``-step (default 1000)
DDPM trajectory of B=64 motions of 196 frames with classifier-free guidance (cond + uncond batched) and the
root_horizontal inpainting blend, i.e. BASELINE.json configs[1], through the public API
(``create_gaussian_diffusion`` -> ``ClassifierFreeSampleModel(MDM)`` -> ``diffusion.p_sample_loop``).
Weights are PyTorch default init under ``torch.manual_seed(0)``, inputs are synthetic (seeds 1-3), CLIP text
features are a random [B,512] tensor (CLIP is outside the hot path).

Printed JSON line (rank 0):
  value          frames/s, whole job: n_gpus * B * T * K / (max-over-ranks device time), inputs resident in HBM
  e2e            same metric with HOST buffers: per step the pinned-host inputs are copied to the device and the
                 final sample is read back, all inside the timed region
  ms_per_step    ms per trajectory; ms_per_denoise_step = that / traj_steps (the second half of BASELINE's metric)
  roofline       dominant kernel (the bf16 tcgen05 GEMM family) measured live with CUDA events between launches
                 (mst_profile_begin/end) on one eager denoise step of the same workload
  cpu_baseline   the CPU oracle port (oracle/) timed on this host's cores on a bounded sample (a few denoise
                 steps of the same B=64 workload), extrapolated to frames/s; rank 0 at N=1 only
``--impl reference`` times that CPU port alone, on the same config / metric / unit.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

F_FEATS, D_MODEL, N_HEADS, D_FF, N_LAYERS = 181, 512, 4, 1024, 8
METRIC, UNIT = "sampled_motion_frames_per_sec", "frames/s"


class Args:
    dataset, latent_dim, layers, cond_mask_prob, arch = "stylexia_posrot", 512, 8, 0.1, "trans_enc"
    emb_trans_dec, unconstrained, diffusion_steps, noise_schedule, sigma_small = False, False, 1000, "cosine", True
    lambda_vel = lambda_rcxyz = lambda_fc = 0.0


def model_flops_per_seq(T, F=F_FEATS, d=D_MODEL, ff=D_FF, L=N_LAYERS):
    """ALGORITHMIC FLOPs of one denoiser evaluation of one sequence: 2MNK of every GEMM incl. QK^T, PV."""
    S = T + 1
    return 4 * T * F * d + 6 * d * d + L * (8 * S * d * d + 4 * S * S * d + 4 * S * d * ff)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: float(m[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["src"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                c = [v.strip() for v in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


# --------------------------------------------------------------------------------------- CPU oracle port
def cpu_port_step_seconds(B, T, n_denoise, threads=None):
    """Seconds per denoise step of the CPU oracle port (oracle/) on this host: CFG forward (2 passes) + the
    p_sample update, B sequences of T frames, timed over `n_denoise` steps after one untimed step."""
    from oracle import denoiser as OD, sampler as OS, schedule as OSch
    from oracle.weights import mdm_state_dict
    # all host cores (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is a single process)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    state = mdm_state_dict(F_FEATS, seed=0)
    g = torch.Generator().manual_seed(1)
    shape = (B, F_FEATS, 1, T)
    x, x_inp, feat = torch.randn(shape, generator=g), torch.randn(shape, generator=g), torch.randn(B, 512, generator=g)
    mask = torch.zeros(shape)
    mask[:, :3] = 1.0
    scale = torch.full((B,), 2.5)
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    times = []
    with torch.no_grad():
        for k in range(n_denoise + 1):
            t = torch.full((B,), 999 - k, dtype=torch.long)
            t0 = time.perf_counter()
            out = OD.cfg_forward(state, x, t, feat, scale)
            x, _ = OS.p_sample(sch, out, x, t, torch.randn(shape, generator=g), mask, x_inp, False)
            if k > 0:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T, N = a.batch, a.frames, a.traj_steps
    per = []
    threads = None
    for k in range(a.warmup + a.steps):
        sec, threads = cpu_port_step_seconds(B, T, 1)
        if k >= a.warmup:
            per.append(sec)
    sec = sum(per) / len(per)
    value = B * T / (N * sec)
    sample = f"1 denoise step of B={B},T={T} CFG+inpainting per bench step, extrapolated x{N} to a full trajectory"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3 * N, "ms_per_denoise_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def workload_config(a, where):
    return {"workload": f"B={a.batch}/GPU x T={a.frames} frames x F={F_FEATS}, {a.traj_steps}-step DDPM p_sample_loop, "
                        f"CFG (cond+uncond batched, scale 2.5) + root_horizontal inpainting, clip_denoised=False, "
                        f"MDM 8L/d512/4h/ff1024 random-init (BASELINE configs[1])",
            "batch_per_gpu": a.batch, "frames": a.frames, "feats": F_FEATS, "traj_steps": a.traj_steps,
            "precision": a.precision if where == "gpu" else "fp32",
            "residual_stream": ("fp16 (LayerNorm outputs), bf16 elsewhere" if a.precision == "bf16" else "fp32") if where == "gpu" else "fp32",
            "rng": "philox-in-kernel" if where == "gpu" else "torch-cpu",
            "trajectory_submission": os.environ.get("MST_TRAJ_GRAPH", "full") + " (whole trajectory = one CUDA-graph launch)" if where == "gpu" else "n/a",
            "l2": "per-denoise-step working set (~250 MB of 16-bit activations at B=64 CFG) exceeds the 126 MB L2; "
                  "a 256 MB buffer is also rewritten between trajectories" if where == "gpu" else "n/a"}


# --------------------------------------------------------------------------------------- native arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=196)
    ap.add_argument("--traj-steps", type=int, default=1000, dest="traj_steps")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling and finetune sub-records")
    ap.add_argument("--strong-batch", type=int, default=4096, dest="strong_batch")
    ap.add_argument("--strong-steps", type=int, default=50, dest="strong_steps")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: park the real stdout and send everything else that any library prints to
    # file descriptor 1 (NCCL's version banner ignores NCCL_DEBUG_FILE, the model factories print) to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL's log (communicator / ring / NVLS lines of NCCL_DEBUG=INFO) is NOT silenced: it is written to fd 1, which
        # points at stderr in this process (see above)
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
    if a.gpus != world and rank == 0:
        print(f"bench.py: --gpus {a.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run for N>1", file=sys.stderr)

    from mst_b200 import engine as K
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    from mst_b200.model.cfg_sampler import ClassifierFreeSampleModel
    from mst_b200.model.mdm_forstyledataset import MDM
    from mst_b200.utils import model_util as mu

    B, T, N = a.batch, a.frames, a.traj_steps
    shape = (B, F_FEATS, 1, T)
    torch.manual_seed(0)
    model = MDM(load_clip=False, **mu.get_transfer_args(Args()))
    model.mst_precision = a.precision
    model.to(dev).eval()
    cfg_model = ClassifierFreeSampleModel(model)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        diffusion = mu.create_gaussian_diffusion(Args(), mu.InpaintingGaussianDiffusion,
                                                 timestep_respacing="" if N == 1000 else str(N))
    diffusion.rng, diffusion.philox_seed, diffusion.philox_sample_offset = "philox", 3, rank * B

    # synthetic inputs in pinned host memory
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    h_inp = torch.randn(shape, generator=g1).pin_memory()
    h_mask = torch.from_numpy(get_inpainting_mask("root_horizontal", shape)).float().pin_memory()
    h_feat = torch.randn(B, 512, generator=g2).pin_memory()
    h_scale = torch.full((B,), 2.5).pin_memory()
    h_out = torch.empty(shape).pin_memory()

    def to_dev():
        return {"y": {"text": [""] * B, "text_feat": h_feat.to(dev, non_blocking=True),
                      "scale": h_scale.to(dev, non_blocking=True),
                      "inpainted_motion": h_inp.to(dev, non_blocking=True),
                      "inpainting_mask": h_mask.to(dev, non_blocking=True),
                      "mask": torch.ones(B, 1, 1, T, device=dev), "lengths": torch.full((B,), T, device=dev)}}

    kw_dev = to_dev()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def trajectory(kw):
        return diffusion.p_sample_loop(cfg_model, shape, clip_denoised=False, model_kwargs=kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            flush.zero_()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up, then `value`: inputs resident in HBM
    for _ in range(max(a.warmup, 3)):
        trajectory(kw_dev)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0, s0 = K.launch_count(), K.graph_replays()
    ms_total = timed(lambda: trajectory(kw_dev), a.steps)
    l1, s1 = K.launch_count(), K.graph_replays()
    clock_rec = clocks.stop() if rank == 0 else None
    launches_per_denoise = diffusion.last_plan_launches
    gpu_launches = (l1 - l0) + (s1 - s0) * launches_per_denoise

    # ---- e2e: host buffers in, host buffer out, every step
    def e2e_step():
        out = trajectory(to_dev())
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    e2e_step()
    ms_e2e = timed(e2e_step, a.steps)
    h2d = sum(t.numel() * t.element_size() for t in (h_inp, h_mask, h_feat, h_scale))
    d2h = h_out.numel() * h_out.element_size()

    frames = world * B * T * a.steps
    value = frames / (ms_total * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    ms_step = ms_total / a.steps

    # ---- roofline: one eager denoise step of the same workload with an event after every launch
    pk = peaks()
    roof, stages = roofline_leg(K, model, dev, B, T, pk, ms_step / N)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_step, "ms_per_denoise_step": ms_step / N, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.precision, "data": "synthetic", "config": workload_config(a, "gpu"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(gpu_launches), "launches_per_denoise_step": int(launches_per_denoise),
        "clocks": clock_rec, "roofline": roof, "stages": stages,
        "model_tflops_per_s_per_gpu": 2 * B * model_flops_per_seq(T) * N / (ms_step * 1e-3) / 1e12,
        "frac_of_bf16_sustained_peak": 2 * B * model_flops_per_seq(T) * N / (ms_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
        "peaks": pk,
    }
    # ---- BASELINE configs[2] / configs[3] inside the same line, so the driver's 1/2/4/8 runs carry them
    if not a.no_extra:
        if os.environ.get("MST_BENCH_FT_DEBUG"):  # diagnosis: the finetune leg before AND after the strong-scaling leg
            line["finetune_before_strong"] = finetune_leg(a, dev, rank, world, dist)["value"]
        try:
            line["strong"] = strong_scaling_leg(a, dev, rank, world, dist, cfg_model, mu, get_inpainting_mask)
        except Exception as exc:
            line["strong"] = {"error": str(exc)[:300]}
        try:
            line["finetune"] = finetune_leg(a, dev, rank, world, dist)
        except Exception as exc:
            line["finetune"] = {"error": str(exc)[:300]}
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            sec, threads = cpu_port_step_seconds(B, T, 3)
            line["cpu_baseline"] = {"value": B * T / (N * sec), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"3 denoise steps of the same B={B},T={T} CFG+inpainting workload on the CPU "
                                              f"oracle port ({sec:.2f} s/step), extrapolated x{N}",
                                    "host_cpus": os.cpu_count()}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def strong_scaling_leg(a, dev, rank, world, dist, cfg_model, mu, get_inpainting_mask):
    """BASELINE configs[2]: a FIXED total batch (default 4096 motions x 196 frames, CFG + inpainting) sharded over the
    ranks (contiguous shards, Philox keyed by the global sample index, no collective in the loop), on a respaced
    trajectory of --strong-steps steps so that the default run stays short.  value = frames of COMPLETE
    strong_steps-step trajectories per second over all ranks (max-over-ranks device time)."""
    import contextlib
    import io
    total, n = a.strong_batch, a.strong_steps
    if total % world:
        raise ValueError(f"strong batch {total} is not a multiple of {world} ranks")
    B, T = total // world, a.frames
    shape = (B, F_FEATS, 1, T)
    with contextlib.redirect_stdout(io.StringIO()):
        d = mu.create_gaussian_diffusion(Args(), mu.InpaintingGaussianDiffusion, timestep_respacing=str(n))
    d.rng, d.philox_seed, d.philox_sample_offset = "philox", 3, rank * B
    g = torch.Generator().manual_seed(100 + rank)
    kw = {"y": {"text": [""] * B, "text_feat": torch.randn(B, 512, generator=g).to(dev), "scale": torch.full((B,), 2.5, device=dev),
                "inpainted_motion": torch.randn(shape, generator=g).to(dev),
                "inpainting_mask": torch.from_numpy(get_inpainting_mask("root_horizontal", shape)).float().to(dev),
                "mask": torch.ones(B, 1, 1, T, device=dev), "lengths": torch.full((B,), T, device=dev)}}
    run = lambda: d.p_sample_loop(cfg_model, shape, clip_denoised=False, model_kwargs=kw)
    run()
    reps = 2
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / reps
    del kw, d
    torch.cuda.empty_cache()
    flops = 2 * total * model_flops_per_seq(T) * n
    return {"metric": METRIC, "scaling": "strong", "total_batch": total, "batch_per_gpu": B, "traj_steps": n, "n_gpus": world,
            "ms_per_trajectory": ms, "ms_per_denoise_step": ms / n, "value": total * T / (ms * 1e-3), "unit": UNIT,
            "frames_per_s_1000_step_equiv": total * T / (ms * 1e-3) * n / 1000.0,
            "model_tflops_per_s_all_gpus": flops / (ms * 1e-3) / 1e12,
            "frac_of_bf16_sustained_peak": flops / (ms * 1e-3) / 1e12 / (world * peaks()["bf16_tflops_sustained"]),
            "workload": f"B={total} total ({B}/GPU) x T={T}, {n}-step respaced DDPM, CFG + inpainting (BASELINE configs[2])"}


def finetune_leg(a, dev, rank, world, dist, B=64, T=76, steps=20, warmup=5):
    """BASELINE configs[3]: the few-shot style finetune step (t2m batch B=64 x T=76 sharded over the ranks, style example
    B=1 x 6 differentiable DDIM steps replicated, semantic guidance on, fused AdamW; NCCL all-reduce of the 67 MB fp32
    gradient arena when world > 1).  ms per step = max over ranks (CUDA events); the all-reduce is timed on the device."""
    import numpy as np
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import bench_finetune as BF
    import contextlib
    np.random.seed(0)
    with contextlib.redirect_stdout(sys.stderr):  # the model factories print; stdout carries the ONE JSON line
        loop, batch = BF.build(dev, B, T, 1, "bf16")
    for _ in range(warmup):
        loop.run_step(*batch)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    # The step is ~9 ms with ~5.5 ms of host work: one full collection of Python's cyclic GC inside the 8 timed steps (this
    # process holds the sampler's models and graphs by now) shows up as +3-4 ms per step.  Collect now, keep the collector
    # out of the timed region (a training script would gc.freeze() after setup for the same reason).
    import gc
    gc_tweak = not os.environ.get("MST_BENCH_NO_GC_TWEAK")
    if gc_tweak:
        gc.collect()
        gc.disable()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exposed, window = [], []
    e0.record()
    for _ in range(steps):
        loop.run_step(*batch)
        if world > 1:
            ov = loop.last_overlap()
            if ov is not None:
                window.append(ov[0])
                exposed.append(ov[1])
    e1.record()
    torch.cuda.synchronize(dev)
    if gc_tweak:
        gc.enable()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"metric": "finetune_ms_per_step", "value": float(ms.item()), "unit": "ms", "n_gpus": world, "higher_is_better": False,
           "scaling": "strong (t2m batch sharded, style term replicated)", "loss": float(loop.last_losses["loss"]),
           "workload": f"t2m B={B} x T={T} + style B=1 x 6 DDIM steps with grad, semantic_guidance=1, AdamW (BASELINE configs[3])"}
    if world > 1:
        # the collective alone (same 67 MB buffer, blocking, CUDA events), then how much of it the step hides
        buf = torch.zeros_like(loop.mp_trainer.flat.grads)
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize(dev)
        ar_ms = a0.elapsed_time(a1) / 10
        nbytes = buf.numel() * 4
        exp_ms = sum(exposed) / len(exposed) if exposed else ar_ms
        out.update(allreduce_ms=ar_ms, allreduce_bytes=nbytes,
                   allreduce_busbw_gbs=2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                   allreduce_exposed_ms=exp_ms, allreduce_overlap_pct=100.0 * max(0.0, 1.0 - exp_ms / ar_ms),
                   allreduce_window_ms=sum(window) / len(window) if window else 0.0,
                   allreduce_mode="t2m gradient all-reduced asynchronously behind the replicated style term" if exposed
                   else "one blocking all-reduce of the gradient arena after the backward pass",
                   allreduce_share_of_step=ar_ms / float(ms.item()))
        del buf
    del loop, batch
    torch.cuda.empty_cache()
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the four GEMM launches of one layer, from one
# `ncu --set full` capture of the B=64, T=196 step; re-derived every round (tools/ncu_raw_summary.py on the capture)
NCU_TRAFFIC_B64 = {(64, 196): 55.7e6}
NCU_TRAFFIC_SRC = ("profiles/r02h_ncu_full_step_kernels_summary.txt (round-2 capture: dram read + write bytes per launch, mean of the "
                   "4 GEMM launches of a layer: QKV 53.1 MB, out-proj+LN 55.2 MB, FFN1+GELU 30.1 MB, FFN2+LN 84.5 MB)")


def roofline_leg(K, model, dev, B, T, pk, ms_denoise_in_graph, steady=True):
    """Per-launch CUDA-event timing of one denoise step (forward with cond+uncond batched, then the fused update),
    captured as a CUDA graph with an event after every kernel and replayed in steady state -> roofline of the
    dominant kernel family + the per-kernel table (each entry = kernel time + the in-graph gap to the next one)."""
    from mst_b200 import _lib as L
    eng = model.mst_engine(dev)
    shape = (B, F_FEATS, 1, T)
    x = torch.randn(shape, device=dev)
    temb = eng.time_embed(torch.arange(1000, device=dev))
    text = eng.text_embed(torch.randn(B, 512, device=dev))
    t_dev = torch.tensor([500], dtype=torch.int32, device=dev)
    oc, ou = torch.empty_like(x), torch.empty_like(x)
    tab = torch.rand(1000, device=dev)
    scale = torch.full((B,), 2.5, device=dev)
    mask = torch.zeros(F_FEATS, device=dev)
    mask[:3] = 1

    def step():
        eng.forward(x, temb, text, cfg=True, out_cond=oc, out_uncond=ou, temb_row_dev=t_dev)
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=x, mask=mask,
                      x_inpaint=oc, t_scalar_dev=t_dev, coef1=tab, coef2=tab, sigma=tab, noise_kind=L.NOISE_PHILOX,
                      philox_seed=1)

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    # capture the step with an event-record node after every kernel, replay it long enough to reach the
    # steady (power-capped) state of a real trajectory, then read the per-launch times of the last replays
    agg = {}
    reps = 5
    if not steady:  # eager variant (for runs under ncu): events between eager launches
        for _ in range(reps):
            with K.profile() as p:
                step()
            for name, ms in p.records:
                n, tot = agg.get(name, (0, 0.0))
                agg[name] = (n + 1, tot + ms)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            with K.profile(deferred=True) as p:
                step()
    torch.cuda.current_stream(dev).wait_stream(side)
    for _ in range(300 if steady else 0):
        graph.replay()
    for _ in range(reps if steady else 0):
        graph.replay()
        torch.cuda.synchronize(dev)
        for name, ms in p.collect():
            n, tot = agg.get(name, (0, 0.0))
            agg[name] = (n + 1, tot + ms)
    S, M = T + 1, 2 * B * (T + 1)
    d, ff, F = D_MODEL, D_FF, F_FEATS
    flops = {  # algorithmic FLOPs per launch
        "tc_gemm_qkv": 2 * M * 3 * d * d, "tc_gemm_ffn1_gelu": 2 * M * ff * d,
        "tc_gemm_res_ln": 2 * M * d * (d + ff) / 2,  # out-proj (K=d) and FFN2 (K=ff) launches alternate
        "tc_gemm_inproj": 2 * B * T * F * d, "tc_gemm_outproj": 2 * M * F * d,
        "tc_attention": 2 * B * N_HEADS * 4 * S * S * (d // N_HEADS) / 1.0 / 1.0 * 1.0,
        "gemm_f32": None, "attention_f32": None,
    }
    stages, total = [], sum(tot for _, tot in agg.values()) / reps
    for name, (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        avg = tot / n
        rec = {"kernel": name, "launches_per_step": n // reps, "avg_ms": avg, "share": (tot / reps) / total}
        if flops.get(name):
            rec["tflops"] = flops[name] / (avg * 1e-3) / 1e12
        stages.append(rec)
    tc = [s for s in stages if s["kernel"].startswith("tc_gemm") and "tflops" in s]
    roof = None
    if tc:
        # dominant kernel = the tcgen05 GEMM (one kernel template, several epilogues): flop-weighted over its launches
        fl = sum(flops[s["kernel"]] * s["launches_per_step"] for s in tc)
        ms = sum(s["avg_ms"] * s["launches_per_step"] for s in tc)
        achieved = fl / (ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "tc_gemm_kernel<BN,EPI> (all epilogues)", "achieved": achieved,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                "peak_kind": f"{pk['src']} sustained cuBLAS bf16 (kernel timed inside a long step)",
                "traffic": NCU_TRAFFIC_B64.get((B, T)), "traffic_src": NCU_TRAFFIC_SRC if (B, T) in NCU_TRAFFIC_B64 else None,
                "share_of_step": ms / total, "profiled_step_ms": total, "graph_step_ms": ms_denoise_in_graph}
    if roof is not None:
        # the secondary bound that explains the tensor fraction: every tile design here stages 128 A rows + 128 W rows
        # per CTA and k-block = 64 B/clk/SM of shared-memory fill at the full MMA rate, read back by the tensor core at
        # another 64 B/clk/SM (tools/ubench/tma_fill.cu: the fill alone tops out at 60.5 B/clk/SM from L2)
        roof["fabric_note"] = ("ncu (profiles/r02h_*): crossbar->L1 at 28 %, LTS->crossbar at 36 % of their peaks in the QKV GEMM - "
                               "the L2 fabric is not the bound (round 1's '0.66 L2 cap' is withdrawn); measured bulk-copy fill "
                               "rate 60.5 B/clk/SM from L2 (profiles/r02p_ubench_tma_fill_rate.txt) against 64 needed at the "
                               "full MMA rate, plus 64 B/clk/SM of operand reads: epilogue shared-memory traffic comes out of the "
                               "MMAs' share (LN GEMMs: probes in DESIGN.md section 4); an A-stationary QKV variant with half "
                               "the fill traffic measured no faster, so QKV at 0.70 is not fill-bound")
    upd = [s for s in stages if s["kernel"] == "update"]
    if upd and roof is not None:
        by = 20 * B * F_FEATS * T  # out_c, out_u, x_t, x_inp read + x_{t-1} write, [F] mask, in-kernel Philox
        roof["update_kernel"] = {"bound": "hbm", "achieved": by / (upd[0]["avg_ms"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                 "unit": "GB/s", "frac": by / (upd[0]["avg_ms"] * 1e-3) / 1e9 / pk["hbm_gbs"],
                                 "bytes_per_launch": by, "note": "L2-resident at B=64 (45 MB); HBM-bound only at B>=2048"}
    if roof is not None and "update_kernel" in roof:
        try:
            roof["update_kernel"]["hbm_bound"] = update_kernel_hbm_leg(K, dev, pk)
        except Exception as exc:  # an out-of-memory on a shared box must not lose the headline number
            roof["update_kernel"]["hbm_bound"] = {"error": str(exc)[:200]}
    return roof, stages


def update_kernel_hbm_leg(K, dev, pk, B=2048, T=196):
    """The fused update kernel alone at a size where it IS HBM-bound (B=2048: 1.45 GB per launch, 11x the L2), CUDA events
    around the launches, live: achieved algorithmic GB/s (20 B/element) against the measured HBM copy peak."""
    from mst_b200 import _lib as L
    shape = (B, F_FEATS, 1, T)
    oc, ou, x, xi = (torch.randn(shape, device=dev) for _ in range(4))
    out = torch.empty_like(x)
    scale = torch.full((B,), 2.5, device=dev)
    mask = torch.zeros(F_FEATS, device=dev)
    mask[:3] = 1
    tab = torch.rand(1000, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)

    def run():
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=out, mask=mask,
                      x_inpaint=xi, mask_noise=True, clip_denoised=False, t_vec=t, coef1=tab, coef2=tab, sigma=tab,
                      noise_kind=L.NOISE_PHILOX, philox_seed=3)
    for _ in range(3):
        run()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    by = 20 * B * F_FEATS * T
    gbs = by / (ms * 1e-3) / 1e9
    del oc, ou, x, xi, out
    torch.cuda.empty_cache()
    return {"batch": B, "bytes_per_launch": by, "ms": ms, "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": gbs / pk["hbm_gbs"], "frac_of_8TBps_nominal": gbs / 8000.0,
            "note": "inputs (5.8 GB live, 1.45 GB per launch) exceed the 126 MB L2; back-to-back launches, CUDA events"}


if __name__ == "__main__":
    main()
