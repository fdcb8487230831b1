"""Shared helpers of the test-suite."""
import hashlib

import numpy as np
import torch


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def relerr(a, b):
    """max |a-b| / max |b|  - the 'relative error' every tolerance in this suite refers to"""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class Args:
    """The argparse namespace fields the reference's factories read (utils/model_util.py)."""
    dataset = "stylexia_posrot"
    latent_dim = 512
    layers = 8
    cond_mask_prob = 0.1
    arch = "trans_enc"
    emb_trans_dec = False
    unconstrained = False
    diffusion_steps = 1000
    noise_schedule = "cosine"
    sigma_small = True
    lambda_vel = 0.0
    lambda_rcxyz = 0.0
    lambda_fc = 0.0
