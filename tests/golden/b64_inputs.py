"""Seeded inputs of the B=64, T=196 headline-shape golden (shared by make_golden_b64.py and the tests)."""
import numpy as np
import torch

from oracle.weights import text_features

KEEP = [0, 21, 42, 63]


def b64_inputs(B=64, F=181, T=196):
    shape = (B, F, 1, T)
    g = torch.Generator().manual_seed(1)
    x_inp = torch.randn(shape, generator=g)
    texts = [f"a person performs motion number {i}" for i in range(B)]
    mask = torch.zeros(shape)
    mask[:, 0:3] = 1.0  # root_horizontal: rows {0, 1, 2} (checked against the reference's mask in make_golden_b64.py)
    return dict(shape=shape, x_inp=x_inp, texts=texts, feat=text_features(texts), scale=torch.full((B,), 2.5), mask=mask)


def b64_digests(x0):
    """three per-sample digests of an x0 prediction [B,F,1,T] (float64 accumulation)"""
    x = x0.double().flatten(1)
    g = torch.Generator().manual_seed(77)
    probe = torch.randn(x.shape[1], generator=g, dtype=torch.float64)
    return {"sum": x.sum(1), "l2": x.norm(dim=1), "probe": x @ probe}
